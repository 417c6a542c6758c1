"""UpstreamSyncer drift repair (internal/controller/upstreamsyncer_controller.go:77-159) over the in-memory
cluster.  KATs: internal/controller/upstreamsyncer_controller_test.go:329 (start tracking), :387 (within the
grace period), :450 (grace exceeded -> detach CR created, first reconcile copies the device id)."""

DEV = "GPU-device00-uuid-temp-0000-000000000000"
RES = "GPU-device00-uuid-temp-0000-000000000res"
UP = [{"node_name": "worker-0", "machine_uuid": "machine0-uuid-temp-0000-000000000000", "device_type": "gpu",
       "model": "NVIDIA-A100-PCIE-80GB", "device_id": DEV, "cdi_device_id": RES}]
T0 = 1_800_000_000


def _detach_crs(dump):
    return {n: r for n, r in dump["resources"].items() if r["labels"].get("cohdi.io/ready-to-detach-device-id") == DEV}


def test_first_sighting_starts_tracking(cro):                      # :329, isCreated false
    with cro.Cluster({"nodes": ["worker-0"], "uuids": [DEV]}) as c:
        assert c.sync_upstream(UP, T0) == ""
        d = c.dump()
        assert _detach_crs(d) == {} and d["missing_devices"] == {DEV: T0}


def test_within_grace_period_waits(cro):                           # :387, isCreated false
    with cro.Cluster({"nodes": ["worker-0"], "uuids": [DEV]}) as c:
        c.sync_upstream(UP, T0)
        assert c.sync_upstream(UP, T0 + 5 * 60) == ""
        d = c.dump()
        assert _detach_crs(d) == {} and d["missing_devices"] == {DEV: T0}
        assert c.sync_upstream(UP, T0 + 10 * 60) == ""              # time.Since(first) > grace is strict
        assert _detach_crs(c.dump()) == {}


def test_grace_exceeded_creates_detach_cr(cro):                    # :450, isCreated true
    with cro.Cluster({"nodes": ["worker-0"], "uuids": [DEV]}) as c:
        c.sync_upstream(UP, T0 - 20 * 60)
        assert c.sync_upstream(UP, T0) == ""
        d = c.dump()
        crs = _detach_crs(d)
        assert len(crs) == 1 and d["missing_devices"] == {}
        (name, cr), = crs.items()
        assert name.startswith("gpu-") and len(name) == 45             # generateName "gpu-<uuid4>" + 5 chars
        assert cr["labels"]["cohdi.io/ready-to-detach-cdi-device-id"] == RES
        assert cr["spec"] == {"type": "gpu", "model": "NVIDIA-A100-PCIE-80GB", "target_node": "worker-0"}
        # the test then triggers ONE reconcile and expects Status.DeviceID to be the device (handleNoneState :186-193)
        assert c.reconcile_resource(name) == ""
        st = c.dump()["resources"][name]["status"]
        assert st == {"state": "Attaching", "device_id": DEV, "cdi_device_id": RES}
        # left to itself the CR walks Attaching -> Online -> (label) delete -> Detaching -> Deleting -> gone
        c.run()
        d = c.dump()
        assert d["resources"] == {} and d["missing_devices"] == {}


def test_owned_device_is_not_drift_and_vanished_device_is_forgotten(cro):
    with cro.Cluster({"nodes": ["worker-0"], "uuids": [DEV]}) as c:
        c.plant({"kind": "ComposableResource", "name": "gpu-owned", "spec": {"type": "gpu", "model": "m", "target_node": "worker-0"},
                 "status": {"state": "Online", "device_id": DEV}})
        c.sync_upstream(UP, T0)
        assert c.dump()["missing_devices"] == {}
        other = [dict(UP[0], device_id="GPU-other")]
        c.sync_upstream(other, T0)
        assert c.dump()["missing_devices"] == {"GPU-other": T0}
        c.sync_upstream([], T0 + 60)                                 # gone upstream: stop tracking (:123-133)
        assert c.dump()["missing_devices"] == {}


def test_bad_upstream_payload(cro):
    with cro.Cluster({"nodes": ["worker-0"]}) as c:
        assert c.sync_upstream({"not": "a list"}, T0).startswith("failed to fetch data from upstream server")


def test_syncer_fed_by_the_real_cm_client(cro):
    """The whole tick the way the reference's entries run it (upstreamsyncer_controller_test.go:329-515): the CM client
    lists the fabric's devices (cluster ...0001 holds DEV), the syncer tracks the orphan and, past the grace period,
    creates the detach CR.  Note the CM flavour reports no model (cm/client.go:335-341), so the CR's Spec.Model is ""."""
    from test_reference_entries import ENTRIES, HTTP, objects_for
    env = {"DEVICE_RESOURCE_TYPE": "DRA", "CDI_PROVIDER_TYPE": "FTI_CDI", "FTI_CDI_API_TYPE": "CM",
           "FTI_CDI_TENANT_ID": "tenant00-uuid-temp-0000-000000000000", "FTI_CDI_CLUSTER_ID": "cluster0-uuid-temp-0000-000000000001"}
    listed = cro.fabric_list_devices({"env": env, "fabric": {"http": HTTP, "objects": objects_for(ENTRIES[1739])}})
    assert listed["error"] == "" and [d["device_id"] for d in listed["devices"]] == [DEV]
    with cro.Cluster({"nodes": ["worker-0"], "uuids": [DEV]}) as c:
        assert c.sync_upstream(listed["devices"], T0 - 20 * 60) == ""
        assert c.dump()["missing_devices"] == {DEV: T0 - 20 * 60}
        assert c.sync_upstream(listed["devices"], T0) == ""
        (name, cr), = _detach_crs(c.dump()).items()
        assert cr["spec"] == {"type": "gpu", "model": "", "target_node": "worker-0"}
        assert cr["labels"]["cohdi.io/ready-to-detach-cdi-device-id"] == RES
