"""The reference's ComposableResource table tests, entry by entry, with the REAL provider clients.

tests/golden/reference_entries.json (made by tests/golden/make_reference_entries.py from
internal/controller/composableresource_controller_test.go) holds, per Entry, what the reference
asserts — expectedReconcileError, the expected Status, or expectedRequestDeleted — plus the switches
that select a scenario: tenant/cluster uuid, the BareMetalHost's machine uuid, which metal3 objects
exist, the secret's username.  Its "routes" table is the reference's fake fabric server
(:663-930) as data.  Here every entry is replayed through cro_reconcile_attach with
  env     = the env vars the Describe block sets (adapter selection, composableresource_adapter.go:39-72),
  fabric  = those routes + those objects, served to csrc/provider.cpp's FM / CM client,
and the node-side inputs (what nvidia-smi printed, which DaemonSets exist) the entry's mocks return.
Citations are ":<line of the Entry>".  Entries other test modules already pin are replayed here too:
this module drives the full client, those drive the canned provider."""
import json
import os
import random

import pytest

import __graft_entry__ as g
from test_cm_provider import cm_machine_data
from test_fabric_codec import fm_machine_data

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "reference_entries.json")))
KATS = json.load(open(os.path.join(HERE, "golden", "reference_kats.json")))
ENTRIES = {e["line"]: e for e in GOLD["entries"]}

DEV, RES = "GPU-device00-uuid-temp-0000-000000000000", "GPU-device00-uuid-temp-0000-000000000res"
FAILDEV, FAILRES = "GPU-device00-uuid-temp-fail-000000000000", "GPU-device00-uuid-temp-fail-000000000res"
MODEL = "NVIDIA-A100-PCIE-80GB"
DRA_DS = "nvidia-dra-driver-gpu/nvidia-dra-driver-gpu-kubelet-plugin"
DP_DS, DCGM_DS = "nvidia-gpu-operator/nvidia-device-plugin-daemonset", "nvidia-gpu-operator/nvidia-dcgm"
READY = {"desired": 1, "ready": 1, "current": 1, "unavailable": 0, "misscheduled": 0}
NOW = "2025-06-01T12:00:00Z"


# ---- the fake fabric: golden route table -> scripted "http" rules -----------------------------
def _generated(gen, args):
    if gen == "generateCMMachineData":                      # (isAttachFailed, isDetachFailed, isSucceeded, isCompleteButError)
        attach_failed, detach_failed, ok, op = args
        devs = None
        if attach_failed:
            devs = [(FAILDEV, "ADD_FAILED", "add failed due to some reasons", FAILRES, "0")]
        if detach_failed:
            devs = [(FAILDEV, "REMOVE_FAILED", "remove failed due to some reasons", FAILRES, "0")]
        if ok:
            devs = [(DEV, "ADD_COMPLETE", "", RES, "0")]
        if op is not None:
            devs = [(DEV, "ADD_COMPLETE", "", RES, op)]
        return cm_machine_data(devs)
    if gen == "generateFMUpdateData":
        return KATS["fixtures"]["fm_update_body"][args[0] or "none"]
    if gen == "generateFMMachineData":
        op = {"isNormal": "0", "isWarning": "1", "isCritical": "2", "isUnknown": "3"}.get(args[0])
        return fm_machine_data([] if op is None else [(RES, "gpu", op, DEV, MODEL)])
    raise AssertionError(gen)


def fabric_http():
    rules = []
    for path, r in GOLD["routes"].items():
        body = _generated(r["generator"], r["args"]) if "generator" in r else r.get("body", "")
        rules.append({"path": path[1:], "status": r["status"], "body": body})
    rules.append({"path_contains": "", "status": 404, "body": '{"error":"not found"}'})     # the handler's default case
    return rules


HTTP = fabric_http()
NOW_UNIX = 1748779200       # NOW ("2025-06-01T12:00:00Z") below


def _go_encode_map(fields):
    """json.NewEncoder(w).Encode(map[string]interface{}): keys sorted, HTML-safe escapes, trailing newline."""
    import json as _j
    s = _j.dumps(fields, sort_keys=True, separators=(",", ":"), ensure_ascii=False)
    return s.replace("<", "\\u003c").replace(">", "\\u003e").replace("&", "\\u0026") + "\n"


def _b64url(raw):
    import base64
    return base64.urlsafe_b64encode(raw).rstrip(b"=").decode()


def id_manager_reply(user, now=NOW_UNIX):
    import json as _j
    r = GOLD["id_manager"].get(user) or {"status": 400, "body": '{"error":"unsupported_test_user"}'}
    if "body" in r:
        return {"status": r["status"], "body": r["body"]}
    fields = {}
    for k, v in r["encode"].items():
        if isinstance(v, dict):
            p = v["payload"]
            if "raw" in p:
                mid = _b64url(p["raw"].encode())
            else:            # createTokenPayload (:617-627)
                mid = _b64url(_j.dumps({"admin": True, "exp": now + 3600 * p["claims_exp_in_hours"], "iat": now,
                                        "name": "John Doe", "sub": "1234567890"}, separators=(",", ":")).encode())
            v = v["prefix"] + mid + v["suffix"]
        fields[k] = v
    return {"status": r["status"], "body": _go_encode_map(fields)}


def objects_for(e, bmh_uuid=None):
    o, objs = e["objects"], {}
    if o["node"]:
        objs["nodes"] = {"worker-0": {"annotations": {"machine.openshift.io/machine": "openshift-machine-api/machine-worker-0"}
                                      if o["node_annotation"] else {}}}
    if o["machine"]:
        objs["metal3machines"] = {"openshift-machine-api/machine-worker-0": {
            "annotations": {"metal3.io/BareMetalHost": "openshift-machine-api/bmh-worker-0"} if o["machine_annotation"] else {}}}
    if o["bmh"]:
        uuid = bmh_uuid or o.get("bmh_machine_uuid")
        objs["baremetalhosts"] = {"openshift-machine-api/bmh-worker-0": {
            "annotations": {"cluster-manager.cdi.io/machine": uuid} if uuid else {}}}
    return objs


def env_for(e):
    ctx = " ".join(e["context"])
    if "FM and DEVICE_PLUGIN" in ctx:
        api, drt = "FM", "DEVICE_PLUGIN"
    else:
        api, drt = "CM", "DRA"
    return {"DEVICE_RESOURCE_TYPE": drt, "CDI_PROVIDER_TYPE": "FTI_CDI", "FTI_CDI_API_TYPE": api,
            "FTI_CDI_TENANT_ID": e.get("tenant_uuid", ""), "FTI_CDI_CLUSTER_ID": e.get("cluster_uuid", "")}


def state_of(e):
    ctx = " ".join(e["context"])
    for s in ("Attaching", "Online", "Detaching", "Deleting", "None"):
        if "in %s state" % s in ctx:
            return "" if s == "None" else s
    raise AssertionError(ctx)


def request_for(e, **extra):
    init = e.get("initial_status") or {}
    # what the fake id_manager answers this entry's Secret (:665-716); a missing Secret is the clientset's NotFound
    token = id_manager_reply(e.get("username", "good_user"))
    if not e["objects"]["secret"]:
        token = {"secret_error": 'secrets "credentials" not found'}
    r = {"name": e.get("resourceName", "test-composable-resource"),
         "spec": {"type": "gpu", "model": MODEL, "target_node": "worker-0"},
         "status": {"state": state_of(e), "error": init.get("Error", ""), "device_id": init.get("DeviceID", ""),
                    "cdi_device_id": init.get("CDIDeviceID", "")},
         "deleting": False, "probe": False, "env": env_for(e), "now": NOW,
         "fabric": {"http": HTTP, "objects": objects_for(e), "token": token},
         "enumeration": {"stdout": "", "stderr": ""}, "resource_slices": []}
    r.update(extra)
    return r


def expected_status(e, oracle):
    s = e["expected_status"]
    return oracle.emit_status(s.get("State", ""), s.get("Error", ""), s.get("DeviceID", ""), s.get("CDIDeviceID", ""))


def check(cro, oracle, e, out):
    if "expected_error" in e:
        assert out["error"] == e["expected_error"], (e["line"], out["error"])
        assert out["status"].get("error", "") == e["expected_error"], e["line"]     # requeueOnErr records it (:423-433)
        assert out["status"]["state"] == state_of(e), e["line"]
    else:
        assert out["error"] == "", (e["line"], out["error"])
        assert g.json_status(out) == expected_status(e, oracle), (e["line"], out["_raw"])


# ---- 1. entries whose outcome is decided inside the provider client -----------------------------
PROVIDER_SIDE = [
    # Attaching, CM: metal3 walk, token, GET machine, POST resize
    1194, 1219, 1246, 1281, 1319, 1365, 1414, 1479, 1544, 1609, 1674, 1739, 1804, 1869, 1934, 1999, 2064, 2133,
    # Attaching, FM: metal3 walk, PATCH update and the op-status gate
    6122, 6146, 6173, 6208, 6246, 6292, 6357, 6422, 6487, 6552, 6617,
    # Online: CheckResource over GET machine
    3446, 3485, 3562, 3639, 3716, 3793, 3870, 3962, 7671, 7710, 7787, 7864, 7941, 8017, 8095, 8172, 8249,
]


@pytest.mark.parametrize("line", PROVIDER_SIDE, ids=[":%d" % n for n in PROVIDER_SIDE])
def test_provider_side_entries(cro, oracle, line):
    e = ENTRIES[line]
    out = cro.reconcile_attach(None, request_for(e))
    check(cro, oracle, e, out)
    if state_of(e) == "Online":
        assert out["requeue_after_s"] == 30                                           # :305-317
    if line == 2064:      # "should wait when the GPU has not yet been added in CM": the resize request went out
        assert out["requeue_after_s"] == 30
        last = out["fabric_requests"][-1]
        assert (last["method"], last["path"].rsplit("/", 2)[-2:]) == ("POST", ["actions", "resize"])
        assert last["body"] == oracle.emit_cm_scale_up("spec0000-uuid-temp-0000-000000000002", 1)


def test_fm_requests_on_the_wire(cro, oracle):
    """:6487 replayed for what went out: PATCH .../machines/<uuid>/update?tenant_uuid=..., ScaleUpBody bytes."""
    e = ENTRIES[6487]
    out = cro.reconcile_attach(None, request_for(e))
    (req,) = out["fabric_requests"]
    mid = e["objects"]["bmh_machine_uuid"]
    assert req == {"method": "PATCH", "path": "fabric_manager/api/v1/machines/%s/update" % mid,
                   "query": "tenant_uuid=" + e["tenant_uuid"], "body": oracle.emit_fm_scale_up(e["tenant_uuid"], mid, "gpu", MODEL)}


# ---- 2. Attaching entries decided on the node side (CM + DRA block :2198-3192) -------------------
def test_cm_dra_attaching_node_side_entries(cro, oracle):
    smi = {"stdout": DEV, "stderr": ""}
    cases = {
        2198: dict(driver_pod_missing=True),                                          # nvidia-driver pod missing: recorded, continues
        2311: dict(enumeration=smi, daemonsets={}),                                   # kubelet-plugin DaemonSet missing: recorded, continues
        2426: dict(enumeration=smi, daemonsets={DRA_DS: READY}),                      # restarted now
        2566: dict(enumeration=smi, daemonsets={DRA_DS: dict(READY, restarted_at=NOW)}),                      # restarted <= 10 s ago
        2709: dict(enumeration=smi, daemonsets={DRA_DS: dict(READY, ready=0, unavailable=1, restarted_at=NOW)}),   # not ready
        2863: dict(enumeration=smi, daemonsets={DRA_DS: dict(READY, restarted_at="error")}),
        3000: dict(enumeration=smi, daemonsets={DRA_DS: READY}, resource_slices=[{"devices": [{"attributes": {"uuid": DEV}}]}]),
    }
    for line, extra in cases.items():
        e = ENTRIES[line]
        out = cro.reconcile_attach(None, request_for(e, **extra))
        check(cro, oracle, e, out)
        assert out["requeue_after_s"] == (0 if line == 3000 else 30), line
        restarts = [r.split("@")[0] for r in out.get("daemonset_restarts", [])]
        assert restarts == ([DRA_DS] if line in (2426, 3000) else []), (line, restarts)
        if restarts:
            assert out["daemonset_restarts"][0].endswith("@" + NOW)                   # the stamp is time.Now().Format(RFC3339)


def test_attaching_deleted_entries(cro, oracle):
    """:3150 / :3166 / :3192 — deletionTimestamp set while Attaching (composableresource_controller.go:203-213)."""
    smi = {"stdout": DEV, "stderr": ""}
    for line in (3150, 3166, 3192):
        e = ENTRIES[line]
        out = cro.reconcile_attach(None, request_for(e, deleting=True, enumeration=smi, daemonsets={DRA_DS: READY},
                                                     resource_slices=[{"devices": [{"attributes": {"uuid": DEV}}]}]))
        check(cro, oracle, e, out)


# ---- 3. FM + DEVICE_PLUGIN Attaching, node side (:6682-7398) -------------------------------------
def test_fm_device_plugin_attaching_node_side_entries(cro, oracle):
    ok = {DP_DS: READY, DCGM_DS: READY}
    cases = {
        6682: dict(driver_pod_missing=True, daemonsets=ok),
        6747: dict(enumeration={"stdout": "", "stderr": ""}, daemonsets={DP_DS: READY}),              # nvidia-dcgm missing: recorded
        6873: dict(enumeration={"stdout": "", "stderr": ""}, daemonsets=ok),                          # not visible yet
        7050: dict(enumeration={"stdout": "", "stderr": "nvidia-smi: command not found"}, daemonsets=ok),
        7221: dict(enumeration={"stdout": DEV, "stderr": ""}, daemonsets=ok),
        7398: dict(enumeration={"stdout": DEV, "stderr": ""}, daemonsets=ok),                         # Warning in FM still goes Online
    }
    for line, extra in cases.items():
        e = ENTRIES[line]
        out = cro.reconcile_attach(None, request_for(e, **extra))
        if line == 7050:
            assert out["error"].startswith("get gpu info command failed: err: '<nil>', stderr: 'nvidia-smi: command not found'")
        check(cro, oracle, e, out)


# ---- 4. Detaching entries through the real clients ------------------------------------------------
def test_detaching_entries_through_the_clients(cro, oracle):
    gone = dict(enumeration_after_remove={"stdout": "", "stderr": ""}, resource_slices_after_remove=[])
    cases = {
        # CM + DRA
        4579: {}, 4756: {}, 4934: {},                                                 # GET / resize fail
        5113: {},                                                                     # resize accepted: wait (ErrWaitingDeviceDetaching)
        5296: {},                                                                     # REMOVE_FAILED: reason recorded, resize sent again
        5495: dict(daemonsets={}, **gone),                                            # device gone upstream; kubelet-plugin DaemonSet missing
        5694: dict(daemonsets={DRA_DS: READY}, **gone),
        # FM + DEVICE_PLUGIN
        8749: {}, 8895: {},
        9035: dict(daemonsets={}, **gone), 9187: dict(daemonsets={DP_DS: READY}, **gone),
        9366: dict(daemonsets={DP_DS: READY, DCGM_DS: READY}, **gone),
    }
    for line, extra in cases.items():
        e = ENTRIES[line]
        r = request_for(e, deleting=True, **extra)
        if line in (5495, 5694):      # the device is no longer in CM (routes of cluster ...0000 serve a machine without it)
            pass
        if line in (9035, 9187, 9366):    # FM: the resource is already gone -> RemoveResource returns nil without a DELETE
            r["fabric"]["objects"] = objects_for(e, bmh_uuid="machine0-uuid-temp-fail-000000000003")
        out = cro.reconcile_attach(None, r)
        check(cro, oracle, e, out)
        if line in (5113, 5296):
            assert out["requeue_after_s"] == 30
            assert out["fabric_requests"][-1]["method"] == "POST"
        if line == 5296:
            assert json.loads(out["fabric_requests"][-1]["body"]) == {"remove_resources": {
                "spec_uuid": "spec0000-uuid-temp-0000-000000000002", "device_count": -1, "devices": [FAILDEV]}}
        if line in (9035, 9187, 9366):
            assert [q["method"] for q in out["fabric_requests"]] == ["GET"]


def test_detaching_node_side_entries(cro, oracle):
    """Load check / drain failures happen before the provider is asked (:4131-4352, :8441-8569)."""
    cases = {
        4131: dict(load_check={"stdout": "", "stderr": "nvidia-smi: command not found"}),
        4194: dict(load_check={"stdout": DEV + ", gpu_load_progress", "stderr": ""}),
        4257: dict(drain={"fd_scan": {"stdout": "nvidia-persist", "stderr": ""}}),
        4352: dict(drain={"error": "no Pod named 'nvidia-dra-driver-gpu-kubelet-plugin' found on node worker-0"}),
        8441: dict(load_check={"stdout": "", "stderr": "nvidia-smi: command not found"}),
        8505: dict(load_check={"stdout": DEV + ", gpu_load_progress", "stderr": ""}),
        8569: dict(drain={"fd_scan": {"stdout": "nvidia-persist", "stderr": ""}}),
    }
    for line, extra in cases.items():
        e = ENTRIES[line]
        out = cro.reconcile_attach(None, request_for(e, deleting=True, **extra))
        check(cro, oracle, e, out)
        assert out.get("fabric_requests") == [], line                                 # RemoveResource was never reached


# ---- 5. state transitions that need no provider --------------------------------------------------
def test_transitions(cro, oracle):
    e = ENTRIES[1042]                                                                 # None -> Attaching
    check(cro, oracle, e, cro.reconcile_attach(None, request_for(e)))
    for line in (3946, 8327):                                                         # Online, deleted -> Detaching
        e = ENTRIES[line]
        check(cro, oracle, e, cro.reconcile_attach(None, request_for(e, deleting=True)))


def test_wrong_env_entries(cro):
    """:9702 / :9714 / :9726 / :9864 — NewComposableResourceAdapter refuses, the error lands in Status.Error."""
    base = {"DEVICE_RESOURCE_TYPE": "DRA", "CDI_PROVIDER_TYPE": "FTI_CDI", "FTI_CDI_API_TYPE": "CM",
            "FTI_CDI_TENANT_ID": "t", "FTI_CDI_CLUSTER_ID": "c"}
    for line, key, state in ((9702, "CDI_PROVIDER_TYPE", "Attaching"), (9714, "FTI_CDI_API_TYPE", "Attaching"),
                             (9726, "DEVICE_RESOURCE_TYPE", "Attaching"), (9864, "DEVICE_RESOURCE_TYPE", "Detaching")):
        e = ENTRIES[line]
        out = cro.reconcile_attach(None, {"status": {"state": state, "device_id": DEV if state == "Detaching" else ""},
                                          "env": dict(base, **{key: "ERROR"}), "fabric": {}})
        assert out["error"] == e["expected_error"] == out["status"]["error"], line
    out = cro.reconcile_attach(None, {"status": {"state": "Attaching"}, "fabric": {},
                                      "env": dict(base, DEVICE_RESOURCE_TYPE="DEVICE_PLUGIN", FTI_CDI_CLUSTER_ID="")})
    assert out["error"] == "The cluster in RKE2 does not support DEVICE_PLUGIN, please use DRA"      # composableresource_adapter.go:52-54


# ---- 6. garbage collection and the not-found object, through the in-memory API -------------------
def test_garbage_collection_entries(cro):
    """:1180 / :4454 / :6112 / :8654 — the target Node is gone: the CR is marked Deleting with
    "target node worker-0 not found" and deleted (composableresource_controller.go:128-174)."""
    for line, state in ((1180, "Attaching"), (6112, "Attaching"), (4454, "Detaching"), (8654, "Detaching")):
        e = ENTRIES[line]
        assert e["expected_deleted"] is True
        dtype = "DEVICE_PLUGIN" if "DEVICE_PLUGIN" in " ".join(e["context"]) else "DRA"
        with cro.Cluster({"nodes": [], "probe": False, "device_resource_type": dtype}) as c:
            init = e.get("initial_status") or {}
            c.plant({"kind": "ComposableResource", "name": "test-composable-resource", "finalizer": False,   # as the reference creates it
                     "spec": {"type": "gpu", "model": MODEL, "target_node": "worker-0"},
                     "status": {"state": state, "device_id": init.get("DeviceID", ""), "cdi_device_id": init.get("CDIDeviceID", "")}})
            assert c.reconcile_resource("test-composable-resource") == ""
            assert c.dump()["resources"] == {}, line


def test_none_state_update_failure(cro):
    """:1055 — AddFinalizer + Update fails ("update fails", the reference's MockUpdate hook): nothing is written."""
    with cro.Cluster({"nodes": ["worker-0"], "probe": False, "device_resource_type": "DRA"}) as c:
        c.plant({"kind": "ComposableResource", "name": "test-composable-resource", "finalizer": False,
                 "spec": {"type": "gpu", "model": MODEL, "target_node": "worker-0"}, "status": {"state": ""}})
        c.plant({"kind": "Fault", "update": "update fails"})
        assert c.reconcile_resource("test-composable-resource") == "update fails"
        r = c.dump()["resources"]["test-composable-resource"]
        assert r["finalizers"] == [] and r["status"]["state"] == ""
        c.plant({"kind": "Fault"})
        assert c.reconcile_resource("test-composable-resource") == ""
        r = c.dump()["resources"]["test-composable-resource"]
        assert r["finalizers"] == ["com.ie.ibm.hpsys/finalizer"] and r["status"]["state"] == "Attaching"    # :1042


def test_missing_object_and_direct_delete(cro):
    with cro.Cluster({"nodes": ["worker-0"], "probe": False, "device_resource_type": "DRA"}) as c:
        assert c.reconcile_resource("unexisted-composable-resource") == ""            # :986 NotFound: stop, no requeue
        c.plant({"kind": "ComposableResource", "name": "r", "finalizer": True, "deleting": True,
                 "spec": {"type": "gpu", "model": MODEL, "target_node": "worker-0"}, "status": {"state": "Deleting"}})
        assert c.reconcile_resource("r") == ""                                        # :5983 Deleting: finalizer off, object gone
        assert c.dump()["resources"] == {}


# ---- 7. UpstreamSyncer's read of the fabric ------------------------------------------------------
def test_syncer_get_resources_entries(cro):
    """SYNC:243 / SYNC:269 — syncUpstreamData wraps the client's error (upstreamsyncer_controller.go:80-84)."""
    env = {"DEVICE_RESOURCE_TYPE": "DRA", "CDI_PROVIDER_TYPE": "FTI_CDI", "FTI_CDI_API_TYPE": "CM",
           "FTI_CDI_TENANT_ID": "tenant00-uuid-temp-0000-000000000000", "FTI_CDI_CLUSTER_ID": "cluster0-uuid-temp-0000-000000000000"}
    node = {"worker-0": {"annotations": {"machine.openshift.io/machine": "openshift-machine-api/machine-worker-0"}}}
    out = cro.fabric_list_devices({"env": env, "fabric": {"http": HTTP, "objects": {"nodes": node}}})
    assert "failed to fetch data from upstream server: " + out["error"] == \
        "failed to fetch data from upstream server: metal3machines.infrastructure.cluster.x-k8s.io \"machine-worker-0\" not found"
    full = objects_for(ENTRIES[1739])
    out = cro.fabric_list_devices({"env": dict(env, FTI_CDI_CLUSTER_ID="cluster0-uuid-temp-fail-000000000000"),
                                   "fabric": {"http": HTTP, "objects": full}})
    assert out["error"] == "failed to process CM get request. http returned status: '404', cm return code: 'E02XXXX', error message: 'machine not found'"
    assert out["devices"] == []
    # the healthy cluster: one device, CM flavour leaves Model empty (cm/client.go:335-341)
    out = cro.fabric_list_devices({"env": dict(env, FTI_CDI_CLUSTER_ID="cluster0-uuid-temp-0000-000000000001"),
                                   "fabric": {"http": HTTP, "objects": full}})
    assert out["error"] == "" and out["devices"] == [{
        "node_name": "worker-0", "machine_uuid": "machine0-uuid-temp-0000-000000000000", "device_type": "gpu", "model": "",
        "device_id": DEV, "cdi_device_id": RES}]
    # FM flavour: a node that fails is skipped, the others are still listed (fm/client.go:373-383)
    two = json.loads(json.dumps(full))
    two["nodes"]["worker-1"] = {"annotations": {}}
    out = cro.fabric_list_devices({"env": dict(env, FTI_CDI_API_TYPE="FM"), "fabric": {"http": HTTP, "objects": two}})
    assert out["error"] == "" and [d["device_id"] for d in out["devices"]] == ["", DEV]


# ---- 8. C++ clients vs the Python restatement on mutated fabrics ----------------------------------
def test_clients_fuzz_vs_oracle(cro, oracle):
    import fabric_clients as fc
    rng = random.Random(20260921)
    bodies = [r["body"] for r in HTTP if r["body"]] + ["", "null", "[]", "7", '"x"', '{"status":"404","detail":[]}',
                                                      '{"status":404.5,"detail":{"code":7,"message":{"k": [1, 2]}}}',
                                                      '{"status":404,"detail":{"code":"E","message": [1,  2] }}', '{"detail":null}']

    def mutate(s):
        if not s or rng.random() < 0.3:
            return s
        i = rng.randrange(len(s))
        return rng.choice([s[:i], s[:i] + rng.choice('{}[]",:x1 \n\\') + s[i:], s[:i] + s[i + 1:], s + rng.choice(["x", " ", "}", "\n"])])

    import collections
    tally = collections.Counter()
    for it in range(600):
        kind = rng.choice(["cm", "fm"])
        tenant = rng.choice(["tenant-0", "tenant-1", "tenant 2/ü&x=y"])       # the last one exercises url.Values.Encode
        cluster = rng.choice(["", "cluster-a"]) if kind == "fm" else "cluster-a"
        def mostly(good, *bad):      # the metal3 walk usually succeeds so the HTTP legs get exercised
            return good if rng.random() < 0.93 else rng.choice(bad)
        objs = {"nodes": {"worker-0": {"annotations": {"machine.openshift.io/machine": mostly("ns/m0", "m0", "a/b/c", "")},
                                       "provider_id": mostly("fsas-cdi://mach-9", "aws://x", "")}},
                "metal3machines": {"ns/m0": {"annotations": {"metal3.io/BareMetalHost": mostly("ns/b0", "b0", "")}}},
                "baremetalhosts": {"ns/b0": mostly({"annotations": {"cluster-manager.cdi.io/machine": "mach-1"}}, {"annotations": {}}, {})},
                "composable_resource_device_ids": rng.choice([[], [DEV], [FAILDEV]])}
        if rng.random() < 0.05:
            objs.pop(rng.choice(["nodes", "metal3machines", "baremetalhosts"]))
        # replies that fit the flavour and the verb most of the time, anything at all otherwise
        fitting = {("cm", "GET"): [cm_machine_data(None), cm_machine_data([(DEV, "ADD_COMPLETE", "", RES, rng.choice("0123"))]),
                                   cm_machine_data([(FAILDEV, rng.choice(["ADD_FAILED", "REMOVE_FAILED"]), "some reason", FAILRES, "0")])],
                   ("fm", "GET"): [fm_machine_data([]), fm_machine_data([(RES, "gpu", rng.choice("0123"), DEV, MODEL)]),
                                   fm_machine_data([(FAILRES, "gpu", "0", FAILDEV, MODEL)])],
                   ("fm", "PATCH"): [v for k, v in KATS["fixtures"]["fm_update_body"].items() if k != "_cite"]}
        http = []
        for m in ("GET", "POST", "PATCH", "DELETE"):
            if rng.random() < 0.75:
                http.append({"method": m, "path_contains": "", "status": 200, "body": rng.choice(fitting.get((kind, m), [""]))})
            else:
                http.append({"method": m, "path_contains": "", "status": rng.choice([200, 204, 404, 500]), "body": mutate(rng.choice(bodies))})
        fabric = {"http": http, "objects": objs, "token_error": rng.choice(["", "", "", "unable to rotate token: boom"])}
        env = {"DEVICE_RESOURCE_TYPE": "DRA", "CDI_PROVIDER_TYPE": "FTI_CDI", "FTI_CDI_API_TYPE": kind.upper(),
               "FTI_CDI_TENANT_ID": tenant, "FTI_CDI_CLUSTER_ID": cluster}
        state = rng.choice(["Attaching", "Online", "Detaching"])
        dev, res = rng.choice([(DEV, RES), (FAILDEV, FAILRES)])
        req = {"name": "cr-x", "spec": {"type": "gpu", "model": MODEL, "target_node": "worker-0"},
               "status": {"state": state, "device_id": "" if state == "Attaching" else dev, "cdi_device_id": "" if state == "Attaching" else res},
               "deleting": state == "Detaching", "probe": False, "env": env, "fabric": fabric,
               "enumeration": {"stdout": "", "stderr": ""}, "resource_slices": [],
               "enumeration_after_remove": {"stdout": "", "stderr": ""}, "resource_slices_after_remove": []}
        out = cro.reconcile_attach(None, req)
        f = fc.Fabric(fabric)
        client = (fc.CMClient if kind == "cm" else fc.FMClient)(f, tenant, cluster)

        def is_panic(e):     # a Go run-time panic in the client (res_op_status[:1] on "", machines[0] of an empty list):
            return e.startswith("runtime error: ")   # it unwinds past requeueOnErr — no status write, "panic: ... [recovered]"
        if state == "Attaching":
            d, c, err = client.add("cr-x", "gpu", MODEL, "worker-0")
            if err == fc.ERR_ATTACHING:
                want = ("", "Attaching", "")
            elif is_panic(err):
                want = ("panic: %s [recovered]" % err, "Attaching", "")
            elif err:                # also when an ADD_FAILED device came back with ids: the error wins (:217-229)
                want = (err, "Attaching", err)
            else:
                want = None
            tally[("add", "wait" if err == fc.ERR_ATTACHING else "err" if err else "ids")] += 1
            if want:
                assert (out["error"], out["status"]["state"], out["status"].get("error", "")) == want, (it, req, out)
            else:                    # ids recorded; nothing is visible yet -> stays Attaching, requeue 30 s
                assert (out["error"], out["status"].get("device_id", ""), out["status"].get("cdi_device_id", "")) == ("", d, c), (it, out)
                assert out["requeue_after_s"] == 30
        elif state == "Online":
            err = client.check("gpu", MODEL, "worker-0", dev)
            tally[("check", err.split(":")[0][:32])] += 1
            if is_panic(err):        # CheckResource's error is recorded, never returned — a panic is not an error
                assert out["error"] == "panic: %s [recovered]" % err and out["status"].get("error", "") == "" and out["status_updates"] == [], (it, out)
            else:
                assert out["error"] == "" and out["status"].get("error", "") == err, (it, fabric, out)
        else:
            if kind == "cm":
                err, recorded = client.remove("gpu", MODEL, "worker-0", dev)
            else:
                err, recorded = client.remove("gpu", "worker-0", res), None
            tally[("remove", "wait" if err == fc.ERR_DETACHING else "err" if err else "gone")] += 1
            if err == fc.ERR_DETACHING:
                assert out["error"] == "" and out["requeue_after_s"] == 30, (it, out)
                if recorded is not None:
                    assert out["status"].get("error", "") == recorded
            elif err:
                assert out["error"] == ("panic: %s [recovered]" % err if is_panic(err) else err), (it, fabric, out["error"], err)
            else:
                assert out["error"] == "" and out["status"]["state"] == "Deleting", (it, out)
        assert out["fabric_requests"] == f.requests, (it, out["fabric_requests"], f.requests)
    # the generator reaches every branch family, not just the early exits
    for key in (("add", "wait"), ("add", "err"), ("add", "ids"), ("remove", "wait"), ("remove", "err"), ("remove", "gone")):
        assert tally[key] >= 3, tally


def test_restart_rule_and_time_parse_fuzz_vs_oracle(cro):
    import fabric_clients as fc
    rng = random.Random(7)
    now, _ = fc.parse_rfc3339(NOW)
    seeds = [NOW, "2025-06-01T11:59:55Z", "2025-06-01T11:59:49Z", "2025-06-01T21:00:00+09:00", "2024-02-29T23:59:60Z", "error", "",
             "2025-06-01T12:00:00.25Z", "2025-6-01T12:00:00Z", "2025-06-01T9:00:00Z", "2025-06-01 12:00:00Z", "2025-06-31T00:00:00Z",
             "2025-06-01T12:00:00+24:61", "2025-06-01T12:00:00-00:00", "2025-06-01T12:00:00,5Z", "٢٠٢٥-06-01T12:00:00Z", "2025-06-01T12:00:00Z\"\\"]
    for it in range(600):
        s = rng.choice(seeds)
        if it >= len(seeds) and rng.random() < 0.7:
            i = rng.randrange(len(s) + 1)
            s = rng.choice([s[:i] + rng.choice("0123456789-:TZ+. x") + s[i:], s[:i] + s[i + 1:], s[:i]])
        else:
            s = seeds[it % len(seeds)]
        ds = dict(READY, restarted_at=s)
        if rng.random() < 0.2:
            ds[rng.choice(["ready", "current"])] = 0
        restart, err = fc.restart_daemonset("nvidia-dra-driver-gpu", "nvidia-dra-driver-gpu-kubelet-plugin", ds, now)
        out = cro.reconcile_attach(None, {"status": {"state": "Attaching", "device_id": DEV, "cdi_device_id": RES}, "device_resource_type": "DRA",
                                          "spec": {"type": "gpu", "model": MODEL, "target_node": "worker-0"}, "probe": False, "provider": {},
                                          "enumeration": {"stdout": DEV, "stderr": ""}, "resource_slices": [], "daemonsets": {DRA_DS: ds}, "now": NOW})
        assert out["status"].get("error", "") == err, (s, out["status"], err)
        assert bool(out.get("daemonset_restarts")) == restart, (s, ds, out)


# ---- 9. the third provider: Sunfish (internal/cdi/sunfish/client.go; the reference has no test for it) ---
def test_sunfish_client_vs_oracle(cro, oracle):
    import fabric_clients as fc
    env = {"DEVICE_RESOURCE_TYPE": "DEVICE_PLUGIN", "CDI_PROVIDER_TYPE": "SUNFISH"}
    for model in ("NVIDIA-A100-PCIE-40GB", "Tesla-V100-PCIE-16GB", MODEL, "x\"<y"):
        for status in (200, 204, 500):
            fabric = {"http": [{"method": "PATCH", "path": "redfish/v1/Systems/System", "status": status, "body": ""}]}
            for state in ("Attaching", "Detaching"):
                req = {"name": "cr", "spec": {"type": "gpu", "model": model, "target_node": "worker-0"},
                       "status": {"state": state, "device_id": DEV if state == "Detaching" else "", "cdi_device_id": ""},
                       "deleting": state == "Detaching", "probe": False, "env": env, "fabric": fabric,
                       "enumeration": {"stdout": "", "stderr": ""}, "enumeration_after_remove": {"stdout": "", "stderr": ""}}
                out = cro.reconcile_attach(None, req)
                f = fc.Fabric(fabric)
                c = fc.SunfishClient(f)
                err = c.add("worker-0", model)[2] if state == "Attaching" else c.remove("worker-0", model)
                assert out["error"] == err == ("" if status != 500 else "http returned code 500")
                assert out["fabric_requests"] == f.requests
                body = json.loads(out["fabric_requests"][0]["body"])
                known = model in fc.SunfishClient.MODELS
                assert body == {"Name": "worker-0", "Processors": {"Members": [{
                    "@Redfish.RequestCount": 1 if (known and state == "Attaching") else 0,
                    "ProcessorType": "GPU" if known else "", "Model": model if known else ""}]}}
                if state == "Attaching" and not err:       # no ids come back (the reference's TODO): AddResource is asked again
                    assert out["status"].get("device_id", "") == ""
