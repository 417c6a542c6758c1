"""The remaining ComposabilityRequest states (None, Updating, Running, Cleaning, Deleting, child-status sync,
invalid state) on the reference's table entries (internal/controller/composabilityrequest_controller_test.go:
:527, :574, :604, :682, :1448-1600, :1635-1677, :1716-1742, :1779)."""
CR0, CR1 = "gpu-00000000-temp-uuid-0000-000000000000", "gpu-00000000-temp-uuid-0000-000000000001"
BASE = {"type": "gpu", "model": "NVIDIA-A100-PCIE-80GB", "size": 2, "allocation_policy": "samenode"}
FIN = "com.ie.ibm.hpsys/finalizer"


def plant_request(c, state, spec=None, res_states=("", ""), scalar=None, **kw):
    resources = {CR0: {"node_name": "worker-0", "state": res_states[0]}, CR1: {"node_name": "worker-0", "state": res_states[1]}}
    c.plant(dict({"kind": "ComposabilityRequest", "name": "test-composability-request", "resource": spec or BASE,
                  "status": {"state": state, "resources": resources, "scalarResource": scalar or BASE}}, **kw))


def plant_children(c, names=(CR0, CR1), owner="test-composability-request", state=""):
    for n in names:
        c.plant({"kind": "ComposableResource", "name": n, "labels": {"app.kubernetes.io/managed-by": owner},
                 "spec": {"type": BASE["type"], "model": BASE["model"], "target_node": "worker-0"}, "status": {"state": state}})


def status(c):
    return c.dump()["requests"]["test-composability-request"]["status"]


def test_invalid_state(cro):                                       # :527
    with cro.Cluster({"nodes": ["worker-0"]}) as c:
        plant_request(c, "unknown")
        assert c.reconcile_request("test-composability-request") == "the composabilityRequest state 'unknown' is invalid"
        assert status(c)["error"] == "the composabilityRequest state 'unknown' is invalid"


def test_child_status_is_mirrored_up(cro):                         # :574
    with cro.Cluster({"nodes": ["worker-0"]}) as c:
        plant_request(c, "Running", res_states=("Updating", ""))
        plant_children(c, names=(CR0,))
        assert c.reconcile_request(CR0) == ""                       # the request controller is keyed by the CHILD's name
        st = status(c)
        assert st["state"] == "Running" and st["resources"][CR0] == {"state": "", "node_name": "worker-0"}


def test_child_of_a_missing_request(cro):                          # :604
    with cro.Cluster({"nodes": ["worker-0"]}) as c:
        plant_children(c, names=(CR0,), owner="unknown-composability-request")
        assert c.reconcile_request(CR0) == 'composabilityrequests.cro.hpsys.ibm.ie.com "unknown-composability-request" not found'


def test_none_state(cro):                                          # :682
    with cro.Cluster({"nodes": ["worker-0"]}) as c:
        assert c.apply("test-composability-request", BASE) == ""
        assert c.reconcile_request("test-composability-request") == ""
        r = c.dump()["requests"]["test-composability-request"]
        assert r["finalizers"] == [FIN] and r["status"]["state"] == "NodeAllocating" and "resources" not in r["status"]
        assert r["status"]["scalarResource"] == BASE


def test_updating_all_online_goes_running(cro):                    # :1448
    with cro.Cluster({"nodes": ["worker-0"]}) as c:
        plant_request(c, "Updating", res_states=("Online", "Online"))
        plant_children(c)
        assert c.reconcile_request("test-composability-request") == ""
        st = status(c)
        assert st["state"] == "Running" and all(r["state"] == "Online" for r in st["resources"].values())
        assert set(c.dump()["resources"]) == {CR0, CR1}


def test_updating_creates_missing_children_and_waits(cro):         # :1499
    with cro.Cluster({"nodes": ["worker-0"]}) as c:
        plant_request(c, "Updating")
        assert c.reconcile_request("test-composability-request") == ""
        assert status(c)["state"] == "Updating"
        kids = c.dump()["resources"]
        assert set(kids) == {CR0, CR1}
        for k in kids.values():
            assert k["spec"] == {"type": "gpu", "model": BASE["model"], "target_node": "worker-0"}
            assert k["status"] == {"state": ""} and k["labels"] == {"app.kubernetes.io/managed-by": "test-composability-request"}


def test_updating_deletes_surplus_children(cro):                   # :1528
    with cro.Cluster({"nodes": ["worker-0"]}) as c:
        plant_request(c, "Updating")
        plant_children(c, names=(CR0, CR1, "gpu-00000000-temp-uuid-0000-000000000002"))
        assert c.reconcile_request("test-composability-request") == ""
        d = c.dump()
        assert d["requests"]["test-composability-request"]["status"]["state"] == "Updating"
        assert d["resources"]["gpu-00000000-temp-uuid-0000-000000000002"]["deleting"] is True   # finalizer keeps it until it detaches
        assert not d["resources"][CR0]["deleting"] and not d["resources"][CR1]["deleting"]


def test_updating_and_running_notice_spec_changes_and_deletion(cro):   # :1570, :1586, :1646, :1662
    for state in ("Updating", "Running"):
        with cro.Cluster({"nodes": ["worker-0"]}) as c:
            plant_request(c, state, spec=dict(BASE, size=3), res_states=("Online", "Online"))
            plant_children(c, state="Online")
            assert c.reconcile_request("test-composability-request") == ""
            st = status(c)
            assert st["state"] == "NodeAllocating" and st["scalarResource"]["size"] == 3
        with cro.Cluster({"nodes": ["worker-0"]}) as c:
            plant_request(c, state, res_states=("Online", "Online"), deleting=True)
            plant_children(c, state="Online")
            assert c.reconcile_request("test-composability-request") == ""
            assert status(c)["state"] == "Cleaning"


def test_running_stays_running(cro):                               # :1635
    with cro.Cluster({"nodes": ["worker-0"]}) as c:
        plant_request(c, "Running", res_states=("Online", "Online"))
        plant_children(c, state="Online")
        assert c.reconcile_request("test-composability-request") == ""
        assert status(c)["state"] == "Running" and "error" not in status(c)


def test_cleaning(cro):                                            # :1716, :1729
    with cro.Cluster({"nodes": ["worker-0"]}) as c:
        plant_request(c, "Cleaning", deleting=True)
        plant_children(c)
        assert c.reconcile_request("test-composability-request") == ""
        d = c.dump()
        assert d["requests"]["test-composability-request"]["status"]["state"] == "Cleaning"
        assert all(r["deleting"] for r in d["resources"].values())
    with cro.Cluster({"nodes": ["worker-0"]}) as c:
        plant_request(c, "Cleaning", deleting=True)
        assert c.reconcile_request("test-composability-request") == ""
        assert status(c)["state"] == "Deleting"


def test_deleting_drops_the_finalizer_and_the_object(cro):         # :1779
    with cro.Cluster({"nodes": ["worker-0"]}) as c:
        plant_request(c, "Deleting", deleting=True)
        assert c.reconcile_request("test-composability-request") == ""
        assert c.dump()["requests"] == {}


# ---- CRD validation: the API server's answer to a bad spec (:324-:412) --------------------------------
INVALID = [
    (":324", {"type": "UnknownType"}, 'spec.resource.type: Unsupported value: "UnknownType": supported values: "gpu", "cxlmemory"'),
    (":334", {"size": -1}, "spec.resource.size: Invalid value: -1: spec.resource.size in body should be greater than or equal to 0"),
    (":344", {"allocation_policy": "UnknownPolicy"},
     'spec.resource.allocation_policy: Unsupported value: "UnknownPolicy": supported values: "samenode", "differentnode"'),
    (":354", {"other_spec": {"milli_cpu": -1, "memory": 1, "ephemeral_storage": 1, "allowed_pod_number": 1}},
     "spec.resource.other_spec.milli_cpu: Invalid value: -1: spec.resource.other_spec.milli_cpu in body should be greater than or equal to 0"),
    (":369", {"other_spec": {"milli_cpu": 1, "memory": -1, "ephemeral_storage": 1, "allowed_pod_number": 1}},
     "spec.resource.other_spec.memory: Invalid value: -1: spec.resource.other_spec.memory in body should be greater than or equal to 0"),
    (":384", {"other_spec": {"milli_cpu": 1, "memory": 1, "ephemeral_storage": -1, "allowed_pod_number": 1}},
     "spec.resource.other_spec.ephemeral_storage: Invalid value: -1: spec.resource.other_spec.ephemeral_storage in body should be greater than or equal to 0"),
    (":399", {"other_spec": {"milli_cpu": 1, "memory": 1, "ephemeral_storage": 1, "allowed_pod_number": -1}},
     "spec.resource.other_spec.allowed_pod_number: Invalid value: -1: spec.resource.other_spec.allowed_pod_number in body should be greater than or equal to 0"),
]


def test_invalid_specs_are_refused_with_the_api_servers_words(cro):
    with cro.Cluster({"nodes": ["worker-0"]}) as c:
        for cite, over, detail in INVALID:
            got = c.apply("test-composability-request", dict(BASE, **over))
            assert got == 'ComposabilityRequest.cro.hpsys.ibm.ie.com "test-composability-request" is invalid: ' + detail, cite
        assert c.dump()["requests"] == {}


# ---- API-server write failures (the reference's MockStatusUpdate / MockUpdate hooks, :452-:513, :694, :1784) ----
def test_unknown_request_is_not_requeued(cro):                     # :452 "should wait when the request is invalid"
    with cro.Cluster({"nodes": ["worker-0"]}) as c:
        assert c.reconcile_request("unknown-request") == ""


def test_status_update_failures_surface_from_every_state(cro):
    for cite, state in ((":457", "NodeAllocating"), (":471", "Updating"), (":485", "Running"), (":499", "Cleaning")):
        with cro.Cluster({"nodes": ["worker-0"]}) as c:
            plant_request(c, state, res_states=("Online", "Online"), finalizer=True)
            assert c.delete("test-composability-request")           # "Use Delete() to trigger r.Status().Update" (:431)
            c.plant({"kind": "Fault", "status_update": "status update fails"})
            assert c.reconcile_request("test-composability-request") == "status update fails", cite
            assert status(c)["state"] == state                      # nothing was written
            c.plant({"kind": "Fault"})                              # cleared: the same reconcile now goes through
            assert c.reconcile_request("test-composability-request") == ""
            assert status(c)["state"] in ("Cleaning", "Deleting")


def test_update_failures_when_the_finalizer_changes(cro):
    with cro.Cluster({"nodes": ["worker-0"]}) as c:                # :513 Deleting: RemoveFinalizer + Update
        plant_request(c, "Deleting", finalizer=True)
        c.delete("test-composability-request")
        c.plant({"kind": "Fault", "update": "update fails"})
        assert c.reconcile_request("test-composability-request") == "update fails"
        assert c.dump()["requests"]["test-composability-request"]["finalizers"] == [FIN]
    with cro.Cluster({"nodes": ["worker-0"]}) as c:                # :694 None: AddFinalizer + Update
        assert c.apply("test-composability-request", BASE) == ""
        c.plant({"kind": "Fault", "update": "update fails"})
        assert c.reconcile_request("test-composability-request") == "update fails"
        r = c.dump()["requests"]["test-composability-request"]
        assert r["finalizers"] == [] and r["status"]["state"] == ""
    with cro.Cluster({"nodes": ["worker-0"]}) as c:                # :1784 Deleting (its own table)
        plant_request(c, "Deleting", finalizer=True)
        c.delete("test-composability-request")
        c.plant({"kind": "Fault", "update": "update fails"})
        assert c.reconcile_request("test-composability-request") == "update fails"
