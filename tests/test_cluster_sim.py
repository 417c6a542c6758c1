"""The caller of the hot path: both reconcilers over the in-memory API (cro_sim_*).

KATs are the NodeAllocating entries of the reference's own table
(internal/controller/composabilityrequest_controller_test.go:729-818 node matrix,
:889-1018 fresh-request entries): same nodes, same specs, same expected node sets /
error strings.  Then the storm (BASELINE config 4) and churn (config 5) shapes on CPU
with the probe off (the GPU versions are in test_gpu_parity.py / bench)."""
import random

import pytest

GI = 1 << 30
NODES = [  # composabilityrequest_controller_test.go:729-818
    {"name": "worker-0", "cpu": 8, "memory": 16 * GI, "ephemeral_storage": 512 * GI, "pods": 100},
    {"name": "worker-1", "cpu": 2, "memory": 16 * GI, "ephemeral_storage": 512 * GI, "pods": 100},
    {"name": "worker-2", "cpu": 8, "memory": 4 * GI, "ephemeral_storage": 512 * GI, "pods": 100},
    {"name": "worker-3", "cpu": 8, "memory": 16 * GI, "ephemeral_storage": 128 * GI, "pods": 100},
    {"name": "worker-4", "cpu": 8, "memory": 16 * GI, "ephemeral_storage": 512 * GI, "pods": 20},
    {"name": "worker-5", "cpu": 8, "memory": 16 * GI, "ephemeral_storage": 512 * GI, "pods": 100},
    {"name": "worker-6", "cpu": 32, "memory": 64 * GI, "ephemeral_storage": 2048 * GI, "pods": 100},
    {"name": "worker-7", "cpu": 32, "memory": 64 * GI, "ephemeral_storage": 2048 * GI, "pods": 100},
]
BASE = {"type": "gpu", "model": "NVIDIA-A100-PCIE-80GB", "size": 2, "allocation_policy": "samenode"}   # :161-172
BIG = {"milli_cpu": 32, "memory": 64 * GI, "ephemeral_storage": 2048 * GI, "allowed_pod_number": 100}
FIT0 = {"milli_cpu": 8, "memory": 16 * GI, "ephemeral_storage": 512 * GI, "allowed_pod_number": 100}
TOO = {"milli_cpu": 64, "memory": 16 * GI, "ephemeral_storage": 512 * GI, "allowed_pod_number": 100}

ALLOC_KATS = [  # (cite, spec overrides, expected used nodes | error)
    (":889 samenode, no TargetNode, no OtherSpec", {}, ["worker-0", "worker-0"]),
    (":894 samenode, no TargetNode, satisfiable OtherSpec", {"other_spec": BIG}, ["worker-6", "worker-6"]),
    (":909 samenode, no TargetNode, unsatisfiable OtherSpec",
     {"other_spec": {"milli_cpu": 64, "memory": 8 * GI, "ephemeral_storage": 256 * GI, "allowed_pod_number": 50}},
     "insufficient number of available nodes"),
    (":924 samenode, existed TargetNode", {"target_node": "worker-0"}, ["worker-0", "worker-0"]),
    (":934 samenode, existed TargetNode, satisfiable OtherSpec", {"target_node": "worker-0", "other_spec": FIT0}, ["worker-0", "worker-0"]),
    (":950 samenode, existed TargetNode, unsatisfiable OtherSpec", {"target_node": "worker-0", "other_spec": TOO},
     "TargetNode does not meet spec's requirements"),
    (":976 differentnode, no OtherSpec", {"allocation_policy": "differentnode"}, ["worker-0", "worker-1"]),
    (":986 differentnode, satisfiable OtherSpec", {"allocation_policy": "differentnode", "other_spec": FIT0}, ["worker-0", "worker-5"]),
    (":1002 differentnode, unsatisfiable OtherSpec", {"allocation_policy": "differentnode", "other_spec": TOO},
     "insufficient number of available nodes"),
]


@pytest.mark.parametrize("cite,over,expect", ALLOC_KATS, ids=[k[0] for k in ALLOC_KATS])
def test_node_allocating_kats(cro, cite, over, expect):
    with cro.Cluster({"nodes": NODES}) as c:
        spec = dict(BASE, **over)
        assert c.plant({"kind": "ComposabilityRequest", "name": "test-composability-request", "resource": spec,
                        "status": {"state": "NodeAllocating"}}) == ""
        err = c.reconcile_request("test-composability-request")
        st = c.dump()["requests"]["test-composability-request"]["status"]
        if isinstance(expect, str):
            assert err == expect and st["error"] == expect and st["state"] == "NodeAllocating"
        else:
            assert err == "" and st["state"] == "Updating"
            assert sorted(r["node_name"] for r in st["resources"].values()) == sorted(expect)
            assert all(n.startswith("gpu-") and len(n) == 40 for n in st["resources"])


def test_unknown_target_node_is_garbage_collected(cro):
    """:966 'samenode, unexisted TargetNode' -> the request is deleted (expectedRequestDeleted)."""
    with cro.Cluster({"nodes": NODES}) as c:
        c.plant({"kind": "ComposabilityRequest", "name": "r", "resource": dict(BASE, target_node="worker-unknown"),
                 "status": {"state": "NodeAllocating"}})
        assert c.reconcile_request("r") == ""
        assert c.dump()["requests"]["r"]["deleting"] is True


def test_other_request_occupies_node(cro):
    """:1374 'should succeed when there are ComposabilityRequests existed': worker-0 and worker-1 are taken."""
    with cro.Cluster({"nodes": NODES}) as c:
        c.plant({"kind": "ComposabilityRequest", "name": "other-0", "resource": dict(BASE, target_node="worker-0"),
                 "status": {"state": "Running"}})
        c.plant({"kind": "ComposabilityRequest", "name": "other-1", "resource": dict(BASE, target_node="worker-1"),
                 "status": {"state": "Running"}})
        c.plant({"kind": "ComposabilityRequest", "name": "r", "resource": BASE, "status": {"state": "NodeAllocating"}})
        assert c.reconcile_request("r") == ""
        st = c.dump()["requests"]["r"]["status"]
        assert sorted(x["node_name"] for x in st["resources"].values()) == ["worker-2", "worker-2"]


def test_size_shrink_evicts_by_priority(cro):
    """:1209 'changes the size when there are extra ComposableResource CRs': the Online child survives,
    the Attaching-without-device one goes first (bucket 0 of :326-338)."""
    with cro.Cluster({"nodes": NODES}) as c:
        res = {"gpu-a": {"node_name": "worker-0", "state": "Online"}, "gpu-b": {"node_name": "worker-0", "state": "Attaching"}}
        c.plant({"kind": "ComposabilityRequest", "name": "r", "resource": dict(BASE, size=1),
                 "status": {"state": "NodeAllocating", "resources": res}})
        for name, state, dev in (("gpu-a", "Online", "GPU-x"), ("gpu-b", "Attaching", "")):
            c.plant({"kind": "ComposableResource", "name": name, "labels": {"app.kubernetes.io/managed-by": "r"},
                     "spec": {"type": "gpu", "model": BASE["model"], "target_node": "worker-0"},
                     "status": {"state": state, "device_id": dev}})
        assert c.reconcile_request("r") == ""
        st = c.dump()["requests"]["r"]["status"]
        assert st["state"] == "Updating" and list(st["resources"]) == ["gpu-a"]


CR0, CR1 = "gpu-00000000-temp-uuid-0000-000000000000", "gpu-00000000-temp-uuid-0000-000000000001"
SPEC_CHANGE_KATS = [  # (cite, base policy, spec overrides, children must be dropped?, expected used nodes)
    (":1019 type, samenode", "samenode", {"type": "cxlmemory"}, True, ["worker-0", "worker-0"]),
    (":1033 model, samenode", "samenode", {"model": "NVIDIA-H100-PCIE-80GB"}, True, ["worker-0", "worker-0"]),
    (":1047 size 3, samenode", "samenode", {"size": 3}, False, ["worker-0", "worker-0", "worker-0"]),
    (":1060 ForceDetach, samenode", "samenode", {"force_detach": True}, True, ["worker-0", "worker-0"]),
    (":1074 policy -> differentnode", "samenode", {"allocation_policy": "differentnode"}, False, ["worker-0", "worker-1"]),
    (":1087 TargetNode -> worker-1", "samenode", {"target_node": "worker-1"}, True, ["worker-1", "worker-1"]),
    (":1101 OtherSpec, samenode", "samenode", {"other_spec": BIG}, True, ["worker-6", "worker-6"]),
    (":1121 type, differentnode", "differentnode", {"type": "cxlmemory"}, True, ["worker-0", "worker-1"]),
    (":1135 model, differentnode", "differentnode", {"model": "NVIDIA-H100-PCIE-80GB"}, True, ["worker-0", "worker-1"]),
    (":1149 size 3, differentnode", "differentnode", {"size": 3}, False, ["worker-0", "worker-1", "worker-2"]),
    (":1162 ForceDetach, differentnode", "differentnode", {"force_detach": True}, True, ["worker-0", "worker-1"]),
    (":1176 policy -> samenode", "differentnode", {"allocation_policy": "samenode"}, False, ["worker-0", "worker-0"]),
    (":1189 OtherSpec, differentnode", "differentnode", {"other_spec": BIG}, True, ["worker-6", "worker-7"]),
]


@pytest.mark.parametrize("cite,policy,over,dropped,expect", SPEC_CHANGE_KATS, ids=[k[0] for k in SPEC_CHANGE_KATS])
def test_spec_change_kats(cro, cite, policy, over, dropped, expect):
    """'should succeed when user changes the ...' (composabilityrequest_controller_test.go:1019-1205): two children
    exist from the old spec; after one NodeAllocating reconcile the listed nodes are used and, where the change
    invalidates them, neither old child survives in Status.Resources."""
    with cro.Cluster({"nodes": NODES}) as c:
        old = dict(BASE, allocation_policy=policy)
        nodes = ["worker-0", "worker-0"] if policy == "samenode" else ["worker-0", "worker-1"]
        c.plant({"kind": "ComposabilityRequest", "name": "r", "resource": dict(old, **over),
                 "status": {"state": "NodeAllocating", "scalarResource": old,
                            "resources": {CR0: {"node_name": nodes[0]}, CR1: {"node_name": nodes[1]}}}})
        for name, node in zip((CR0, CR1), nodes):
            c.plant({"kind": "ComposableResource", "name": name, "labels": {"app.kubernetes.io/managed-by": "r"},
                     "spec": {"type": old["type"], "model": old["model"], "target_node": node}, "status": {"state": ""}})
        assert c.reconcile_request("r") == ""
        st = c.dump()["requests"]["r"]["status"]
        assert st["state"] == "Updating"
        assert sorted(x["node_name"] for x in st["resources"].values()) == sorted(expect)
        if dropped:
            assert CR0 not in st["resources"] and CR1 not in st["resources"]


def test_admission_rules(cro):
    with cro.Cluster({"nodes": NODES}) as c:
        assert c.apply("a", dict(BASE, allocation_policy="differentnode", target_node="worker-0")) == \
            'admission webhook "vcomposabilityrequest.kb.io" denied the request: ' \
            "TargetNode cannot be specified when AllocationPolicy is set to 'differentnode'"
        assert "Unsupported value" in c.apply("b", dict(BASE, type="fpga"))
        assert c.apply("c", BASE) == ""


def storm_requests(n, n_nodes, seed=20260921):
    """BASELINE config 4: request i -> worker-(i mod 8), model NVIDIA-B200-<i div 8>, size U{1..4}."""
    rng = random.Random(seed)
    return [("req-%04d" % i, {"type": "gpu", "model": "NVIDIA-B200-%d" % (i // n_nodes), "size": rng.randint(1, 4),
                             "allocation_policy": "samenode", "target_node": "worker-%d" % (i % n_nodes)}) for i in range(n)]


def test_storm_reaches_running_on_cpu(cro):
    uuids = ["GPU-%08x-0000-0000-0000-000000000000" % i for i in range(8)]
    with cro.Cluster({"nodes": ["worker-%d" % i for i in range(8)], "uuids": uuids, "probe": False}) as c:
        reqs = storm_requests(200, 8)
        for name, spec in reqs:
            assert c.apply(name, spec) == ""
        stats = c.run()
        assert stats["requests_running"] == 200 and stats["reconcile_errors"] == 0
        assert stats["resources_online"] == sum(s["size"] for _, s in reqs)
        d = c.dump()
        for name, spec in reqs:
            st = d["requests"][name]["status"]
            assert st["state"] == "Running" and len(st["resources"]) == spec["size"]
            node_idx = int(spec["target_node"].split("-")[1])
            for child, cs in st["resources"].items():
                assert cs == {"state": "Online", "device_id": uuids[node_idx], "cdi_device_id": cs["cdi_device_id"],
                              "node_name": spec["target_node"]}
                assert cs["cdi_device_id"].startswith("res-%s-" % name)
                assert d["resources"][child]["status"]["state"] == "Online"
        # tear everything down: finalizers run, children detach, nothing is left
        for name, _ in reqs:
            assert c.delete(name)
        stats = c.run()
        d = c.dump()
        assert d["requests"] == {} and d["resources"] == {}


def test_churn_cycles_on_cpu(cro):
    """BASELINE config 5 shape: compose 4 GPUs -> Online -> decompose, 10 cycles, subset (4c+j) mod 8."""
    uuids = ["GPU-%08x-0000-0000-0000-000000000000" % i for i in range(8)]
    with cro.Cluster({"nodes": ["worker-%d" % i for i in range(8)], "uuids": uuids, "probe": False}) as c:
        for cycle in range(10):
            names = []
            for j in range(4):
                node = (4 * cycle + j) % 8
                name = "churn-%d-%d" % (cycle, j)
                names.append(name)
                assert c.apply(name, {"type": "gpu", "model": "NVIDIA-B200", "size": 1, "target_node": "worker-%d" % node}) == ""
            stats = c.run()
            d = c.dump()
            assert all(d["requests"][n]["status"]["state"] == "Running" for n in names)
            for n in names:
                c.delete(n)
            c.run()
            d = c.dump()
            assert d["requests"] == {} and d["resources"] == {}
        assert stats["reconcile_errors"] == 0
