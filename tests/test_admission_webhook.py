"""The validating webhook (internal/webhook/v1alpha1/composabilityrequest_webhook.go:100-147) as the in-memory API
applies it on create and update, replayed on the reference's two scenarios
(internal/webhook/v1alpha1/composabilityrequest_webhook_test.go:104-185 create, :192-267 update)."""
BASE = {"type": "gpu", "model": "NVIDIA-A100-PCIE-80GB", "size": 2, "allocation_policy": "samenode"}
DENIED = 'admission webhook "vcomposabilityrequest.kb.io" denied the request: '
EXISTS = DENIED + "composabilityRequest resource %s with type gpu and model NVIDIA-A100-PCIE-80GB already exists"
NODES = ["worker-%d" % i for i in range(8)]
CR0, CR1 = "gpu-nvidia-a100-pcie-80gb-00000000-temp-uuid-0000-000000000000", "gpu-nvidia-a100-pcie-80gb-00000000-temp-uuid-0000-000000000001"


def give_children_on_worker0(c, name, spec):
    """`Status().Update` with baseComposabilityRequest.Status: two children on worker-0 (:58-92)."""
    c.plant({"kind": "ComposabilityRequest", "name": name, "resource": spec, "finalizer": False,
             "status": {"state": "", "resources": {CR0: {"node_name": "worker-0", "state": ""}, CR1: {"node_name": "worker-0", "state": ""}},
                        "scalarResource": spec}})


def test_create_scenario(cro):
    with cro.Cluster({"nodes": NODES}) as c:
        assert c.apply("test-composability-request1", dict(BASE, target_node="worker-1")) == ""
        assert c.apply("test-composability-request6", BASE) == ""
        # :131 a second untargeted samenode request for the same (type, model): both resolve to node ""
        assert c.apply("test-composability-request0", BASE) == EXISTS % "test-composability-request6"
        give_children_on_worker0(c, "test-composability-request6", BASE)
        assert c.apply("test-composability-request4", dict(BASE, allocation_policy="differentnode")) == ""
        # :156 request6 now counts for worker-0 (its first child's node)
        assert c.apply("test-composability-request2", dict(BASE, target_node="worker-0")) == EXISTS % "test-composability-request6"
        # :169
        assert c.apply("test-composability-request3", dict(BASE, allocation_policy="differentnode", target_node="worker-0")) == \
            DENIED + "TargetNode cannot be specified when AllocationPolicy is set to 'differentnode'"
        # :181 one differentnode request per (type, model)
        assert c.apply("test-composability-request5", dict(BASE, allocation_policy="differentnode")) == EXISTS % "test-composability-request4"
        assert sorted(c.dump()["requests"]) == ["test-composability-request1", "test-composability-request4", "test-composability-request6"]


def test_update_scenario(cro):
    with cro.Cluster({"nodes": NODES}) as c:
        assert c.apply("test-composability-request1", dict(BASE, target_node="worker-1")) == ""
        assert c.apply("test-composability-request0", BASE) == ""
        give_children_on_worker0(c, "test-composability-request0", BASE)
        assert c.apply("test-composability-request4", dict(BASE, allocation_policy="differentnode")) == ""
        assert c.apply("test-composability-request2", dict(BASE, target_node="worker-7")) == ""
        # :236 moving request2 onto worker-0 collides with request0, which lives there
        assert c.apply("test-composability-request2", dict(BASE, target_node="worker-0")) == EXISTS % "test-composability-request0"
        assert c.dump()["requests"]["test-composability-request2"]["spec"]["target_node"] == "worker-7"      # the update was refused
        assert c.apply("test-composability-request3", dict(BASE, target_node="worker-5")) == ""
        # :251
        assert c.apply("test-composability-request3", dict(BASE, target_node="worker-5", allocation_policy="differentnode")) == \
            DENIED + "TargetNode cannot be specified when AllocationPolicy is set to 'differentnode'"
        assert c.apply("test-composability-request5", dict(BASE, model="NVIDIA-H100-PCIE-80GB", allocation_policy="differentnode")) == ""
        # :266 changing request5's model to the one request4 already covers
        assert c.apply("test-composability-request5", dict(BASE, allocation_policy="differentnode")) == EXISTS % "test-composability-request4"
