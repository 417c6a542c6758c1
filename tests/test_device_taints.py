"""DeviceTaintRule bookkeeping of the DRA detach (internal/utils/gpus.go:691-786) inside the Detaching step
(internal/controller/composableresource_controller.go:339-347, 389-397): taint first, drain, remove upstream,
untaint only once the device is really gone.  No reference entry asserts on the taint objects themselves (the
envtest API server just stores them), so the expectations here are read off the code: unpinned, stated."""
from test_node_side_entries import ENTRIES, replay

DEV = "GPU-device00-uuid-temp-0000-000000000000"
SLICE = [{"driver": "gpu.nvidia.com", "pool": {"name": "worker-0"}, "devices": [{"name": "gpu-0", "attributes": {"uuid": DEV}}]}]


def run(cro, e, slices, **cluster_extra):
    import test_node_side_entries as m
    orig = m.cluster_for

    def patched(x):
        c = orig(x)
        c.update(cluster_extra)
        return c
    m.cluster_for = patched
    m.SLICES[e["line"]] = slices
    try:
        return replay(cro, e)
    finally:
        m.cluster_for = orig
        m.SLICES.pop(e["line"], None)


def test_taint_is_created_before_the_drain_and_kept_while_the_device_is_visible(cro):
    # :5113 — CM accepted the resize, the device is still being removed: requeue, the taint stays
    out = run(cro, ENTRIES[5113], SLICE)
    assert out["error"] == "" and out["requeue_after_s"] == 30
    assert out["taint_ops"] == ["create test-composable-resource-taint driver=gpu.nvidia.com pool=worker-0 device=gpu-0 "
                                "k8s.io/device-uuid=%s:NoSchedule" % DEV]
    # the taint exists already: no second create (gpus.go:694-696)
    out = run(cro, ENTRIES[5113], SLICE, taints=["test-composable-resource-taint"])
    assert out["taint_ops"] == []
    # the ResourceSlice does not list the device any more: nothing to taint (:724-727)
    out = run(cro, ENTRIES[5113], [])
    assert out["taint_ops"] == []


def test_taint_is_deleted_once_the_device_is_gone(cro):
    # :5694 — gone upstream and not in any ResourceSlice: Deleting, taint removed if there was one
    out = run(cro, ENTRIES[5694], [], taints=["test-composable-resource-taint"])
    assert out["status"]["state"] == "Deleting" and out["taint_ops"] == ["delete test-composable-resource-taint"]
    out = run(cro, ENTRIES[5694], [])
    assert out["status"]["state"] == "Deleting" and out["taint_ops"] == []


def test_taint_api_failures_carry_the_references_wording(cro):
    out = run(cro, ENTRIES[5113], SLICE, taint_create_error="admission webhook denied the request")
    assert out["error"] == "failed to create DeviceTaintRule test-composable-resource-taint: admission webhook denied the request"
    assert out["exec_log"][1:] == [] or all(x["kind"] != "fd_scan" for x in out["exec_log"])      # the drain never started
    out = run(cro, ENTRIES[5694], [], taints=["test-composable-resource-taint"], taint_delete_error="conflict")
    assert out["error"] == "failed to delete DeviceTaintRule test-composable-resource-taint: conflict"
    out = run(cro, ENTRIES[5694], [], taint_get_error="etcd leader changed")
    assert out["error"] in ("etcd leader changed",                                                  # CreateDeviceTaint returns the Get error bare (:697-699)
                            "failed to get DeviceTaintRule test-composable-resource-taint: etcd leader changed")


def test_device_plugin_detach_never_touches_taints(cro):
    out = run(cro, ENTRIES[9366], SLICE)
    assert out["taint_ops"] == []
