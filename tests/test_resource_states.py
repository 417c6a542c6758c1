"""Online / Detaching states of ComposableResourceReconciler through the C-ABI harness
(internal/controller/composableresource_controller.go:289-407), on the reference's table entries
(internal/controller/composableresource_controller_test.go, CM+DRA block :3410-5945, FM+DEVICE_PLUGIN :7635-9577)."""
import pytest

import __graft_entry__ as g
from test_cm_provider import cm_machine_data
from test_fabric_codec import fm_machine_data

DEV, RES = "GPU-device00-uuid-temp-0000-000000000000", "GPU-device00-uuid-temp-0000-000000000res"
MODEL = "NVIDIA-A100-PCIE-80GB"


def req(state, dtype="DRA", error="", deleting=False, **kw):
    r = {"name": "test-composable-resource", "spec": {"type": "gpu", "model": MODEL, "target_node": "worker-0"},
         "status": {"state": state, "error": error, "device_id": DEV, "cdi_device_id": RES},
         "deleting": deleting, "device_resource_type": dtype, "probe": False, "enumeration": {"stdout": DEV, "stderr": ""}}
    r.update(kw)
    return r


ONLINE_KATS = [   # (cite, provider, expected Status.Error)
    (":3870 stay Online (CM)", {"cm_check_body": cm_machine_data([(DEV, "ADD_COMPLETE", "", RES, "0")])}, ""),
    (":3562/:3633 not found in CM", {"cm_check_body": cm_machine_data(None)}, "the target device '%s' cannot be found in CDI system" % DEV),
    (":3639/:3710 Warning in CM", {"cm_check_body": cm_machine_data([(DEV, "ADD_COMPLETE", "", RES, "1")])}, "the target gpu '%s' is showing a Warning status in CM" % DEV),
    (":3716/:3787 Critical in CM", {"cm_check_body": cm_machine_data([(DEV, "ADD_COMPLETE", "", RES, "2")])}, "the target gpu '%s' is showing a Critical status in CM" % DEV),
    (":3793/:3864 unknown in CM", {"cm_check_body": cm_machine_data([(DEV, "ADD_COMPLETE", "", RES, "3")])}, "the target gpu '%s' has unknown status '3' in CM" % DEV),
    (":7941 stay Online (FM)", {"fm_machine_body": fm_machine_data([(RES, "gpu", "0", DEV, MODEL)])}, ""),
    (":8017/:8088 not found in FM", {"fm_machine_body": fm_machine_data([])}, "the target device '%s' cannot be found in CDI system" % DEV),
    (":8095/:8166 Warning in FM", {"fm_machine_body": fm_machine_data([(RES, "gpu", "1", DEV, MODEL)])}, "the target gpu '%s' is showing a Warning status in FM" % DEV),
    (":8172/:8243 Critical in FM", {"fm_machine_body": fm_machine_data([(RES, "gpu", "2", DEV, MODEL)])}, "the target gpu '%s' is showing a Critical status in FM" % DEV),
    (":8249/:8320 unknown in FM", {"fm_machine_body": fm_machine_data([(RES, "gpu", "3", DEV, MODEL)])}, "the target gpu '%s' has unknown status '3' in FM" % DEV),
]


@pytest.mark.parametrize("cite,provider,want_err", ONLINE_KATS, ids=[k[0] for k in ONLINE_KATS])
def test_online_check_resource_kats(cro, oracle, cite, provider, want_err):
    out = cro.reconcile_attach(None, req("Online", provider=provider))
    assert out["error"] == ""                                   # CheckResource failures are recorded, not returned (:305-310)
    assert g.json_status(out) == oracle.emit_status("Online", want_err, DEV, RES)
    assert out["requeue_after_s"] == 30


def test_online_clears_a_stale_error(cro, oracle):             # :3962 "should clean the error message when in normal state"
    out = cro.reconcile_attach(None, req("Online", error="stale", provider={"cm_check_body": cm_machine_data([(DEV, "ADD_COMPLETE", "", RES, "0")])}))
    assert g.json_status(out) == oracle.emit_status("Online", "", DEV, RES)


def test_online_deleted_goes_detaching(cro, oracle):            # :3946 / :8327
    out = cro.reconcile_attach(None, req("Online", deleting=True, provider={}))
    assert g.json_status(out) == oracle.emit_status("Detaching", "", DEV, RES) and out["requeue_after_s"] == 0


def test_online_ready_to_detach_label_deletes_itself(cro):      # composableresource_controller.go:297-302
    out = cro.reconcile_attach(None, req("Online", provider={}, labels={"cohdi.io/ready-to-detach-device-id": DEV}))
    assert out["delete_requested"] is True and out["status"]["state"] == "Online"


DETACH_KATS = [   # (cite, device_resource_type, extra request keys, expected reconcile error | expected status tuple)
    (":4131/:4192 load check: nvidia-smi failed", "DRA", {"load_check": {"stdout": "", "stderr": "nvidia-smi: command not found"}},
     "run nvidia-smi in pod 'nvidia-driver-daemonset-test' to check gpu loads failed: '<nil>', stderr: 'nvidia-smi: command not found', stdout: ''"),
    (":4194/:4255 gpu loads exist", "DRA", {"load_check": {"stdout": DEV + ", gpu_load_progress", "stderr": ""}},
     "found gpu loads on node 'worker-0': '[GPUUUID: '%s', ProcessName: 'gpu_load_progress']'" % DEV),
    (":8505/:8567 gpu loads exist (DEVICE_PLUGIN)", "DEVICE_PLUGIN", {"load_check": {"stdout": DEV + ", gpu_load_progress", "stderr": ""}},
     "found gpu loads on node 'worker-0': '[GPUUUID: '%s', ProcessName: 'gpu_load_progress']'" % DEV),
    (":4257/:4350 nvidiaX occupied", "DRA", {"drain": {"fd_scan": {"stdout": "nvidia-persist", "stderr": ""}}},
     "check /dev/nvidiaX command failed: there is a process nvidia-persist occupied the nvidiaX file"),
    (":8569/:8652 nvidiaX occupied (DEVICE_PLUGIN)", "DEVICE_PLUGIN", {"drain": {"fd_scan": {"stdout": "nvidia-persist", "stderr": ""}}},
     "check /dev/nvidiaX command failed: there is a process nvidia-persist occupied the nvidiaX file"),
    (":5113/:5290 being removed upstream -> wait", "DRA", {"provider": {"remove": {"waiting": True}}}, ("Detaching", "", DEV, RES, 30)),
    (":5694 removed -> Deleting, ids cleared", "DRA", {"provider": {}, "enumeration_after_remove": {"stdout": "", "stderr": ""},
                                                        "resource_slices_after_remove": []}, ("Deleting", "", "", "", 0)),
    (":9366 removed -> Deleting (DEVICE_PLUGIN)", "DEVICE_PLUGIN", {"provider": {}, "enumeration_after_remove": {"stdout": "", "stderr": ""}},
     ("Deleting", "", "", "", 0)),
    ("composableresource_controller.go:383-387 still visible -> 3 s", "DEVICE_PLUGIN", {"provider": {}}, ("Detaching", "", DEV, RES, 3)),
    (":9035 device-plugin DaemonSet missing is fatal when detaching", "DEVICE_PLUGIN",
     {"provider": {}, "daemonset_errors": {"nvidia-gpu-operator/nvidia-device-plugin-daemonset": "daemonsets.apps \"nvidia-device-plugin-daemonset\" not found"}},
     "daemonsets.apps \"nvidia-device-plugin-daemonset\" not found"),
]


@pytest.mark.parametrize("cite,dtype,extra,expect", DETACH_KATS, ids=[k[0] for k in DETACH_KATS])
def test_detaching_kats(cro, oracle, cite, dtype, extra, expect):
    r = req("Detaching", dtype=dtype, deleting=True, **extra)
    r.setdefault("provider", {})
    out = cro.reconcile_attach(None, r)
    if isinstance(expect, str):
        assert out["error"] == expect
        assert g.json_status(out) == oracle.emit_status("Detaching", expect, DEV, RES)     # requeueOnErr records it
    else:
        state, err, dev, cdi, rq = expect
        assert out["error"] == "" and out["requeue_after_s"] == rq
        assert g.json_status(out) == oracle.emit_status(state, err, dev, cdi)


def test_force_detach_skips_the_load_check(cro):               # composableresource_controller.go:327
    r = req("Detaching", deleting=True, provider={}, load_check={"stdout": DEV + ", busy", "stderr": ""},
            enumeration_after_remove={"stdout": "", "stderr": ""}, resource_slices_after_remove=[])
    r["spec"]["force_detach"] = True
    out = cro.reconcile_attach(None, r)
    assert out["error"] == "" and out["status"]["state"] == "Deleting"
