"""The node side of the reference's table entries through csrc/gpus.cpp (internal/utils/gpus.go restated):
pod choice, every pod-exec, DrainGPU's step order — driven by each entry's OWN mock executor.

tests/golden/reference_entries.json carries, per Entry, the pods its extraHandling creates, whether it
creates the gpu-operator ClusterPolicy, and its gomonkey patch of remotecommand.NewSPDYExecutor as an
ordered rule list: `strings.Contains(url.RawQuery, needle)` -> (stdout, stderr), with the catch-all
"this error should be reported" last.  The harness rebuilds client-go's exec URL query for every
command it issues and lets those rules answer, so a command the reference would not have issued (or
would have spelled differently) trips the catch-all and the entry's expectation fails.
Citations are ":<line of the Entry>"; the fabric side of each entry is the real FM / CM client as in
tests/test_reference_entries.py."""
import json

import pytest

from test_reference_entries import (DCGM_DS, DEV, DP_DS, DRA_DS, ENTRIES, NOW, READY, check, request_for, state_of)

WITH_EXEC = sorted(n for n, e in ENTRIES.items() if "exec_rules" in e)

# DaemonSets each entry creates (appsv1.DaemonSet literals in its extraHandling), by hand
DAEMONSETS = {
    2311: {}, 2426: {DRA_DS: READY}, 2566: {DRA_DS: dict(READY, restarted_at=NOW)},
    2709: {DRA_DS: dict(READY, ready=0, unavailable=1, restarted_at=NOW)}, 2863: {DRA_DS: dict(READY, restarted_at="error")},
    3000: {DRA_DS: READY}, 3192: {DRA_DS: READY},
    5495: {}, 5694: {DRA_DS: READY},
    6747: {DP_DS: READY}, 6873: {DP_DS: READY, DCGM_DS: READY}, 7050: {DP_DS: READY, DCGM_DS: READY},
    7221: {DP_DS: READY, DCGM_DS: READY}, 7398: {DP_DS: READY, DCGM_DS: READY},
    9035: {}, 9187: {DP_DS: READY}, 9366: {DP_DS: READY, DCGM_DS: READY},
}
SLICES = {3000: [{"devices": [{"attributes": {"uuid": DEV}}]}], 3192: [{"devices": [{"attributes": {"uuid": DEV}}]}]}


def cluster_for(e):
    return {"cluster_policy": {"driver_enabled": True} if e["objects"]["cluster_policy"] else None,
            "pods": e.get("pods", []),
            "exec": [{"needle": r["needle"], "stdout": r["stdout"], "stderr": r["stderr"]} for r in e.get("exec_rules", [])]}


def replay(cro, e, **extra):
    r = request_for(e, cluster=cluster_for(e), deleting=(state_of(e) == "Detaching" or e["line"] in (3150, 3166, 3192)), **extra)
    for legacy in ("enumeration", "enumeration_after_remove", "load_check", "drain"):
        r.pop(legacy, None)
    if e["line"] in DAEMONSETS:
        r["daemonsets"] = DAEMONSETS[e["line"]]
    r["resource_slices"] = SLICES.get(e["line"], [])
    if e["line"] in (9035, 9187, 9366):        # FM: the resource is gone upstream (machine ...0003 has no such device)
        from test_reference_entries import objects_for
        r["fabric"]["objects"] = objects_for(e, bmh_uuid="machine0-uuid-temp-fail-000000000003")
    return cro.reconcile_attach(None, r)


# (:9726 / :9864 are about the env, :4454 / :8654 about garbage collection: tests/test_reference_entries.py)
@pytest.mark.parametrize("line", [n for n in WITH_EXEC if n not in (9726, 9864) and not ENTRIES[n].get("expected_deleted")],
                         ids=lambda n: ":%d" % n)
def test_entries_with_their_own_mock_executor(cro, oracle, line):
    e = ENTRIES[line]
    out = replay(cro, e)
    check(cro, oracle, e, out)
    # no command fell through to the mock's catch-all unless the entry is ABOUT such a failure
    for x in out["exec_log"]:
        assert "query" in x and x["query"].endswith("&stderr=true&stdout=true")


def cmds(out):
    return [" ".join(x["argv"][:6])[:90] for x in out["exec_log"]]


def test_drain_sequence_ocp_dra(cro, oracle):
    """:5694 (CM + DRA, device gone upstream): the exact exec sequence of the detach step."""
    out = replay(cro, ENTRIES[5694])
    drv, plug = "nvidia-gpu-operator/nvidia-driver-daemonset-test", "nvidia-dra-driver-gpu/nvidia-dra-driver-gpu-kubelet-plugin-test"
    got = [(x["pod"], x["kind"], x["argv"]) for x in out["exec_log"]]
    assert [(p, k) for p, k, _ in got] == [
        (drv, "command"),      # CheckNoGPULoads: --query-compute-apps
        (drv, "command"),      # DrainGPU: --query-gpu=device_minor,gpu_uuid,pci.bus_id
        (drv, "command"),      # nvidia-smi -i <uuid> -pm 0
        (drv, "fd_scan"),      # who holds /dev/nvidia0
        (drv, "command"),      # rm -f /run/nvidia/driver/dev/nvidia0
        (plug, "command"),     # rm -f /dev/nvidia0 in the kubelet plugin
        (drv, "command"),      # nvidia-smi drain -p 0000:1F:00.0 -m 1
        (drv, "command"),      # nvidia-smi drain -p 0000:1F:00.0 -r
    ]
    argv = [a for _, _, a in got]
    assert argv[0] == ["/usr/bin/nvidia-smi", "--query-compute-apps=gpu_uuid,process_name", "--format=csv,noheader,nounits"]
    assert argv[1] == ["/usr/bin/nvidia-smi", "--query-gpu=device_minor,gpu_uuid,pci.bus_id", "--format=csv,noheader,nounits"]
    assert argv[2] == ["/usr/bin/nvidia-smi", "-i", DEV, "-pm", "0"]
    assert argv[3][:2] == ["sh", "-c"] and 'TARGET_FILE="/dev/nvidia0"' in argv[3][2]
    assert argv[4] == ["/usr/bin/rm", "-f", "/run/nvidia/driver/dev/nvidia0"]
    assert argv[5] == ["/usr/bin/rm", "-f", "/dev/nvidia0"]
    # TrimPrefix("00000000:1F:00.0", "0000") takes ONE "0000" off: nvidia-smi's 8-digit domain becomes the 4-digit one (gpus.go:406)
    assert argv[6] == ["/usr/bin/nvidia-smi", "drain", "-p", "0000:1F:00.0", "-m", "1"]
    assert argv[7] == ["/usr/bin/nvidia-smi", "drain", "-p", "0000:1F:00.0", "-r"]
    assert out["exec_log"][5]["container"] == "compute-domains" and out["exec_log"][0]["container"] == "nvidia-driver-ctr"
    assert out["exec_log"][2]["query"] == ("command=%2Fusr%2Fbin%2Fnvidia-smi&command=-i&command=" + DEV +
                                           "&command=-pm&command=0&container=nvidia-driver-ctr&stderr=true&stdout=true")


def test_drain_sequence_ocp_device_plugin(cro, oracle):
    """:9366 (FM + DEVICE_PLUGIN): no rm steps, and the visibility check afterwards is one more nvidia-smi."""
    out = replay(cro, ENTRIES[9366])
    kinds = [(x["kind"], x["argv"][1] if x["kind"] == "command" else "") for x in out["exec_log"]]
    assert kinds == [("command", "--query-compute-apps=gpu_uuid,process_name"), ("command", "--query-gpu=device_minor,gpu_uuid,pci.bus_id"),
                     ("command", "-i"), ("fd_scan", ""), ("command", "drain"), ("command", "drain"),
                     ("command", "--query-gpu=gpu_uuid")]
    assert out["status"]["state"] == "Deleting"


def test_failed_reset_is_ignored_but_failed_maintenance_is_not(cro, oracle):
    e = json.loads(json.dumps(ENTRIES[9366]))
    for r in e["exec_rules"]:
        if r["needle"] == {"literal": "command=-r"}:
            r["stderr"] = "reset refused"
    out = replay(cro, e)
    assert out["error"] == "" and out["status"]["state"] == "Deleting"               # gpus.go:656-658
    e = json.loads(json.dumps(ENTRIES[9366]))
    for r in e["exec_rules"]:
        if r["needle"] == {"literal": "command=-m&command=1"}:
            r["stderr"] = "no such device"
    out = replay(cro, e)
    assert out["error"] == "detach command 'set maintenance mode' failed: '<nil>', stderr: 'no such device', stdout: ''"
    e = json.loads(json.dumps(ENTRIES[9366]))
    for r in e["exec_rules"]:
        if r["needle"] == {"literal": "command=-pm&command=0"}:
            r["stderr"] = "Unable to set persistence mode"
    out = replay(cro, e)     # the reference's own misspelling is part of the contract
    assert out["error"] == "deatch command 'disable persistence mode' failed: '<nil>', stderr: 'Unable to set persistence mode', stdout: ''"


def test_already_drained_and_no_driver_pod(cro, oracle):
    e = json.loads(json.dumps(ENTRIES[9366]))
    for r in e["exec_rules"]:      # the GPU is no longer enumerated: DrainGPU returns before touching anything
        if r["needle"] == {"escape": "--query-gpu=device_minor,gpu_uuid,pci.bus_id"}:
            r["stdout"] = "1, GPU-other, 00000000:2F:00.0"
    out = replay(cro, e)
    assert out["error"] == "" and [x["argv"][1] for x in out["exec_log"] if x["kind"] == "command"] == [
        "--query-compute-apps=gpu_uuid,process_name", "--query-gpu=device_minor,gpu_uuid,pci.bus_id", "--query-gpu=gpu_uuid"]
    # DRA without a driver pod on the node: nothing to drain (gpus.go:389-393); DEVICE_PLUGIN: an error (:551-554)
    e = json.loads(json.dumps(ENTRIES[5694]))
    e["pods"] = [p for p in e["pods"] if p["name"] != "nvidia-driver-daemonset-test"]
    out = replay(cro, e)
    assert out["error"] == "" and out["exec_log"] == []
    e = json.loads(json.dumps(ENTRIES[9366]))
    e["pods"] = []
    out = replay(cro, e)
    assert out["error"] == "no Pod with label 'app.kubernetes.io/component=nvidia-driver' found on node worker-0"


def test_driver_pod_choice_quirk(cro, oracle):
    """getNvidiaDriverDaemonsetPod checks that a driver pod exists on the node but returns Items[0]
    (gpus.go:827-839): with another node's pod listed first, THAT pod gets the exec."""
    e = json.loads(json.dumps(ENTRIES[7221]))
    e["pods"] = [{"namespace": "nvidia-gpu-operator", "name": "nvidia-driver-daemonset-other", "node": "worker-9",
                  "labels": {"app.kubernetes.io/component": "nvidia-driver"}, "containers": ["ctr-other"]}] + e["pods"]
    out = replay(cro, e)
    assert out["status"]["state"] == "Online"
    assert {x["pod"] for x in out["exec_log"]} == {"nvidia-gpu-operator/nvidia-driver-daemonset-other"}
    assert {x["container"] for x in out["exec_log"]} == {"ctr-other"}


RKE2_PODS = [{"namespace": "cro-system", "name": "cro-node-agent-abcde", "node": "worker-0", "labels": {"app": "cro-node-agent"},
              "containers": ["agent"]}]
PROC2 = "0,%s,0000:1f:00.0\n1,GPU-other,0000:2f:00.0\n" % DEV


def rke2(cro, rules, state="Detaching", pods=RKE2_PODS):
    r = {"name": "cr", "spec": {"type": "gpu", "model": "m", "target_node": "worker-0"},
         "status": {"state": state, "device_id": DEV, "cdi_device_id": "res"}, "deleting": state == "Detaching",
         "device_resource_type": "DRA", "probe": False, "provider": {}, "resource_slices": [],
         "cluster": {"cluster_policy": None, "pods": pods, "exec": rules + [{"needle": None, "stdout": "", "stderr": "this error should be reported"}]}}
    return cro.reconcile_attach(None, r)


def test_rke2_dra_drain_with_other_gpus_left(cro):
    """RKE2 + DRA (gpus.go:196-298): everything goes through the cro-node-agent pod, chrooted; no reference entry covers it."""
    rules = [{"needle": {"escape": "/proc/driver/nvidia/gpus"}, "stdout": PROC2, "stderr": ""},
             {"needle": {"escape": "--query-compute-apps=gpu_uuid,process_name"}, "stdout": "", "stderr": ""},
             {"needle": {"literal": "command=-q"}, "stdout": "GPU 0000:1F:00.0 is currently: not draining\n", "stderr": ""},
             {"needle": {"literal": "command=-pm&command=0"}, "stdout": "", "stderr": ""},
             {"needle": {"escape": "TARGET_FILE"}, "stdout": "", "stderr": ""},
             {"needle": {"literal": "command=-m&command=1"}, "stdout": "", "stderr": ""},
             {"needle": {"escape": "/dev/nvidia0"}, "stdout": "", "stderr": ""},
             {"needle": {"literal": "command=-r"}, "stdout": "", "stderr": ""}]
    out = rke2(cro, rules)
    assert out["error"] == "" and out["status"]["state"] == "Deleting", out
    seq = [(x["kind"], x["argv"][2:7] if x["kind"] == "command" else x["argv"][:2]) for x in out["exec_log"]]
    assert seq == [
        ("proc_scan", ["/bin/chroot", "/host-root"]),                                       # CheckNoGPULoads: is the GPU still there
        ("command", ["/usr/bin/nvidia-smi", "--query-compute-apps=gpu_uuid,process_name", "--format=csv,noheader,nounits"]),
        ("proc_scan", ["/bin/chroot", "/host-root"]),                                       # DrainGPU: minor + bus id
        ("command", ["/usr/bin/nvidia-smi", "drain", "-p", "0000:1F:00.0", "-q"]),          # bus id upper-cased (:218)
        ("command", ["/usr/bin/nvidia-smi", "-i", DEV, "-pm", "0"]),
        ("fd_scan", ["/bin/chroot", "/host-root"]),
        ("command", ["/usr/bin/nvidia-smi", "drain", "-p", "0000:1F:00.0", "-m"]),
        ("command", ["/usr/bin/rm", "-f", "/dev/nvidia0"]),
        ("command", ["/usr/bin/nvidia-smi", "drain", "-p", "0000:1F:00.0", "-r"]),
    ]
    assert all(x["argv"][:2] == ["/bin/chroot", "/host-root"] for x in out["exec_log"])
    # already draining: persistence-mode and maintenance-mode steps are skipped (:262-265)
    rules[2] = {"needle": {"literal": "command=-q"}, "stdout": "GPU 0000:1F:00.0 is currently: draining\n", "stderr": ""}
    out = rke2(cro, rules)
    assert [x["argv"][2:] for x in out["exec_log"] if x["kind"] == "command" and x["argv"][3:4] in (["-i"], ["drain"])] == [
        ["/usr/bin/nvidia-smi", "drain", "-p", "0000:1F:00.0", "-q"], ["/usr/bin/nvidia-smi", "drain", "-p", "0000:1F:00.0", "-r"]]
    # a busy device node stops the drain with the RKE2 wording (:289-291)
    rules[4] = {"needle": {"escape": "TARGET_FILE"}, "stdout": "4242 python3, 77 nvidia-persist", "stderr": ""}
    out = rke2(cro, rules)
    assert out["error"] == "check /dev/nvidiaX command failed: /dev/nvidiaX is in use by one or more processes: 4242 python3, 77 nvidia-persist"


def test_rke2_dra_drain_of_the_last_gpu(cro):
    """The last GPU (gpus.go:300-385): unload nvidia_drm / nvidia_uvm, remove the PCI function through sysfs
    without waiting, unload again, wait a second, then judge by whether the sysfs writer is still there."""
    base = [{"needle": {"escape": "/proc/driver/nvidia/gpus"}, "stdout": "0,%s,0000:1f:00.0\n" % DEV, "stderr": ""},
            {"needle": {"escape": "--query-compute-apps=gpu_uuid,process_name"}, "stdout": "", "stderr": ""},
            {"needle": {"literal": "command=-q"}, "stdout": "currently: not draining", "stderr": ""},
            {"needle": {"literal": "command=-pm&command=0"}, "stdout": "", "stderr": ""},
            {"needle": {"escape": "TARGET_FILE"}, "stdout": "", "stderr": ""},
            {"needle": {"literal": "command=-m&command=1"}, "stdout": "", "stderr": ""},
            {"needle": {"escape": "/usr/sbin/lsmod"}, "stdout": "Module Size Used by\nnvidia_uvm 1 0\nnvidia_drm 2 0\nnvidia 3 2\n", "stderr": ""},
            {"needle": {"escape": "/usr/sbin/modprobe"}, "stdout": "", "stderr": ""},
            {"needle": {"escape": "/usr/bin/tee /sys/bus/pci/devices/0000:1f:00.0/remove"}, "stdout": "", "stderr": ""},
            {"needle": {"escape": "TARGET=\"/sys/bus/pci/devices/0000:1f:00.0/remove\""}, "stdout": "", "stderr": ""},
            {"needle": {"escape": "/dev/nvidia0"}, "stdout": "", "stderr": ""}]
    out = rke2(cro, base)
    assert out["error"] == "" and out["status"]["state"] == "Deleting", out
    tail = [(x["kind"], x["argv"][2:], x["detached"]) for x in out["exec_log"]][8:]
    assert tail == [
        ("command", ["/usr/sbin/lsmod"], False),
        ("command", ["/usr/sbin/modprobe", "-r", "nvidia_drm"], False), ("command", ["/usr/sbin/modprobe", "-r", "nvidia_uvm"], False),
        ("cmdline_scan", ["/bin/sh", "-c", tail_text(out, 11)], False),
        ("command", ["/bin/sh", "-c", "/usr/bin/echo 1 | /usr/bin/tee /sys/bus/pci/devices/0000:1f:00.0/remove > /dev/null"], True),
        ("command", ["/usr/sbin/lsmod"], False),
        ("command", ["/usr/sbin/modprobe", "-r", "nvidia_drm"], False), ("command", ["/usr/sbin/modprobe", "-r", "nvidia_uvm"], False),
        ("cmdline_scan", ["/bin/sh", "-c", tail_text(out, 11)], False),
    ]
    assert out["slept_s"] == 1
    # the writer is still there after the pause: the drain did not complete
    still = [dict(r) for r in base]
    still[9] = dict(still[9], stdout="true\n")
    out = rke2(cro, still)
    assert out["error"] == ("detach command 'reset GPU' did not complete, so it failed to drain the last GPU: targetNodeName=worker-0, "
                            "targetGPUUUID=%s, resetCommandRunning=true, resetCommandError=false" % DEV)
    assert not any(x["detached"] for x in out["exec_log"])          # a running remove is not started twice (:352-354)
    # lsmod failing is fatal, with its own wording
    broken = [dict(r) for r in base]
    broken[6] = dict(broken[6], stdout="", stderr="lsmod: not found")
    out = rke2(cro, broken)
    assert out["error"] == "detach command 'lsmod' failed: '<nil>', stderr: 'lsmod: not found', stdout: ''"


def tail_text(out, i):
    return out["exec_log"][i]["argv"][4]


def test_rke2_paths_of_attach(cro):
    """RunNvidiaSmi without a ClusterPolicy goes through the cro-node-agent pod (gpus.go:673-677); without that
    pod the error is recorded and the attach continues (composableresource_controller.go:258-264)."""
    rules = [{"needle": {"escape": "--query-gpu=gpu_uuid"}, "stdout": DEV, "stderr": ""}]
    out = rke2(cro, rules, state="Attaching")
    assert out["status"].get("error", "") == "" and out["requeue_after_s"] == 30
    assert out["exec_log"][0]["argv"] == ["/bin/chroot", "/host-root", "/usr/bin/nvidia-smi", "--query-gpu=gpu_uuid", "--format=csv,noheader,nounits"]
    out = rke2(cro, rules, state="Attaching", pods=[])
    assert out["status"]["error"] == "no Pod named 'cro-node-agent' found on node worker-0" and out["exec_log"] == []
    # spec.driver.enabled unset is its own error (gpus.go:1237-1239)
    r = {"name": "cr", "spec": {"type": "gpu", "model": "m", "target_node": "worker-0"}, "status": {"state": "Attaching", "device_id": DEV},
         "device_resource_type": "DRA", "probe": False, "provider": {}, "resource_slices": [], "cluster": {"cluster_policy": {}, "pods": [], "exec": []}}
    out = cro.reconcile_attach(None, r)
    assert out["status"]["error"] == "'cluster-policy' nvidia container driver configuration (spec.driver.enabled) is not set"


def test_driver_pod_missing_entries(cro, oracle):
    """:2198 (CM + DRA: recorded in Status.Error, the attach goes on) and :6682 (FM + DEVICE_PLUGIN: CheckGPUVisible
    needs the pod, so the reconcile fails) — no pods, no mock executor: the pod look-up itself is what is exercised."""
    for line in (2198, 6682):
        e = ENTRIES[line]
        assert "exec_rules" not in e and not e.get("pods")
        out = replay(cro, e, daemonsets={DRA_DS: READY, DP_DS: READY, DCGM_DS: READY})
        check(cro, oracle, e, out)
        assert out["exec_log"] == []
