"""Detach-side pre-flight: load check, drain-status parse, open-file scan.
KATs from internal/controller/composableresource_controller_test.go; fuzz against the oracle;
the native /proc fd scan against a fake /proc and against this very process."""
import os
import random

from test_oracle_kats import NASTY, rand_text

DEV = "GPU-device00-uuid-temp-0000-000000000000"


def test_gpu_load_kat(cro, oracle):
    # :4247 mock stdout for --query-compute-apps, :4255 / :8567 expected error
    so = "GPU-device00-uuid-temp-0000-000000000000, gpu_load_progress"
    want = "found gpu loads on node 'worker-0': '[GPUUUID: 'GPU-device00-uuid-temp-0000-000000000000', ProcessName: 'gpu_load_progress']'"
    assert cro.CheckNoGPULoadsFromOutput(so, "", None, "nvidia-driver-daemonset-test", "worker-0", None, True) == want
    assert oracle.check_no_gpu_loads(so, "", None, "nvidia-driver-daemonset-test", "worker-0", None, True) == want
    # no load
    assert cro.CheckNoGPULoadsFromOutput("", "", None, "p", "worker-0", None, True) == ""
    assert cro.CheckNoGPULoadsFromOutput("No devices were found\n", "boom", "exit 9", "p", "worker-0", None, True) == ""
    # RKE2 branch: only the target GPU matters
    two = so + "\nGPU-other, x\n"
    assert cro.CheckNoGPULoadsFromOutput(two, "", None, "p", "worker-0", "GPU-none", False) == ""
    assert cro.CheckNoGPULoadsFromOutput(two, "", None, "p", "worker-0", DEV, False) == \
        "found gpu load on gpu '%s': [GPUUUID: '%s', ProcessName: 'gpu_load_progress' GPUUUID: 'GPU-other', ProcessName: 'x']" % (DEV, DEV)


def test_fd_scan_kat(cro, oracle):
    # :4342 mock stdout "nvidia-persist", :4350 / :8652 expected error
    want = "check /dev/nvidiaX command failed: there is a process nvidia-persist occupied the nvidiaX file"
    assert cro.CheckDeviceFileScanResult("nvidia-persist", "", None) == want == oracle.check_device_file_scan("nvidia-persist", "", None)
    assert cro.CheckDeviceFileScanResult("", "", None) == ""
    assert cro.CheckDeviceFileScanResult("1234 python", "", None, rke2=True) == \
        "check /dev/nvidiaX command failed: /dev/nvidiaX is in use by one or more processes: 1234 python"


def test_drain_status(cro, oracle):
    cases = [
        ("GPU 00000000:1F:00.0 is currently being drained.\nDrain state: draining.\n", (True, "")),
        ("Successfully ...\nGPU 0000:1F:00.0 drain state: Not Draining\n", (False, "")),
        ("", (False, "nvidia-smi drain query returned empty output (node=worker-0, busID=0000:1F:00.0)")),
        ("something else entirely", (False, "nvidia-smi drain query did not contain recognizable drain state (node=worker-0, busID=0000:1F:00.0, raw=something else entirely)")),
    ]
    for so, want in cases:
        assert cro.checkGPUDrainStatusFromOutput(so, "", None, "worker-0", " 0000:1F:00.0 ") == want
        assert oracle.check_gpu_drain_status(so, "", None, "worker-0", " 0000:1F:00.0 ") == want
    assert cro.checkGPUDrainStatusFromOutput("x", "", None, "n", "  ")[1] == "target GPU bus ID is empty"
    assert cro.checkGPUDrainStatusFromOutput("x", "bad", None, "n", "b")[1] == \
        "check gpu drain status command failed: '<nil>', stderr: 'bad', stdout: 'x'"


def test_detach_fuzz_vs_oracle(cro, oracle):
    rng = random.Random(77)
    words = ["drain", "Draining", "not draining", ":", ".", "state", "GPU", "\n", " ", "ok", "Drain state: draining..", "NOT DRAINING"]
    for _ in range(2000):
        so = rand_text(rng, NASTY, rng.randrange(0, 30))
        se = "" if rng.random() < 0.85 else "err"
        ee = None if rng.random() < 0.9 else "exit status 2"
        tgt = rng.choice([None, "GPU-a", "a", ""])
        drv = rng.random() < 0.5
        if not drv and tgt is None:
            tgt = "GPU-a"
        assert cro.CheckNoGPULoadsFromOutput(so, se, ee, "pod", "node", tgt, drv) == oracle.check_no_gpu_loads(so, se, ee, "pod", "node", tgt, drv), (so, se, ee, tgt, drv)
        ds = "".join(rng.choice(words) for _ in range(rng.randrange(0, 8)))
        assert cro.checkGPUDrainStatusFromOutput(ds, se, ee, "node", "0000:1F:00.0") == oracle.check_gpu_drain_status(ds, se, ee, "node", "0000:1F:00.0"), ds
        rk = rng.random() < 0.5
        assert cro.CheckDeviceFileScanResult(so, se, ee, rk) == oracle.check_device_file_scan(so, se, ee, rk)


def test_native_fd_scan_fake_proc(cro, tmp_path):
    target = tmp_path / "dev_nvidia0"
    target.write_text("")
    other = tmp_path / "other"
    other.write_text("")
    proc = tmp_path / "proc"
    for pid, comm, links in (("100", "bash", [other]), ("23", "nvidia-persist", [other, target]), ("7", "python", [target, target])):
        d = proc / pid / "fd"
        d.mkdir(parents=True)
        (proc / pid / "comm").write_text(comm + "\n")
        for i, t in enumerate(links):
            os.symlink(t, d / str(i))
    (proc / "self").mkdir()
    # OCP script: lexical PID order (100, 23, 7): first holder is pid 23
    assert cro.scan_device_file_holders(str(target), str(proc)) == "nvidia-persist\n"
    # RKE2 script: every matching link, "PID comm" joined by ", "
    assert cro.scan_device_file_holders(str(target), str(proc), rke2=True) == "23 nvidia-persist, 7 python, 7 python"
    assert cro.scan_device_file_holders(str(tmp_path / "nobody_has_this"), str(proc)) == ""
    err = cro.CheckDeviceFileScanResult(cro.scan_device_file_holders(str(target), str(proc)).strip(), "", None)
    assert err == "check /dev/nvidiaX command failed: there is a process nvidia-persist occupied the nvidiaX file"


def test_native_fd_scan_real_proc(cro, tmp_path):
    """Against the live /proc: this process holds a file open, the scan must find python."""
    f = tmp_path / "held"
    f.write_text("x")
    with open(f):
        out = cro.scan_device_file_holders(str(f), None, rke2=True)
    assert str(os.getpid()) + " " in out
