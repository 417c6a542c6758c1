"""The node-side operations run ON the node (cro_local_node_op, csrc/gpus_local.cpp): the same restated flows
as tests/test_node_side_entries.py, but with the Exec seam answered locally — native /proc scans, spawned
read-only commands, mutating commands only on request.  CPU tests use a fake /proc; the GPU tests run the real
read-only checks against the box."""
import os

import pytest

DEV = "GPU-7cc45b7b-2a6d-f0ac-1b02-6f8de09e1a6c"


def fake_proc(tmp_path, gpus, procs=()):
    root = tmp_path / "proc"
    for bus, (minor, uuid) in gpus.items():
        d = root / "driver" / "nvidia" / "gpus" / bus
        d.mkdir(parents=True)
        (d / "information").write_text("Model: \t NVIDIA B200\nGPU UUID: \t %s\nDevice Minor: \t %s\nBus Location: \t %s\n" % (uuid, minor, bus))
    for pid, cmdline in procs:
        p = root / str(pid)
        p.mkdir(parents=True)
        (p / "cmdline").write_bytes(b"\0".join(a.encode() for a in cmdline) + b"\0")
        (p / "comm").write_text(cmdline[0].rsplit("/", 1)[-1] + "\n")
        (p / "fd").mkdir()
    root.mkdir(exist_ok=True)
    return str(root)


def test_cmdline_scan(cro, tmp_path):
    target = "/sys/bus/pci/devices/0000:1f:00.0/remove"
    root = fake_proc(tmp_path, {}, [(100, ["/bin/sh", "-c", "/usr/bin/echo 1 | /usr/bin/tee %s > /dev/null" % target]), (101, ["/usr/bin/sleep", "9"])])
    assert cro.scan_cmdline_for(root, target) is True
    assert cro.scan_cmdline_for(root, "/sys/bus/pci/devices/0000:2f:00.0/remove") is False
    assert cro.scan_cmdline_for(str(tmp_path / "nowhere"), target) is False


def test_rke2_flavour_with_a_fake_proc(cro, tmp_path):
    req = {"node": "worker-0", "device_id": DEV, "device_resource_type": "DRA", "driver_container": False}
    # the GPU is not under /proc/driver/nvidia/gpus any more: nothing to check, nothing to drain (gpus.go:109-121, :227-230)
    root = fake_proc(tmp_path / "a", {"0000:2f:00.0": ("1", "GPU-other")})
    for op in ("check_no_gpu_loads", "drain"):
        out = cro.local_node_op(None, dict(req, op=op, proc_root=root))
        assert out["error"] == "" and [(x["kind"], x["how"]) for x in out["exec_log"]] == [("proc_scan", "native")], out
    # it is there: the next step needs nvidia-smi, which this container does not have — the spawn error is the exec error
    root = fake_proc(tmp_path / "b", {"0000:1f:00.0": ("0", DEV)})
    out = cro.local_node_op(None, dict(req, op="drain", proc_root=root))
    assert [(x["kind"], x["how"]) for x in out["exec_log"]] == [("proc_scan", "native"), ("command", "spawned")]
    assert out["exec_log"][1]["argv"] == ["/bin/chroot", "/host-root", "/usr/bin/nvidia-smi", "drain", "-p", "0000:1F:00.0", "-q"]
    if not os.path.exists("/usr/bin/nvidia-smi"):
        assert out["error"].startswith("check gpu drain status command failed: 'exec: \"/usr/bin/nvidia-smi\": No such file or directory'")
    with pytest.raises(cro.ProbeError):
        cro.local_node_op(None, dict(req, op="reboot"))


@pytest.mark.gpu
def test_local_checks_on_the_box(cro):
    """Real read-only checks: this very process holds a CUDA context on GPU 0, so the load check finds a compute app
    and the open-file scan finds a holder of /dev/nvidia<minor> — both are the reference's refusals, spelled its way."""
    with cro.ProbeContext(sweep_bytes=64 << 20, devices=[0], read_sweeps=1, copy_sweeps=1) as ctx:
        info = ctx.own_devices()[0]
        uuid = info.gpu_uuid.decode()
        ctx.probe_device(0)
        base = {"node": "worker-0", "device_id": uuid, "driver_container": True}
        out = cro.local_node_op(ctx, dict(base, op="run_nvidia_smi"))
        assert out["error"] == "" and out["exec_log"][0]["how"] == "native"          # answered from the enumeration
        out = cro.local_node_op(ctx, dict(base, op="check_gpu_visible", device_resource_type="DEVICE_PLUGIN"))
        assert out["error"] == "" and out["visible"] is True
        out = cro.local_node_op(ctx, dict(base, op="check_gpu_visible", device_resource_type="DEVICE_PLUGIN", device_id="GPU-nope"))
        assert out["visible"] is False
        out = cro.local_node_op(ctx, dict(base, op="check_no_gpu_loads", device_resource_type="DEVICE_PLUGIN"))
        assert out["exec_log"][0]["how"] == "spawned" and out["exec_log"][0]["argv"][1] == "--query-compute-apps=gpu_uuid,process_name"
        assert out["error"] == "" or out["error"].startswith("found gpu loads on node 'worker-0': '[GPUUUID: '")
        # dry run of the drain: enumeration (with device_minor, which driver 580's nvidia-smi refuses to print) is native,
        # persistence mode is skipped, and the open-file scan stops the drain because WE hold the device node
        out = cro.local_node_op(ctx, dict(base, op="drain", device_resource_type="DEVICE_PLUGIN"))
        hows = [(x["kind"], x["how"]) for x in out["exec_log"]]
        assert hows[:3] == [("command", "native"), ("command", "skipped (dry run)"), ("fd_scan", "native")], out
        assert out["exec_log"][2]["argv"][2].startswith('TARGET_FILE="/dev/nvidia%d"' % info.device_minor)
        assert out["error"].startswith("check /dev/nvidiaX command failed: there is a process ") and "occupied the nvidiaX file" in out["error"]
        print("local drain dry run:", out["error"].strip(), hows)
