"""The node's inventory is re-read on every enumeration: the reference execs a fresh nvidia-smi on
each reconcile (internal/utils/gpus.go:666-689, :878-919), so a GPU composed after the agent started must be listed
and a GPU drained off the bus must stop being listed.  A fake /proc tree is mutated between calls; the probe of a
device that arrived late goes through the helper process (a stand-in script here: no GPU in this container)."""
import ctypes
import os
import stat
import struct
import sys
import time

import pytest

INFO = """Model: \t\t NVIDIA B200
IRQ:   \t\t %d
GPU UUID: \t %s
Video BIOS: \t 97.00.82.00.2e
Bus Type: \t PCIe
DMA Size: \t 52 bits
DMA Mask: \t 0xfffffffffffff
Bus Location: \t %s
Device Minor: \t %d
GPU Excluded:\t No
"""
U = ["GPU-%08x-aaaa-bbbb-cccc-%012x" % (i, i) for i in range(4)]
BUS = ["0000:%02x:00.0" % b for b in (0x1b, 0x43, 0x52, 0x61)]


def put(root, i, minor=None):
    d = os.path.join(root, "driver", "nvidia", "gpus", BUS[i])
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "information"), "w") as f:
        f.write(INFO % (100 + i, U[i], BUS[i], i if minor is None else minor))


def drop(root, i):
    d = os.path.join(root, "driver", "nvidia", "gpus", BUS[i])
    os.remove(os.path.join(d, "information"))
    os.rmdir(d)


def known(cro, i, dev_index):
    d = cro.DevInfo()
    d.cuda_ordinal, d.device_minor, d.gpu_uuid, d.pci_bus_id = dev_index, i, U[i].encode(), ("0000" + BUS[i]).upper().encode()
    d.name, d.identity_source = b"NVIDIA B200", 2
    return d


def rows(devs):
    return [(d.gpu_uuid.decode(), d.device_minor, d.pci_bus_id.decode(), d.flags, d.dev_index) for d in devs]


def test_enumeration_follows_the_node(cro, tmp_path):
    root = str(tmp_path)
    mine = [known(cro, 0, 0), known(cro, 1, 1)]
    # no registry at all: nothing to go by but the context's own devices
    assert rows(cro.node_inventory(root, mine)) == [(U[0], 0, "00000000:1B:00.0", cro.DEV_IN_PROCESS, 0),
                                                    (U[1], 1, "00000000:43:00.0", cro.DEV_IN_PROCESS, 1)]
    put(root, 0), put(root, 1)
    assert [r[0] for r in rows(cro.node_inventory(root, mine))] == [U[0], U[1]]
    # a GPU is composed AFTER the agent initialised CUDA: listed at once, flagged for the helper process
    put(root, 2)
    got = rows(cro.node_inventory(root, mine))
    assert got[2] == (U[2], 2, "00000000:52:00.0", cro.DEV_NEEDS_HELPER, -1) and len(got) == 3
    assert cro.CheckGPUVisible(cro.node_inventory(root, mine), U[2])          # the reference's membership test (gpus.go:78-82)
    assert cro.emit_csv(cro.node_inventory(root, mine), "gpu_uuid").split("\n")[:3] == [U[0], U[1], U[2]]
    # a GPU is drained off the bus (`nvidia-smi drain -r`, /sys/bus/pci/devices/<id>/remove): no longer listed,
    # so Detaching sees visible=false and can finish (composableresource_controller.go:381-394)
    drop(root, 1)
    got = rows(cro.node_inventory(root, mine))
    assert [r[0] for r in got] == [U[0], U[2]] and got[0][4] == 0
    assert not cro.CheckGPUVisible(cro.node_inventory(root, mine), U[1])
    # ... and back under ANOTHER minor (re-bound): the fresh minor wins, the in-process handle stays usable
    put(root, 1, minor=7)
    got = rows(cro.node_inventory(root, mine))
    assert [r[0] for r in got] == [U[0], U[2], U[1]] and got[2][1] == 7 and got[2][3] == cro.DEV_IN_PROCESS and got[2][4] == 1
    # the last GPU leaves: the driver's registry is there but empty -> "No devices were found" (gpus.go:896-898)
    for i in (0, 1, 2):
        drop(root, i)
    assert cro.node_inventory(root, mine) == []
    assert cro.emit_csv([], "gpu_uuid") == "No devices were found\n"


def test_bus_id_spelling(cro, tmp_path):
    root = str(tmp_path)
    put(root, 3)
    (d,) = cro.node_inventory(root, [])
    assert d.pci_bus_id == b"00000000:61:00.0" and d.device_minor == 3 and d.identity_source == 2
    assert cro.normalize(2, d.pci_bus_id.decode()) == "0000:61:00.0"          # what `nvidia-smi drain -p` takes (gpus.go:406)


def fake_helper(tmp_path, body):
    p = os.path.join(str(tmp_path), "fake-croprobe-cli")
    with open(p, "w") as f:
        f.write("#!%s\n" % sys.executable + body)
    os.chmod(p, os.stat(p).st_mode | stat.S_IXUSR)
    return p


def test_late_device_is_probed_through_the_helper_process(cro, tmp_path, monkeypatch):
    """cro_probe_uuid without an in-process handle for the device: fork/exec of the helper with
    CUDA_VISIBLE_DEVICES=<uuid>, 512-byte struct back on its stdout."""
    helper = fake_helper(tmp_path, """
import os, struct, sys
assert sys.argv[1] == "probe-raw" and os.environ["CUDA_VISIBLE_DEVICES"] == sys.argv[2], sys.argv
r = bytearray(512)
struct.pack_into("<Iiii", r, 0, 2, 0, 0, 5)
r[16:16 + len(sys.argv[2])] = sys.argv[2].encode()
struct.pack_into("<Q", r, 96, int(sys.argv[3]) << 20)
sys.stdout.buffer.write(bytes(r))
""")
    monkeypatch.setenv("CRO_HELPER_PATH", helper)
    r = cro.probe_uuid(None, U[2])
    assert r.status == 0 and r.abi_version == 2 and r.gpu_uuid.decode() == U[2] and r.device_minor == 5
    assert r.sweep_bytes == 1 << 30                                         # the helper's default first sweep


def test_helper_failures_are_loud(cro, tmp_path, monkeypatch):
    monkeypatch.setenv("CRO_HELPER_PATH", fake_helper(tmp_path, "import sys\nsys.exit(3)\n"))
    with pytest.raises(cro.ProbeError) as e:
        cro.probe_uuid(None, U[0])
    assert e.value.code == cro.ERR_NO_DEVICE and "not visible to a fresh CUDA process" in str(e.value)
    monkeypatch.setenv("CRO_HELPER_PATH", fake_helper(tmp_path, "import sys\nsys.stdout.write('short')\n"))
    with pytest.raises(cro.ProbeError) as e:
        cro.probe_uuid(None, U[0])
    assert e.value.code == cro.ERR_EXEC and "5 result bytes" in str(e.value)
    monkeypatch.setenv("CRO_HELPER_PATH", os.path.join(str(tmp_path), "missing"))
    with pytest.raises(cro.ProbeError) as e:
        cro.probe_uuid(None, U[0])
    assert e.value.code == cro.ERR_EXEC and "not executable" in str(e.value)


def test_wedged_helper_is_killed_at_the_deadline(cro, tmp_path, monkeypatch):
    """A helper stuck on a GPU that is mid-drain must not hang the reconcile worker: SIGKILL + reap at the deadline."""
    monkeypatch.setenv("CRO_HELPER_PATH", fake_helper(tmp_path, "import time\ntime.sleep(60)\n"))
    monkeypatch.setenv("CRO_HELPER_TIMEOUT_MS", "300")
    assert cro.validate_env() == ""
    t0 = time.monotonic()
    with pytest.raises(cro.ProbeError) as e:
        cro.probe_uuid(None, U[0])
    assert e.value.code == cro.ERR_DEADLINE and "was killed" in str(e.value)
    assert time.monotonic() - t0 < 5
