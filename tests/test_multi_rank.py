"""The N>1 host logic on CPU: world_size-2 gloo all-gather of the 512-byte result structs,
then the rank-0 emit — the same code path bench.py runs over NCCL."""
import importlib
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    cro = importlib.import_module("composable-resource-operator_b200")
    multirank = importlib.import_module("composable-resource-operator_b200.multirank")
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    r = cro.ProbeResult()
    r.abi_version, r.status, r.rank, r.world = 2, 0, rank, world
    r.gpu_uuid = ("GPU-%08x-aaaa-bbbb-cccc-dddddddddddd" % rank).encode()
    r.pci_bus_id = ("00000000:%02X:00.0" % (0x1B + rank)).encode()
    r.device_minor = rank
    r.sweep_bytes, r.read_best_ns, r.fill_ns, r.copy_best_ns, r.copy_sweeps = 4 << 30, 575000 + rank, 580000, 1240000, 5
    r.checksum_xor, r.checksum_sum = 0x1111 * (rank + 1), 0x2222
    got = multirank.all_gather_results(dist, r)
    problem = multirank.check_gathered(got, world)
    out = {"problem": problem, "uuids": [g.gpu_uuid.decode() for g in got], "ranks": [g.rank for g in got],
           "bytes_equal_own": multirank.result_to_bytes(got[rank]) == multirank.result_to_bytes(r)}
    if rank == 0:
        out["annotations"] = [cro.emit_probe_annotations_json(g) for g in got]
        out["status"] = [cro.emit_status_json("Online", "", g.gpu_uuid.decode(), "res-%d-0" % g.rank) for g in got]
        devs = []
        for g in got:
            d = cro.DevInfo()
            d.device_minor, d.gpu_uuid, d.pci_bus_id = g.device_minor, g.gpu_uuid, g.pci_bus_id
            devs.append(d)
        out["csv"] = cro.emit_csv(devs, "device_minor,gpu_uuid,pci.bus_id")
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_all_gather(cro):
    import json
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank in (0, 1):
        o = outs[rank]
        assert o["problem"] is None and o["ranks"] == [0, 1] and o["bytes_equal_own"]
        assert o["uuids"] == ["GPU-00000000-aaaa-bbbb-cccc-dddddddddddd", "GPU-00000001-aaaa-bbbb-cccc-dddddddddddd"]
    a = [json.loads(x) for x in outs[0]["annotations"]]
    assert a[0]["cohdi.io/probe-hbm-read-gbs"] == "7469.5" and a[1]["cohdi.io/probe-device-minor"] == "1"
    assert outs[0]["csv"] == ("0, GPU-00000000-aaaa-bbbb-cccc-dddddddddddd, 00000000:1B:00.0\n"
                              "1, GPU-00000001-aaaa-bbbb-cccc-dddddddddddd, 00000000:1C:00.0\n")
    assert outs[0]["status"][1] == '{"state":"Online","device_id":"GPU-00000001-aaaa-bbbb-cccc-dddddddddddd","cdi_device_id":"res-1-0"}'


def test_check_gathered_flags_duplicates(cro):
    multirank = importlib.import_module("composable-resource-operator_b200.multirank")
    a, b = cro.ProbeResult(), cro.ProbeResult()
    a.gpu_uuid = b.gpu_uuid = b"GPU-same"
    assert "same device" in multirank.check_gathered([a, b], 2)
    b.gpu_uuid = b"GPU-other"
    b.status = cro.ERR_CHECKSUM
    assert "status" in multirank.check_gathered([a, b], 2)
    assert multirank.check_gathered([a], 2).startswith("expected 2")
