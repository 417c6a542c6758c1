"""One tiny full-box probe (every visible GPU, S = 4 MiB, 1 MiB per NVLink leg, 64-hop chases, in-library all-gather) for
runs under compute-sanitizer:  compute-sanitizer --tool memcheck python tools/tiny_fullbox.py"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import __graft_entry__ as g

g.build()
import oracle

cro = importlib.import_module("composable-resource-operator_b200")
co = oracle.COracle()
S, P, HOPS = 4 << 20, 1 << 20, 64
flags = cro.F_SKIP_NCCL if os.environ.get("TINY_SKIP_NCCL") else 0
with cro.ProbeContext(sweep_bytes=S, p2p_bytes=P, read_sweeps=2, copy_sweeps=2, latency_hops=HOPS, flags=flags) as ctx:
    n = ctx.device_count()
    devs = ctx.own_devices()
    for rep in range(2):
        res = ctx.probe_all()
        for i, r in enumerate(res):
            assert r.status == 0 and r.checksum == co.checksum(r.seed, 0, S // 8), (i, r.status, r.fail_code, r.fail_index)
            for j in range(n):
                if j == i:
                    continue
                d = ctx.p2p_detail(i, j)
                assert (d.read_xor, d.read_sum, d.read_wsum) == co.checksum(res[j].seed, 0, P // 8), (i, j)
                assert (d.landed_xor, d.landed_sum, d.landed_wsum) == co.checksum(r.seed, 0, P // 8), (i, j)
                assert d.chase_end == co.chase_end(max(devs[i].device_minor, 0), max(devs[j].device_minor, 0), HOPS), (i, j)
    print("tiny full-box probe ok on", n, "GPUs, host syncs per call:", ctx.fullbox_times().host_syncs)
