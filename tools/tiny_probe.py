"""One tiny full probe (S = 4 MiB, every read variant, verify-copy) for runs under compute-sanitizer."""
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
S = 4 << 20
with cro.ProbeContext(sweep_bytes=S, devices=[0], read_sweeps=2, copy_sweeps=2) as ctx:
    r = ctx.probe_device(0)
    want = oracle.COracle().checksum(r.seed, 0, S // 8)
    assert r.checksum == want == r.copy_checksum, (r.status, want)
    for v in (cro.READ_LDG, cro.READ_TMA, cro.READ_LDG256):
        s = ctx.hbm_read_checksum(0, v)
        assert s.checksum == want
    assert ctx.hbm_copy(0, cro.COPY_TMA_FUSED).checksum == want and r.copy_verified == 2
    print("tiny probe ok", hex(r.checksum_xor))
