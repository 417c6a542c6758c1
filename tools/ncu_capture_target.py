"""The workload for `ncu --set full`: one launch each of fill / read / checksumming copy / plain TMA copy / closed-form
generator at S = 1 GiB (ncu replays every kernel about 40 times, so keep it to a handful of launches).  Run as:  ncu --set full --clock-control none --import-source on
-k regex:hbm_ -o gpurun_out/<round>/full python tools/ncu_capture_target.py"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

g.build()
cro = importlib.import_module("composable-resource-operator_b200")
S = int(float(sys.argv[1]) * (1 << 30)) if len(sys.argv) > 1 else 1 << 30
with cro.ProbeContext(sweep_bytes=S, devices=[0], flags=cro.F_LAZY_ALLOC) as ctx:
    ctx.hbm_fill(0)
    ctx.hbm_read_checksum(0, cro.READ_AUTO)
    ctx.hbm_copy(0, cro.COPY_AUTO)          # the checksumming copy (hbm_copy_fused_kernel)
    ctx.hbm_copy(0, cro.COPY_TMA)           # the plain one it replaced, for comparison
    ctx.hbm_expected_checksum(0)
print("captured fill / read / fused copy / plain copy / closed form at S =", S)
