"""Sweep-size dependence of the three probe kernels (S = 256 MiB, 1 GiB, 4 GiB, 16 GiB) and the
probe as a whole.  JSON lines on stdout."""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g

g.build()
cro = importlib.import_module("composable-resource-operator_b200")

for gib in (0.015625, 0.0625, 0.125, 0.25, 1, 4, 16, 32):
    S = int(gib * (1 << 30))
    iters = max(3, min(40, int(16 / max(gib, 0.25))))
    with cro.ProbeContext(sweep_bytes=S, devices=[0]) as c:
        c.hbm_fill(0)
        row = {"sweep_gib": gib, "iters": iters}
        for name, fn in (("read_tma", lambda: c.hbm_read_checksum(0, cro.READ_TMA, iters)),
                         ("read_ldg", lambda: c.hbm_read_checksum(0, cro.READ_LDG, iters)),
                         ("read_ldg256", lambda: c.hbm_read_checksum(0, cro.READ_LDG256, iters)),
                         ("copy_fused", lambda: c.hbm_copy(0, cro.COPY_TMA_FUSED, iters)),
                         ("copy_tma", lambda: c.hbm_copy(0, cro.COPY_TMA, iters)),
                         ("fill", lambda: c.hbm_fill(0, iters))):
            fn()
            best = max(r.bytes / r.ns for r in (fn() for _ in range(3)))
            row[name + "_gbs"] = round(best, 1)
        probes = [c.probe_device(0) for _ in range(5)]
        assert all(p.status == 0 and p.copy_verified == p.copy_sweeps for p in probes)
        best = min(p.total_ns for p in probes)
        row["probe_ms"] = round(best / 1e6, 3)
        row["probes_per_s"] = round(1e9 / best, 1)
        row["probe_effective_gbs"] = round(16 * S / best, 1)
        print(json.dumps(row), flush=True)
