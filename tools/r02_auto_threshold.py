"""Which read kernel should AUTO pick between 128 MiB and 4 GiB?  Whole probes (graph replay) with the read variant
forced, S = 256 MiB / 512 MiB / 1 GiB / 2 GiB.  JSON lines on stdout."""
import importlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == "child":
    cro = importlib.import_module("composable-resource-operator_b200")
    row = {}
    for mib in (256, 512, 1024, 2048):
        with cro.ProbeContext(sweep_bytes=mib << 20, devices=[0]) as c:
            for _ in range(5):
                c.probe_device(0)
            rs = [c.probe_device(0) for _ in range(30)]
            assert all(r.status == 0 and r.copy_verified == 5 for r in rs)
            row["probe_%d_us" % mib] = round(sorted(r.total_ns for r in rs)[len(rs) // 2] / 1e3, 1)
            row["read_best_%d_us" % mib] = round(min(r.read_best_ns for r in rs) / 1e3, 1)
            row["read_variant_%d" % mib] = int(rs[-1].read_variant)
    print(json.dumps(row))
    sys.exit(0)
import __graft_entry__ as g  # noqa: E402

g.build()
for env in ({}, {"CRO_READ_VARIANT": 2}, {"CRO_READ_VARIANT": 3}):
    e = dict(os.environ)
    e.update({k: str(v) for k, v in env.items()})
    p = subprocess.run([sys.executable, __file__, "child"], env=e, capture_output=True, text=True, timeout=150)
    print(json.dumps({"env": env, **(json.loads(p.stdout.strip().split("\n")[-1]) if p.returncode == 0 else {"error": (p.stdout + p.stderr)[-400:]})}), flush=True)
