"""Where the constant ~12 us per launch of the TMA sweeps goes: per-launch time of the read and the checksumming copy at
16 MiB / 256 MiB / 4 GiB under different ring shapes (the validated CRO_* knobs).  t(S) = S / R_inf + c, so the 16 MiB
row is nearly all c.  JSON lines on stdout."""
import importlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SETTINGS = [
    {},
    {"CRO_TMA_READ_DYN": 0, "CRO_TMA_COPY_DYN": 0},
    {"CRO_TMA_READ_STAGES": 2, "CRO_FUSED_STAGES": 2},
    {"CRO_TMA_READ_STAGES": 8, "CRO_TMA_READ_TILE": 16384, "CRO_FUSED_STAGES": 8, "CRO_FUSED_TILE": 16384},
    {"CRO_TMA_READ_TILE": 16384, "CRO_FUSED_TILE": 16384},
    {"CRO_TMA_READ_TILE": 57344, "CRO_FUSED_TILE": 57344},
    {"CRO_TMA_READ_THREADS": 96, "CRO_FUSED_THREADS": 96},
    {"CRO_TMA_READ_THREADS": 288, "CRO_FUSED_THREADS": 288},
    {"CRO_TMA_READ_CHUNK": 2, "CRO_FUSED_CHUNK": 2},
    {"CRO_TMA_READ_WAVES": 2, "CRO_TMA_READ_TILE": 16384},
]


def run_child():
    cro = importlib.import_module("composable-resource-operator_b200")
    row = {}
    for mib, iters in ((16, 200), (256, 60), (4096, 8)):
        with cro.ProbeContext(sweep_bytes=mib << 20, devices=[0]) as c:
            c.hbm_fill(0)
            for name, fn in (("read_tma", lambda n: c.hbm_read_checksum(0, cro.READ_TMA, n)), ("read_ldg256", lambda n: c.hbm_read_checksum(0, cro.READ_LDG256, n)),
                             ("copy_fused", lambda n: c.hbm_copy(0, cro.COPY_TMA_FUSED, n)), ("fill", lambda n: c.hbm_fill(0, n))):
                fn(3)
                runs = [fn(iters) for _ in range(3)]
                best = min(r.ns / r.launches for r in runs)       # CUDA events around `iters` back-to-back launches
                row["%s_%d_us" % (name, mib)] = round(best / 1e3, 2)
                row["%s_%d_window_us" % (name, mib)] = round(min(r.timer_ns for r in runs) / 1e3, 2)   # the last launch's own %globaltimer window
    print(json.dumps(row))


if len(sys.argv) > 1 and sys.argv[1] == "child":
    run_child()
    sys.exit(0)
import __graft_entry__ as g  # noqa: E402

g.build()
for env in SETTINGS:
    e = dict(os.environ)
    e.update({k: str(v) for k, v in env.items()})
    p = subprocess.run([sys.executable, __file__, "child"], env=e, capture_output=True, text=True, timeout=120)
    if p.returncode != 0:
        print(json.dumps({"env": env, "error": (p.stdout + p.stderr)[-400:]}), flush=True)
        continue
    d = json.loads(p.stdout.strip().split("\n")[-1])
    # fixed cost: what is left of the 16 MiB launch after its bytes at the 4 GiB rate
    for k in ("read_tma", "read_ldg256", "copy_fused", "fill"):
        d[k + "_fixed_us"] = round(d[k + "_16_us"] - d[k + "_4096_us"] / 256.0, 2)
        d[k + "_fixed_in_window_us"] = round(d[k + "_16_window_us"] - d[k + "_4096_window_us"] / 256.0, 2)
        mul = 2 if k == "copy_fused" else 1
        d[k + "_4g_gbs"] = round(mul * (4096 << 20) / (d[k + "_4096_us"] * 1e3), 1)
    print(json.dumps({"env": env, **d}), flush=True)
