"""Focused sweep 3: dynamic tile scheduling on/off for the TMA kernels, high wave counts for the LDG kernels."""
import importlib, itertools, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
g.build()
cro = importlib.import_module("composable-resource-operator_b200")
S = 4 << 30
ITERS, TRIALS = 10, 5

def run(tag, env, fn):
    for k, v in env.items():
        os.environ[k] = str(v)
    try:
        with cro.ProbeContext(sweep_bytes=S, devices=[0]) as c:
            c.hbm_fill(0); fn(c)
            vals = sorted(r.bytes / r.ns for r in (fn(c) for _ in range(TRIALS)))
            print(json.dumps({"tag": tag, "env": env, "best": round(vals[-1], 1), "median": round(vals[len(vals) // 2], 1)}), flush=True)
    except Exception as e:
        print(json.dumps({"tag": tag, "env": env, "error": str(e)}), flush=True)
    for k in env:
        os.environ.pop(k, None)

for dyn in (0, 1):
    for tile, stages, threads in ((32768, 4, 288), (32768, 4, 160), (16384, 8, 288), (65536, 2, 288), (32768, 6, 288), (65536, 3, 288), (16384, 4, 160), (32768, 3, 160), (16384, 12, 288)):
        run("read_tma", {"CRO_TMA_READ_DYN": dyn, "CRO_TMA_READ_TILE": tile, "CRO_TMA_READ_STAGES": stages, "CRO_TMA_READ_THREADS": threads},
            lambda c: c.hbm_read_checksum(0, cro.READ_TMA, ITERS))
for dyn in (0, 1):
    for tile, stages, waves in ((32768, 6, 1), (65536, 3, 1), (32768, 3, 1), (16384, 3, 1), (16384, 6, 1), (8192, 3, 1), (8192, 6, 1), (16384, 3, 8), (8192, 3, 8),
                                (32768, 2, 1), (16384, 2, 1), (16384, 4, 1), (32768, 4, 1), (8192, 4, 1), (4096, 4, 1), (8192, 3, 16), (8192, 3, 32)):
        if dyn and waves > 1:
            continue
        run("copy_tma", {"CRO_TMA_COPY_DYN": dyn, "CRO_TMA_COPY_TILE": tile, "CRO_TMA_COPY_STAGES": stages, "CRO_TMA_COPY_WAVES": waves},
            lambda c: c.hbm_copy(0, cro.COPY_TMA, ITERS))
for waves in (32, 64, 128, 256, 512):
    run("copy_ldg", {"CRO_COPY_WAVES": waves}, lambda c: c.hbm_copy(0, cro.COPY_LDG, ITERS))
for waves in (64, 128, 256, 512):
    run("fill", {"CRO_FILL_WAVES": waves}, lambda c: c.hbm_fill(0, ITERS))
