"""Sweep 5: L2 cache-hint policies on the TMA bulk copies (CRO_TMA_READ_HINT / CRO_TMA_COPY_HINT).
hint bits — copy: 1 evict_first loads, 2 evict_first stores, 4 evict_last stores; read: 1 evict_first loads."""
import importlib, json, os, sys
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


for hint in (0, 1):
    run("read_tma", {"CRO_TMA_READ_HINT": hint}, lambda c: c.hbm_read_checksum(0, cro.READ_TMA, ITERS))
for hint in (0, 1, 2, 3, 4, 5):
    for tile, stages in ((32768, 4), (65536, 3)):
        run("copy_tma", {"CRO_TMA_COPY_HINT": hint, "CRO_TMA_COPY_TILE": tile, "CRO_TMA_COPY_STAGES": stages},
            lambda c: c.hbm_copy(0, cro.COPY_TMA, ITERS))
