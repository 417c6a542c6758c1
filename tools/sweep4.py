"""Sweep 4: chunked tile claims (fewer atomics) for the dynamic TMA kernels; thread counts."""
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

for tile, stages, threads, chunk in ((32768, 4, 160, 1), (32768, 4, 160, 2), (32768, 4, 160, 4), (16384, 8, 160, 2), (16384, 8, 160, 4), (16384, 8, 288, 8),
                                     (8192, 16, 160, 8), (32768, 4, 96, 1), (32768, 4, 64, 1), (32768, 5, 160, 1), (32768, 6, 160, 2), (65536, 3, 160, 1),
                                     (65536, 3, 288, 1), (49152, 4, 160, 1), (24576, 6, 160, 2), (32768, 4, 288, 2)):
    run("read_tma", {"CRO_TMA_READ_TILE": tile, "CRO_TMA_READ_STAGES": stages, "CRO_TMA_READ_THREADS": threads, "CRO_TMA_READ_CHUNK": chunk},
        lambda c: c.hbm_read_checksum(0, cro.READ_TMA, ITERS))
for tile, stages, chunk in ((32768, 4, 1), (32768, 4, 2), (32768, 2, 1), (32768, 2, 2), (16384, 4, 2), (16384, 4, 4), (8192, 8, 4), (8192, 8, 8), (65536, 3, 1), (65536, 2, 1),
                            (49152, 4, 1), (32768, 6, 2), (16384, 8, 2)):
    run("copy_tma", {"CRO_TMA_COPY_TILE": tile, "CRO_TMA_COPY_STAGES": stages, "CRO_TMA_COPY_CHUNK": chunk},
        lambda c: c.hbm_copy(0, cro.COPY_TMA, ITERS))
