"""Times sweep-kernel variants / tuning points on cuda:0 (CUDA events inside libcroprobe).

Usage: python tools/sweep_variants.py [sweep_GiB] [group ...]   -> JSON lines on stdout.
Each point: 5 trials of 10 back-to-back launches; reports best and median GB/s
(algorithmic bytes / event time).  Tuning knobs are the CRO_* environment
variables csrc/kernels.cu reads at plan / launch time."""
import importlib
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

g.build()
cro = importlib.import_module("composable-resource-operator_b200")
S = int(float(sys.argv[1]) * (1 << 30)) if len(sys.argv) > 1 else 4 << 30
GROUPS = set(sys.argv[2:]) or {"read_ldg", "read_tma", "copy_ldg", "copy_tma", "fill"}
ITERS, TRIALS = 10, 5


def run(tag, env, fn):
    for k, v in env.items():
        os.environ[k] = str(v)
    try:
        with cro.ProbeContext(sweep_bytes=S, devices=[0]) as c:
            c.hbm_fill(0)
            fn(c)                     # warm
            vals = sorted(r.bytes / r.ns for r in (fn(c) for _ in range(TRIALS)))
            print(json.dumps({"tag": tag, "env": env, "best": round(vals[-1], 1), "median": round(vals[len(vals) // 2], 1)}), flush=True)
    except Exception as e:  # keep sweeping
        print(json.dumps({"tag": tag, "env": env, "error": str(e)}), flush=True)
    for k in env:
        os.environ.pop(k, None)


if "read_ldg" in GROUPS:
    for waves in (1, 2, 4, 8):
        run("read_ldg128", {"CRO_READ_WAVES": waves}, lambda c: c.hbm_read_checksum(0, cro.READ_LDG, ITERS))
        run("read_ldg256", {"CRO_READ_WAVES": waves}, lambda c: c.hbm_read_checksum(0, cro.READ_LDG256, ITERS))
if "read_tma" in GROUPS:
    for tile, stages, threads, waves in itertools.chain(
            itertools.product((16384, 32768, 65536), (2, 3, 4, 6), (160, 288), (1,)),
            [(8192, 8, 288, 1), (8192, 16, 288, 1), (16384, 8, 288, 1), (32768, 4, 288, 2), (32768, 4, 288, 4), (16384, 4, 288, 4),
             (32768, 2, 160, 1), (32768, 3, 96, 1), (16384, 3, 96, 1)]):
        if tile * stages > 220 * 1024:
            continue
        run("read_tma", {"CRO_TMA_READ_TILE": tile, "CRO_TMA_READ_STAGES": stages, "CRO_TMA_READ_THREADS": threads,
                         "CRO_TMA_READ_WAVES": waves}, lambda c: c.hbm_read_checksum(0, cro.READ_TMA, ITERS))
if "copy_ldg" in GROUPS:
    for waves in (1, 2, 4, 8, 16, 32):
        run("copy_ldg", {"CRO_COPY_WAVES": waves}, lambda c: c.hbm_copy(0, cro.COPY_LDG, ITERS))
if "copy_tma" in GROUPS:
    for tile, stages, waves in itertools.chain(
            itertools.product((16384, 32768, 65536), (2, 3, 4, 6), (1,)),
            [(8192, 4, 1), (8192, 8, 1), (16384, 2, 4), (32768, 2, 4), (32768, 2, 2), (65536, 3, 2), (65536, 3, 4), (32768, 6, 2), (32768, 6, 4),
             (16384, 3, 8), (8192, 3, 8), (4096, 4, 8)]):
        if tile * stages > 220 * 1024:
            continue
        run("copy_tma", {"CRO_TMA_COPY_TILE": tile, "CRO_TMA_COPY_STAGES": stages, "CRO_TMA_COPY_WAVES": waves},
            lambda c: c.hbm_copy(0, cro.COPY_TMA, ITERS))
if "fill" in GROUPS:
    for waves in (1, 2, 4, 8, 16, 32, 64):
        run("fill", {"CRO_FILL_WAVES": waves}, lambda c: c.hbm_fill(0, ITERS))
