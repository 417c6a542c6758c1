"""Generates tests/golden/pattern_kats.json from the Python oracle.

Run from the repo root:  python tests/golden/make_pattern_kats.py
The vectors pin the probe pattern (splitmix64 step of seed+i) and its
(XOR, wrapping-sum) checksum at sizes small enough to recompute anywhere.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle  # noqa: E402

SEED_BASE = 0x00C0FFEE00000000
out = {"_about": "probe pattern golden vectors; made by tests/golden/make_pattern_kats.py (pure-Python oracle)",
       "words": [], "checksums": []}
for seed in (SEED_BASE | 0, SEED_BASE | 3, 0, (1 << 64) - 1):
    for i in (0, 1, 2, 255, (1 << 32) - 1, 1 << 32, (1 << 64) - 1):
        out["words"].append({"seed": "%#018x" % seed, "i": str(i), "word": "%#018x" % oracle.pattern_word(seed, i)})
for seed in (SEED_BASE | 0, SEED_BASE | 7):
    for first, n in ((0, 1), (0, 2), (0, 1000), (5, 4096), (0, 1 << 16), (123456789, 100003), (0, 1 << 20)):
        x, s = 0, 0
        if n <= 4096:   # pure-Python loop for the small ones, numpy beyond
            for k in range(n):
                w = oracle.pattern_word(seed, first + k)
                x ^= w
                s = (s + w) & oracle.MASK
        else:
            x, s = oracle.checksum_np(seed, first, n)
        out["checksums"].append({"seed": "%#018x" % seed, "first": first, "n_words": n,
                                 "xor": "%#018x" % x, "sum": "%#018x" % s})
with open(os.path.join(ROOT, "tests", "golden", "pattern_kats.json"), "w") as f:
    json.dump(out, f, indent=1)
print("wrote", len(out["words"]), "words,", len(out["checksums"]), "checksums")
