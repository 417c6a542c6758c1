"""Generates tests/golden/pattern_kats.json from the Python oracle.

Run from the repo root:  python tests/golden/make_pattern_kats.py
The vectors pin the probe pattern (splitmix64 step of seed+i), its
(XOR, wrapping-sum, position-weighted-sum) checksum at sizes small enough to
recompute anywhere, the per-probe seed schedule and the end points of the
NVLink latency permutations (Sattolo cycle, mt19937_64) — everything with the
pure-Python loop, independent of numpy and of the C oracle.
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
        x, s, ws = 0, 0, 0
        if n <= 4096:   # pure-Python loop for the small ones, numpy beyond
            for k in range(n):
                w = oracle.pattern_word(seed, first + k)
                x ^= w
                s = (s + w) & oracle.MASK
                ws = (ws + w * (2 * (first + k) + 1)) & oracle.MASK      # word at position first + k
        else:
            x, s, ws = oracle.checksum_np(seed, first, n)
        out["checksums"].append({"seed": "%#018x" % seed, "first": first, "n_words": n,
                                 "xor": "%#018x" % x, "sum": "%#018x" % s, "wsum": "%#018x" % ws})
# seed of probe number `nonce` on the device with that minor (nonce 0 is SURVEY.md section 8d's seed)
out["probe_seeds"] = [{"seed_base": "%#018x" % SEED_BASE, "minor": m, "nonce": k,
                       "seed": "%#018x" % (((SEED_BASE | m) + k * 0xD1B54A32D192ED03) & oracle.MASK)}
                      for m in (0, 7) for k in (0, 1, 2, 1000, (1 << 32) - 1)]
# latency permutation end points: std::mt19937_64 restated here (Matsumoto & Nishimura 2004)
def mt64(seed):
    mt = [0] * 312
    mt[0] = seed & oracle.MASK
    for i in range(1, 312):
        mt[i] = (6364136223846793005 * (mt[i - 1] ^ (mt[i - 1] >> 62)) + i) & oracle.MASK
    idx = 312
    while True:
        if idx >= 312:
            for i in range(312):
                x = (mt[i] & 0xFFFFFFFF80000000) | (mt[(i + 1) % 312] & 0x7FFFFFFF)
                mt[i] = mt[(i + 156) % 312] ^ (x >> 1) ^ (0xB5026F5AA96619E9 if x & 1 else 0)
            idx = 0
        y = mt[idx]
        idx += 1
        y ^= (y >> 29) & 0x5555555555555555
        y ^= (y << 17) & 0x71D67FFFEDA60000
        y ^= (y << 37) & 0xFFF7EEE000000000
        y ^= y >> 43
        yield y & oracle.MASK


def chase_end(src, dst, hops, n=65536):
    g = mt64(src * 8 + dst)
    perm = list(range(n))
    for i in range(n - 1, 0, -1):
        j = next(g) % i
        perm[i], perm[j] = perm[j], perm[i]
    at = 0
    for _ in range(hops):
        at = perm[at]
    return at


out["chase_ends"] = [{"minor_src": a, "minor_dst": b, "hops": h, "end": chase_end(a, b, h)}
                     for a, b in ((0, 1), (1, 0), (3, 7), (7, 0)) for h in (1, 2, 4096, 16384, 65535, 65536)]
# the first draw of mt19937_64 seeded with 5489 is the C++ standard's own check value (10000th draw = 9981545732273789042)
g = mt64(5489)
for _ in range(9999):
    next(g)
out["mt19937_64_10000th_of_default_seed"] = str(next(g))
with open(os.path.join(ROOT, "tests", "golden", "pattern_kats.json"), "w") as f:
    json.dump(out, f, indent=1)
print("wrote", len(out["words"]), "words,", len(out["checksums"]), "checksums")
