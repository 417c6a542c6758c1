"""Storm (BASELINE config 4) with the reconcile driver's event trace: where does the one long idle gap per GPU come from?"""
import importlib, json, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
cro = importlib.import_module("composable-resource-operator_b200")
n_req = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
with cro.ProbeContext(sweep_bytes=4 << 30) as ctx:
    n = ctx.device_count()
    rng = random.Random(20260921)
    with cro.Cluster({"nodes": ["worker-%d" % i for i in range(n)], "probe": True, "trace": True}, ctx) as c:
        for i in range(n_req):
            assert c.apply("req-%04d" % i, {"type": "gpu", "model": "NVIDIA-B200-%d" % (i // n), "size": rng.randint(1, 4),
                                            "allocation_policy": "samenode", "target_node": "worker-%d" % (i % n)}) == ""
        t0 = time.perf_counter()
        st = c.run()
        wall = time.perf_counter() - t0
tr = st.pop("trace")
print(json.dumps({k: st[k] for k in st if k != "gpus"}), "wall", round(wall, 3))
print(json.dumps(st["gpus"]))
# the largest silence between consecutive events of device 0, and what surrounds it
ev0 = [e for e in tr if e[2] == 0]
gaps = sorted(((b[0] - a[0], i) for i, (a, b) in enumerate(zip(ev0, ev0[1:]))), reverse=True)[:3]
for g, i in gaps:
    print("dev0 silence %.1f ms between" % (g / 1e3), ev0[max(0, i - 3):i + 4])
allgaps = sorted(((b[0] - a[0], i) for i, (a, b) in enumerate(zip(tr, tr[1:]))), reverse=True)[:3]
for g, i in allgaps:
    print("ANY-device silence %.1f ms between" % (g / 1e3), tr[max(0, i - 2):i + 3])
print("first 40 events", tr[:40])
