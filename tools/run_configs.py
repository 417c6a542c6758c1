"""Runs BASELINE.json configs 3, 4 and 5 on the box's GPUs (single process, as a Go operator would) and prints
one JSON object per config.  Config 2 is bench.py; config 1 is `bench.py --impl reference`.

  config 3  8xB200 full-box compose: cro_probe_all (concurrent probes + NVLink rounds + one NCCL all-gather)
  config 4  reconcile storm: 1000 synthetic ComposabilityRequests over the box's GPUs, warm probe contexts
  config 5  attach/detach churn: 100 cycles x 4-GPU compose / decompose, probe each attach

Usage: python tools/run_configs.py [--sweep-gib 4] [--storm 1000] [--cycles 100] [--configs 3,4,5]
"""
import argparse
import importlib
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import __graft_entry__ as g  # noqa: E402

g.build()
cro = importlib.import_module("composable-resource-operator_b200")

ap = argparse.ArgumentParser()
ap.add_argument("--sweep-gib", type=float, default=4.0)
ap.add_argument("--storm", type=int, default=1000)
ap.add_argument("--cycles", type=int, default=100)
ap.add_argument("--configs", default="3,4,5")
ap.add_argument("--cpu-compare", action="store_true", help="also run storm/churn with the probe off (reference-strength check)")
args = ap.parse_args()
S = int(args.sweep_gib * (1 << 30))
which = set(args.configs.split(","))


def gbs(b, ns):
    return round(b / ns, 1) if ns else None


def config3(ctx):
    t0 = time.perf_counter()
    res = ctx.probe_all()            # first call pays ncclCommInitAll + peer enablement
    cold = time.perf_counter() - t0
    times = []
    for _ in range(5):
        t0 = time.perf_counter()
        res = ctx.probe_all()
        times.append(time.perf_counter() - t0)
    n = len(res)
    import oracle
    co = oracle.COracle()
    parity = all(r.checksum == co.checksum(r.seed, 0, S // 8, threads=os.cpu_count() or 1) for r in res[:2])
    return {
        "config": 3, "n_gpus": n, "sweep_bytes": S, "p2p_bytes": int(res[0].p2p_bytes), "cold_s": round(cold, 3),
        "warm_s_median": round(sorted(times)[len(times) // 2], 4), "probes_per_s": round(n / sorted(times)[len(times) // 2], 1),
        "status": [r.status for r in res], "parity_first_two_vs_oracle": parity,
        "hbm_read_gbs": [gbs(S, r.read_best_ns) for r in res], "hbm_copy_gbs": [gbs(2 * S, r.copy_best_ns) for r in res],
        "hbm_fill_gbs": [gbs(S, r.fill_ns) for r in res],
        "nvlink_read_gbs": [[gbs(r.p2p_bytes, r.p2p_read_ns[j]) if j < 8 and r.p2p_read_ns[j] else None for j in range(n)] for r in res],
        "nvlink_write_gbs": [[gbs(r.p2p_bytes, r.p2p_write_ns[j]) if j < 8 and r.p2p_write_ns[j] else None for j in range(n)] for r in res],
        "nvlink_latency_ns": [[r.p2p_latency_ns_x16[j] // 16 if j < 8 and r.p2p_read_ns[j] else None for j in range(n)] for r in res],
    }


def storm(ctx, n_req, probe):
    n = ctx.device_count()
    nodes = ["worker-%d" % i for i in range(n)]
    rng = random.Random(20260921)
    with cro.Cluster({"nodes": nodes, "probe": probe, "device_resource_type": "DEVICE_PLUGIN"}, ctx) as c:
        sizes = []
        for i in range(n_req):
            size = rng.randint(1, 4)
            sizes.append(size)
            err = c.apply("req-%04d" % i, {"type": "gpu", "model": "NVIDIA-B200-%d" % (i // n), "size": size,
                                           "allocation_policy": "samenode", "target_node": "worker-%d" % (i % n)})
            assert err == "", err
        t0 = time.perf_counter()
        st = c.run()
        wall = time.perf_counter() - t0
        return {"config": 4, "probe": probe, "n_gpus": n, "requests": n_req, "children": sum(sizes), "wall_s": round(wall, 3),
                "requests_running": st["requests_running"], "resources_online": st["resources_online"],
                "requests_per_s": round(st["requests_running"] / wall, 1), "child_probes_per_s": round(st["probes"] / wall, 1) if probe else None,
                "specs_per_s": round(st["status_updates"] / wall, 1), "status_updates": st["status_updates"], "spec_bytes": st["spec_bytes"],
                "reconciles": st["request_reconciles"] + st["resource_reconciles"], "reconcile_p50_us": st["reconcile_p50_ns"] / 1e3,
                "reconcile_p99_us": st["reconcile_p99_ns"] / 1e3, "errors": st["reconcile_errors"], "probe_failures": st["probe_failures"],
                "note": "single reconcile worker per controller (reference default), physical GPUs multiplexed across CRs, timers immediate"}


def churn(ctx, cycles, probe):
    n = ctx.device_count()
    width = min(4, n)
    with cro.Cluster({"nodes": ["worker-%d" % i for i in range(n)], "probe": probe}, ctx) as c:
        t0 = time.perf_counter()
        probes = 0
        for cyc in range(cycles):
            names = []
            for j in range(width):
                name = "churn-%d-%d" % (cyc, j)
                names.append(name)
                assert c.apply(name, {"type": "gpu", "model": "NVIDIA-B200", "size": 1, "target_node": "worker-%d" % ((width * cyc + j) % n)}) == ""
            st = c.run()
            d = c.dump()
            assert all(d["requests"][x]["status"]["state"] == "Running" for x in names), d["requests"]
            for x in names:
                c.delete(x)
            st = c.run()
            probes = st["probes"]
        wall = time.perf_counter() - t0
        d = c.dump()
        return {"config": 5, "probe": probe, "n_gpus": n, "cycles": cycles, "width": width, "attaches": cycles * width, "wall_s": round(wall, 3),
                "probes": probes, "probes_per_s": round(probes / wall, 1) if probe else None, "attach_detach_cycles_per_s": round(cycles / wall, 2),
                "left_over_objects": len(d["requests"]) + len(d["resources"]), "errors": st["reconcile_errors"], "probe_failures": st["probe_failures"],
                "note": "logical attach/detach (CUDA cannot hot-plug inside one process); warm probe contexts"}


t0 = time.perf_counter()
ctx = cro.ProbeContext(sweep_bytes=S)
init_s = time.perf_counter() - t0
print(json.dumps({"init": {"devices": ctx.device_count(), "cold_init_s": round(init_s, 3), "sweep_bytes": S}}), flush=True)
if "3" in which:
    print(json.dumps(config3(ctx)), flush=True)
if "4" in which:
    print(json.dumps(storm(ctx, args.storm, True)), flush=True)
    if args.cpu_compare:
        print(json.dumps(storm(ctx, args.storm, False)), flush=True)
if "5" in which:
    print(json.dumps(churn(ctx, args.cycles, True)), flush=True)
    if args.cpu_compare:
        print(json.dumps(churn(ctx, args.cycles, False)), flush=True)
ctx.close()
