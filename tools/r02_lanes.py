"""Does a second probe queued behind a running one keep the GPU busy?  begin/end loops at depth 1 and 2 (1 GPU)."""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
cro = importlib.import_module("composable-resource-operator_b200")
S = 4 << 30
N = 40
out = {}
with cro.ProbeContext(sweep_bytes=S, devices=[0]) as c:
    for _ in range(3):
        c.probe_device(0)
    def sweeps():
        return [[t.kind, t.index, round(t.bytes / t.event_ns, 1), round(t.bytes / t.timer_ns, 1)] for t in c.sweep_times(0)]
    t0 = time.perf_counter()
    tot = []
    for _ in range(N):
        tot.append(c.probe_device(0).total_ns)
    out["sync_ms"] = (time.perf_counter() - t0) / N * 1e3
    out["sync_total_ms"] = sum(tot) / N / 1e6
    out["sync_sweeps"] = sweeps()
    # depth 1: begin, end, begin, end
    t0 = time.perf_counter()
    tot = []
    per_lane = {0: [], 1: []}
    for i in range(N):
        c.probe_begin(0)
        r = c.probe_end(0)
        tot.append(r.total_ns)
        per_lane[i & 1].append(r.total_ns)
    out["depth1_ms"] = (time.perf_counter() - t0) / N * 1e3
    out["depth1_total_ms"] = sum(tot) / N / 1e6
    out["depth1_total_by_parity_ms"] = [sum(v) / len(v) / 1e6 for v in per_lane.values()]
    out["depth1_sweeps"] = sweeps()
    # depth 2: two in flight at all times
    c.probe_begin(0); c.probe_begin(0); c.probe_end(0); c.probe_end(0)      # lane 1's graph is captured here
    c.probe_begin(0)
    c.probe_begin(0)
    t0 = time.perf_counter()
    rs = []
    for _ in range(N):
        rs.append(c.probe_end(0))
        c.probe_begin(0)
    out["depth2_ms"] = (time.perf_counter() - t0) / N * 1e3
    c.probe_end(0); c.probe_end(0)
    gaps = [(b.t_start_ns - (a.t_start_ns + a.total_ns)) / 1e3 for a, b in zip(rs, rs[1:])]
    out["depth2_device_gap_us"] = {"min": min(gaps), "median": sorted(gaps)[len(gaps) // 2], "max": max(gaps)}
    out["probe_total_ms"] = sum(r.total_ns for r in rs) / len(rs) / 1e6
    out["depth2_sweeps"] = sweeps()
    out["depth2_fill_copy_read_ms"] = [sum(r.fill_ns for r in rs) / len(rs) / 1e6, sum(r.copy_median_ns for r in rs) / len(rs) / 1e6, sum(r.read_median_ns for r in rs) / len(rs) / 1e6]
    out["ok"] = all(r.status == 0 for r in rs)
print(json.dumps(out))
