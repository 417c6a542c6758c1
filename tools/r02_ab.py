"""A/B of one knob on the box's GPUs: probe time alone, two probes in flight, and the full-box HBM phase."""
import importlib, json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    cro = importlib.import_module("composable-resource-operator_b200")
    out = {}
    with cro.ProbeContext(sweep_bytes=4 << 30) as c:
        n = c.device_count()
        for _ in range(3):
            c.probe_device(0)
        out["alone_ms"] = round(sum(c.probe_device(0).total_ns for _ in range(10)) / 10 / 1e6, 4)
        c.probe_begin(0); c.probe_begin(0); c.probe_end(0); c.probe_end(0)
        c.probe_begin(0); c.probe_begin(0)
        rs = []
        for _ in range(20):
            rs.append(c.probe_end(0))
            c.probe_begin(0)
        c.probe_end(0); c.probe_end(0)
        out["two_in_flight_ms"] = round(sum(r.total_ns for r in rs) / len(rs) / 1e6, 4)
        out["fill_ms_two_in_flight"] = round(sum(r.fill_ns for r in rs) / len(rs) / 1e6, 4)
        if n > 1:
            for _ in range(3):
                c.probe_all()
            hb, wall = [], []
            for _ in range(8):
                t0 = time.perf_counter()
                res = c.probe_all()
                wall.append(time.perf_counter() - t0)
                hb.append(c.fullbox_times().hbm_ns)
                assert all(r.status == 0 for r in res)
            out["fullbox_hbm_ms"] = round(sorted(hb)[4] / 1e6, 4)
            out["fullbox_ms"] = round(sorted(wall)[4] * 1e3, 3)
            out["fullbox_fill_ms"] = round(sum(r.fill_ns for r in res) / len(res) / 1e6, 4)
    print(json.dumps(out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
        sys.exit(0)
    knob = sys.argv[1] if len(sys.argv) > 1 else "CRO_CARVEOUT_FILL"
    for rep in range(2):
        for v in ("0", "1"):
            e = dict(os.environ); e[knob] = v
            p = subprocess.run([sys.executable, __file__, "child"], env=e, capture_output=True, text=True, timeout=300)
            print(knob, v, p.stdout.strip().split("\n")[-1] if p.returncode == 0 else (p.stdout + p.stderr)[-400:], flush=True)
