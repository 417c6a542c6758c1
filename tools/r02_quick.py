"""Round-2 first look on one B200: the checksumming copy against the plain TMA copy, the closed-form generator beside
the copy sweeps vs in line, %globaltimer windows against CUDA events.  Prints JSON lines (gpurun_out/r02_quick.jsonl)."""
import importlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def child(env, what):
    e = dict(os.environ)
    e.update({k: str(v) for k, v in env.items()})
    out = subprocess.run([sys.executable, __file__, "child", what], env=e, capture_output=True, text=True, timeout=240)
    if out.returncode != 0:
        return {"what": what, "env": env, "error": (out.stdout + out.stderr)[-800:]}
    d = json.loads(out.stdout.strip().split("\n")[-1])
    d.update({"what": what, "env": env})
    return d


def run_child(what):
    cro = importlib.import_module("composable-resource-operator_b200")
    S = int(os.environ.get("S_GIB", "4")) << 30
    iters = 10
    with cro.ProbeContext(sweep_bytes=S, devices=[0]) as c:
        if what == "copy":
            out = {}
            for v, name in ((cro.COPY_TMA, "tma"), (cro.COPY_TMA_FUSED, "fused"), (cro.COPY_LDG, "ldg")):
                c.hbm_copy(0, v, 2)
                r = c.hbm_copy(0, v, iters)
                out[name + "_gbs"] = round(r.bytes / r.ns, 1)
            for v, name in ((cro.READ_TMA, "read_tma"), (cro.READ_LDG, "read_ldg"), (cro.READ_LDG256, "read_ldg256")):
                c.hbm_read_checksum(0, v, 2)
                r = c.hbm_read_checksum(0, v, iters)
                out[name + "_gbs"] = round(r.bytes / r.ns, 1)
            c.hbm_fill(0, 2)
            r = c.hbm_fill(0, iters)
            out["fill_gbs"] = round(r.bytes / r.ns, 1)
            print(json.dumps(out))
        elif what == "probe":
            for _ in range(3):
                c.probe_device(0)
            tot_t = tot_e = 0
            rows = None
            for _ in range(10):
                r = c.probe_device(0)
                assert r.status == 0, (r.status, r.fail_code, r.fail_index)
                ts = c.sweep_times(0)
                tot_t += r.total_ns
                tot_e += sum(t.event_ns for t in ts)
                rows = ts
            print(json.dumps({"probe_ms_timer": tot_t / 10 / 1e6, "probe_ms_events": tot_e / 10 / 1e6,
                              "sweeps": [[t.kind, t.index, round(t.bytes / t.event_ns, 1), round(t.bytes / max(1, t.timer_ns), 1)] for t in rows],
                              "copy_verified": r.copy_verified}))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        run_child(sys.argv[2])
        sys.exit(0)
    rows = [child({}, "copy"), child({}, "probe"), child({"CRO_EXPECT_OVERLAP": 0}, "probe"),
            child({"CRO_FUSED_THREADS": 96}, "copy"), child({"CRO_FUSED_THREADS": 288}, "copy"),
            child({"CRO_FUSED_STAGES": 6, "CRO_FUSED_TILE": 32768}, "copy"), child({"CRO_FUSED_STAGES": 3, "CRO_FUSED_TILE": 65536}, "copy"),
            child({"CRO_FUSED_STAGES": 8, "CRO_FUSED_TILE": 16384, "CRO_FUSED_CHUNK": 2}, "copy"),
            child({"S_GIB": 1}, "probe")]
    for r in rows:
        print(json.dumps(r))
