"""bench.py's output contract, on what can run without a GPU: the reference arm (`--impl reference`, the oracle port timed
on host cores) prints exactly ONE JSON line with the keys the driver reads — alone and under torchrun, where rank 0
alone runs it — and the product arm fails loudly without a GPU instead of falling back to anything."""
import json
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = json.load(open(os.path.join(ROOT, "BASELINE.json")))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _check_reference_line(out, n, steps, warmup):
    lines = [ln for ln in out.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == n and d["steps"] == steps and d["warmup"] == warmup
    for k in ("metric", "value", "unit", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
              "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0 and d["ms_per_step"] > 0
    assert isinstance(d["config"].get("workload"), str) and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    # the metric is BASELINE.json's
    want = BASE.get("metric") if isinstance(BASE.get("metric"), str) else (BASE.get("metric") or {}).get("name")
    if want:
        assert d["metric"] == want or want in d["metric"] or d["metric"] in want, (d["metric"], want)
    return d


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "2", "--warmup", "1"], cwd=ROOT,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-1500:]
    _check_reference_line(r.stdout, 1, 2, 1)


def test_reference_arm_under_torchrun_only_rank0_prints():
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(_free_port()), "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       cwd=ROOT, capture_output=True, text=True, timeout=400)
    assert r.returncode == 0, r.stderr[-1500:]
    _check_reference_line(r.stdout, 2, 1, 1)


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present: the product arm runs (covered by the driver's own bench run)")
    r = subprocess.run([sys.executable, "bench.py", "--steps", "1", "--warmup", "1"], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode != 0
    assert not [ln for ln in r.stdout.splitlines() if ln.strip().startswith("{")], r.stdout[-500:]    # no number without a GPU
