set -u
OUT=gpurun_out/r02p
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi -L | wc -l
timeout 400 python -m pytest tests -m gpu -x -q -k "multi_device or odd_number or peer_push or probe_by_uuid or storm" 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 2> $OUT/bench_n8.err | tail -1 > $OUT/bench_n8.json; echo "bench n8 rc=$?"
tail -3 $OUT/bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 20 --warmup 5 2> $OUT/bench_n4.err | tail -1 > $OUT/bench_n4.json; echo "bench n4 rc=$?"
python - <<'PY'
import json
for n in (8, 4):
    try:
        d = json.loads(open("gpurun_out/r02p/bench_n%d.json" % n).read())
    except Exception as e:
        print(n, "no line", e); continue
    print("N=%d value %.1f e2e %.1f parity %s" % (n, d["value"], d["e2e"]["value"], d["parity_ok"]))
    f = d.get("fullbox", {})
    print(" fullbox", {k: f.get(k) for k in ("probes_per_s", "ms_per_call", "phases_ms", "first_call_s", "nvlink_read_gbs", "nvlink_push_gbs", "latency_ns", "matrix_flat", "parity_ok", "allgather_us", "error")})
    print(" latency_vs_hops", f.get("latency_vs_hops"))
    for k in ("storm", "churn"):
        if k in d:
            print(" ", k, {x: d[k][x] for x in d[k] if x not in ("note", "gpus")})
PY
