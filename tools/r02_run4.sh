set -u
OUT=gpurun_out/r02o
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -6 $OUT/gpu_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 2> $OUT/bench_n1.err | tail -1 > $OUT/bench_n1.json; echo "bench n1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 2> $OUT/bench_n2.err | tail -1 > $OUT/bench_n2.json; echo "bench n2 rc=$?"
timeout 200 python bench.py --impl reference --steps 20 --warmup 5 2> $OUT/bench_ref.err | tail -1 > $OUT/bench_ref.json; echo "ref rc=$?"
python - <<'PY'
import json
for n in (1, 2):
    d = json.loads(open("gpurun_out/r02o/bench_n%d.json" % n).read())
    print("N=%d value %.1f e2e %.1f (%.2f ms; reconcile %.2f, allgather %.2f) parity %s" % (n, d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["reconcile_ms"], d["e2e"]["allgather_wall_ms"], d["parity_ok"]))
    for k in d["roofline_kernels"]:
        print("   ", k["kernel"], round(k["achieved"], 1), k.get("avg_launch_ms"), k.get("share_of_step"))
    for k in ("cold", "fullbox", "storm", "churn"):
        if k in d:
            print(k, json.dumps(d[k])[:900])
r = json.loads(open("gpurun_out/r02o/bench_ref.json").read())
print("ref", r["value"], r["cpu_best_case"]["value"], r["config"] == json.loads(open("gpurun_out/r02o/bench_n1.json").read())["config"])
PY
