set -u
OUT=gpurun_out/r02f
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -6 $OUT/gpu_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --storm 400 --cycles 40 2> $OUT/bench_n2.err | tail -1 > $OUT/bench_n2.json; echo "bench n2 rc=$?"
tail -3 $OUT/bench_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02f/bench_n2.json").read())
print("N=2 value %.1f e2e %.1f parity %s" % (d["value"], d["e2e"]["value"], d["parity_ok"]))
f = d["fullbox"]
print("fullbox", f["probes_per_s"], f["ms_per_call"], f["phases_ms"], f["latency_hops"], f["first_call_s"])
print("storm", d["storm"]["child_probes_per_s"], d["storm"]["wall_s"], d["storm"]["children"], d["storm"]["busiest_gpu_children"], "churn", d["churn"]["probes_per_s"])
PY
