set -u
OUT=gpurun_out/r02r
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 20 --warmup 5 2> $OUT/bench_n4.err | tail -1 > $OUT/bench_n4.json; echo "bench n4 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02r/bench_n4.json").read())
f = d["fullbox"]
print("N=4 value %.1f e2e %.1f parity %s | fullbox %.3f ms (%s) %s probes/s | storm %s probes/s wall %s bound %s | churn %s" % (
    d["value"], d["e2e"]["value"], d["parity_ok"], f["ms_per_call"], f["phases_ms"], f["probes_per_s"],
    d["storm"]["child_probes_per_s"], d["storm"]["wall_s"], d["storm"]["bound_s_busiest_gpu"], d["churn"]["probes_per_s"]))
PY
