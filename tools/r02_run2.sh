set -u
OUT=gpurun_out/r02c
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi -L > $OUT/gpus.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -30 $OUT/gpu_tests.log
timeout 300 python bench.py --steps 5 --warmup 3 2> $OUT/bench_n1.err | tail -1 > $OUT/bench_n1.json; echo "bench n1 rc=$?"
tail -5 $OUT/bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --storm 200 --cycles 20 2> $OUT/bench_n2.err | tail -1 > $OUT/bench_n2.json; echo "bench n2 rc=$?"
tail -15 $OUT/bench_n2.err
python - <<'PY'
import json
for n in (1, 2):
    try:
        d = json.loads(open("gpurun_out/r02c/bench_n%d.json" % n).read())
    except Exception as e:
        print("n", n, "no line", e); continue
    print("N=%d value %.1f e2e %.1f parity %s roofline %s %.1f" % (n, d["value"], d["e2e"]["value"], d["parity_ok"], d["roofline"]["kernel"], d["roofline"]["achieved"]))
    for k in ("cold", "fullbox", "storm", "churn"):
        if k in d:
            print(k, json.dumps(d[k])[:1500])
PY
