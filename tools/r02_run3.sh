set -u
OUT=gpurun_out/r02d
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -4 $OUT/gpu_tests.log
timeout 200 python tools/r02_exp2.py 0 > $OUT/exp2_dev0.json 2> $OUT/exp2.err; cat $OUT/exp2_dev0.json
timeout 200 python tools/r02_exp2.py 1 > $OUT/exp2_dev1.json 2>> $OUT/exp2.err; cat $OUT/exp2_dev1.json
for ctas in 1 2 4; do CRO_EXPECT_CTAS=$ctas timeout 100 python tools/r02_quick.py child probe 2>>$OUT/exp2.err | tail -1 | cut -c1-700; done
CRO_EXPECT_OVERLAP=0 timeout 100 python tools/r02_quick.py child probe 2>>$OUT/exp2.err | tail -1 | cut -c1-300
tail -5 $OUT/exp2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-storm 2> $OUT/bench_n2.err | tail -1 > $OUT/bench_n2.json; echo "bench n2 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02d/bench_n2.json").read())
print("N=2 value %.1f e2e %.1f (%.2f ms/step) parity %s" % (d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["parity_ok"]))
print({k: d["e2e"].get(k) for k in ("reconcile_ms", "allgather_wall_ms")})
PY
