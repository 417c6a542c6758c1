#!/bin/bash
# One gpurun call that re-measures everything a round's profiles/ entry needs (about 2.5 GPU-minutes on one B200):
#
#   gpurun --timeout 420 -- 'bash tools/gpu_round.sh r02'
#
# writes gpurun_out/<tag>/: gpu_tests.log, smoke.log, bench.json, bench_reference.json, launches.csv (ncu per-launch
# durations of `bench.py --steps 2 --warmup 1`), full.ncu-rep (ncu --set full of fill / read / copy at S = 1 GiB).
# Back in the container:  python tools/ncu_summary.py gpurun_out/<tag>/full.ncu-rep profiles/<tag>_ncu_full
# and copy the JSON lines / launches.csv into profiles/ (README.md there indexes them).
# Numbers printed by a run under ncu are never bench values: bench.json comes from the un-profiled run.
set -u
TAG=${1:-rXX}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests -m gpu -x -q > "$OUT/gpu_tests.log" 2>&1; echo "gpu tests rc=$?" | tee -a "$OUT/gpu_tests.log"
timeout 60 python -c 'import __graft_entry__ as g; g.smoke()' > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$OUT/smoke.log"
timeout 240 python bench.py 2> "$OUT/bench.err" | tail -1 > "$OUT/bench.json"; echo "bench rc=$?"
timeout 150 python bench.py --impl reference 2> "$OUT/bench_reference.err" | tail -1 > "$OUT/bench_reference.json"; echo "reference arm rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" \
    python bench.py --steps 2 --warmup 1 --no-cold --no-cpu-baseline > "$OUT/bench_under_ncu.log" 2>&1; echo "launch list rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k "regex:hbm_|finalize" -f -o "$OUT/full" \
    python tools/ncu_capture_target.py 1 > "$OUT/ncu_full.log" 2>&1; echo "full capture rc=$?"
tail -3 "$OUT/gpu_tests.log"; tail -1 "$OUT/smoke.log"
python - "$OUT" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1] + "/bench.json").read())
print("bench:", d["value"], d["unit"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["achieved"], "/", d["roofline"]["peak"],
      "parity_ok", d.get("parity_ok"), "clocks", d["clocks"])
PY
