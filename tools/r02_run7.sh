set -u
OUT=gpurun_out/r02l
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 120 python tools/tiny_fullbox.py 2>&1 | tail -2
for tool in memcheck racecheck; do
  TINY_SKIP_NCCL=1 timeout 400 compute-sanitizer --tool $tool python tools/tiny_fullbox.py > $OUT/fullbox_$tool.log 2>&1; echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|tiny full-box probe ok" $OUT/fullbox_$tool.log | tail -3
done
timeout 300 python -m pytest tests -m gpu -x -q -k "multi_device or peer_push or probe_by_uuid or storm" 2>&1 | tail -3
