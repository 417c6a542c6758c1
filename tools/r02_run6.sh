set -u
OUT=gpurun_out/r02k
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python tools/size_sweep.py > $OUT/size_sweep.jsonl 2> $OUT/size_sweep.err; echo "size sweep rc=$?"; cat $OUT/size_sweep.jsonl
for tool in memcheck synccheck racecheck initcheck; do
  timeout 280 compute-sanitizer --tool $tool python tools/tiny_probe.py > $OUT/sanitizer_$tool.log 2>&1; echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|tiny probe ok" $OUT/sanitizer_$tool.log | tail -3
done
