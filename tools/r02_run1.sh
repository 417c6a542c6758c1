set -u
mkdir -p gpurun_out/r02b
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r02b/gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -25 gpurun_out/r02b/gpu_tests.log
timeout 120 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r02b/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02b/smoke.log
timeout 600 python tools/r02_quick.py > gpurun_out/r02b/quick.jsonl 2> gpurun_out/r02b/quick.err; echo "quick rc=$?"
cat gpurun_out/r02b/quick.jsonl
