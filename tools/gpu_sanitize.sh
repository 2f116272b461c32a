#!/bin/bash
# compute-sanitizer memcheck over the per-world / randomisation tests and smoke() (small batches; the tool is 10-50x slow)
export DARTB_NO_REBUILD=1
mkdir -p gpurun_out
timeout 280 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/sanitize_tests.txt \
  python -m pytest tests/test_gpu_round2.py -q -x -m gpu -k "randomize or bodynode or per_world" > gpurun_out/sanitize_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/sanitize_tests.log
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/sanitize_smoke.txt \
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/sanitize_smoke.log
tail -5 gpurun_out/sanitize_tests.log; tail -5 gpurun_out/sanitize_tests.txt; tail -3 gpurun_out/sanitize_smoke.log; tail -5 gpurun_out/sanitize_smoke.txt
