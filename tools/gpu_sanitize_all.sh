#!/bin/bash
# compute-sanitizer memcheck over the whole GPU suite
export DARTB_NO_REBUILD=1
mkdir -p gpurun_out
timeout 330 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/sanitize_all.txt \
  python -m pytest tests -q -m gpu > gpurun_out/sanitize_all.log 2>&1
echo "tests rc=$?" >> gpurun_out/sanitize_all.log
tail -6 gpurun_out/sanitize_all.log; tail -12 gpurun_out/sanitize_all.txt
