#!/bin/bash
# A/B session: GPU parity tests on the product library, then the main timing sweep on the product
# library and on each developer variant given as argument (DARTB_SO_SUFFIX values, e.g. _chol).
export DART_ENV_NO_REFERENCE=1 DARTB_NO_REBUILD=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python tools/gpu_sweep.py main > gpurun_out/sweep_main.log 2>&1
for sfx in "$@"; do
  DARTB_SO_SUFFIX=$sfx timeout 600 python tools/gpu_sweep.py main > gpurun_out/sweep_main$sfx.log 2>&1
done
timeout 600 python bench.py --steps 500 --warmup 50 > gpurun_out/bench.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; for f in gpurun_out/sweep_main*.log; do echo $f; cat $f; done; tail -1 gpurun_out/bench.log
