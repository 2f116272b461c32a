#!/bin/bash
# round-2 session A: GPU parity suite, large-batch sweep over the developer variants, full bench line
export DARTB_NO_REBUILD=1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
for sfx in "" "$@"; do
  DARTB_SO_SUFFIX=$sfx timeout 400 python tools/gpu_sweep.py r2big > gpurun_out/r2a_sweep$sfx.log 2>&1
done
timeout 900 python bench.py --steps 300 --warmup 20 > gpurun_out/r2a_bench.log 2>gpurun_out/r2a_bench.err
tail -5 gpurun_out/r2a_pytest.log; for f in gpurun_out/r2a_sweep*.log; do echo $f; cat $f; done; tail -c 3000 gpurun_out/r2a_bench.log; tail -5 gpurun_out/r2a_bench.err
