#!/bin/bash
# session R: branch-free sincos — GPU suite + the automatic choice at every config size
export DARTB_NO_REBUILD=1
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2r_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2r_pytest.log
timeout 900 python tools/gpu_sweep.py r2final > gpurun_out/r2r_sweep.log 2>&1
tail -12 gpurun_out/r2r_pytest.log; cat gpurun_out/r2r_sweep.log
