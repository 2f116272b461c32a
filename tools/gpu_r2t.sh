#!/bin/bash
# session T: ftz rsqrt in the product; A/B of -prec-div=false -prec-sqrt=false -ftz=true (libdartb_fm.so)
export DARTB_NO_REBUILD=1
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.log
timeout 600 python tools/gpu_sweep.py r2s > gpurun_out/r2t_sweep.log 2>&1
DARTB_SO_SUFFIX=_fm timeout 600 python tools/gpu_sweep.py r2s > gpurun_out/r2t_sweep_fm.log 2>&1
DARTB_SO_SUFFIX=_fm timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2t_pytest_fm.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2t_pytest_fm.log
cat gpurun_out/r2t_sweep.log; cat gpurun_out/r2t_sweep_fm.log; tail -12 gpurun_out/r2t_pytest_fm.log
