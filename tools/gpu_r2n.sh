#!/bin/bash
export DARTB_NO_REBUILD=1
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2n_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log
timeout 400 python tools/gpu_sweep.py r2pgs > gpurun_out/r2n_sweep_pgs.log 2>&1
timeout 900 python tools/gpu_soak.py 1500 > gpurun_out/r2n_soak.log 2>&1
tail -4 gpurun_out/r2n_pytest.log; cat gpurun_out/r2n_sweep_pgs.log gpurun_out/r2n_soak.log
