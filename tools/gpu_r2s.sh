#!/bin/bash
# session S: approximate sqrt in the LCP tolerance scales — GPU suite + quick A/B set
export DARTB_NO_REBUILD=1
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2s_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2s_pytest.log
timeout 600 python tools/gpu_sweep.py r2s > gpurun_out/r2s_sweep.log 2>&1
tail -12 gpurun_out/r2s_pytest.log; cat gpurun_out/r2s_sweep.log
