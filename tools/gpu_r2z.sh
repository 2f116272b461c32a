#!/bin/bash
# session Z: GPU suite with the in-kernel dynamics randomisation
export DARTB_NO_REBUILD=1
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2z_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2z_pytest.log
tail -40 gpurun_out/r2z_pytest.log
