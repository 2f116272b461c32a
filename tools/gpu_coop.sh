#!/bin/bash
# cooperative-kernel session: its GPU parity tests, then the per-thread vs cooperative timing sweep
export DART_ENV_NO_REFERENCE=1 DARTB_NO_REBUILD=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_coop.py -m gpu -q --timeout 300 -x > gpurun_out/pytest_coop.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_coop.log
timeout 900 python tools/gpu_sweep.py coop > gpurun_out/sweep_coop.log 2>&1
tail -40 gpurun_out/pytest_coop.log; cat gpurun_out/sweep_coop.log
