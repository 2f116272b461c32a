#!/bin/bash
# round-2 session H: group forms with 2 / 4 / 8 lanes per world
export DARTB_NO_REBUILD=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_quad.py tests/test_contact_free_envs.py -m gpu -q --timeout 600 > gpurun_out/r2h_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
timeout 900 python tools/gpu_sweep.py r2group > gpurun_out/r2h_sweep.log 2>&1
grep -E "passed|failed|^FAILED" gpurun_out/r2h_pytest.log | head -20; cat gpurun_out/r2h_sweep.log
