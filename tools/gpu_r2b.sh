#!/bin/bash
# round-2 session B: GPU suite after the fixes, e2e breakdown, bench line
export DARTB_NO_REBUILD=1
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
timeout 300 python tools/e2e_probe.py > gpurun_out/r2b_e2e_probe.log 2>&1
timeout 900 python bench.py --steps 300 --warmup 20 --no-extras > gpurun_out/r2b_bench.log 2>gpurun_out/r2b_bench.err
grep -E "passed|failed|^FAILED|^E  " gpurun_out/r2b_pytest.log | head -40; cat gpurun_out/r2b_e2e_probe.log; tail -c 1500 gpurun_out/r2b_bench.log; tail -3 gpurun_out/r2b_bench.err
