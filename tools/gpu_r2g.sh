#!/bin/bash
# round-2 session G: the quad form — parity suite, sweep against the other forms, register-cap builds
export DARTB_NO_REBUILD=1
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2g_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
timeout 900 python tools/gpu_sweep.py r2quad > gpurun_out/r2g_sweep.log 2>&1
for sfx in _q2 _q3; do
  DARTB_SO_SUFFIX=$sfx timeout 300 python tools/gpu_sweep.py r2quadonly > gpurun_out/r2g_sweep$sfx.log 2>&1
done
grep -E "passed|failed|^FAILED" gpurun_out/r2g_pytest.log | head -20; cat gpurun_out/r2g_sweep.log; for sfx in _q2 _q3; do echo $sfx; cat gpurun_out/r2g_sweep$sfx.log; done
