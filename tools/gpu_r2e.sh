#!/bin/bash
# round-2 session E: rolled-collision A/B on the large-batch kernels, PGS after the chain restructuring, e2e probe
export DARTB_NO_REBUILD=1
mkdir -p gpurun_out
for sfx in "" _rc; do
  for rep in 1 2; do
    DARTB_SO_SUFFIX=$sfx timeout 400 python tools/gpu_sweep.py r2rc >> gpurun_out/r2e_sweep$sfx.log 2>&1
  done
done
timeout 300 python tools/e2e_probe.py > gpurun_out/r2e_e2e_probe.log 2>&1
DARTB_SO_SUFFIX=_rc timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q --timeout 600 > gpurun_out/r2e_pytest_rc.log 2>&1
for f in gpurun_out/r2e_sweep.log gpurun_out/r2e_sweep_rc.log; do echo $f; cat $f; done; cat gpurun_out/r2e_e2e_probe.log; tail -3 gpurun_out/r2e_pytest_rc.log
