#!/bin/bash
# round-2 session D: suite (incl. TMA digest test), TMA A/B on the headline kernel, PGS sweep, bench line with all configs
export DARTB_NO_REBUILD=1
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
for mode in 0 1 2; do
  for rep in 1 2; do
    DARTB_COOP_TMA=$mode timeout 300 python tools/gpu_sweep.py coopq >> gpurun_out/r2d_tma_$mode.log 2>&1
  done
  DARTB_COOP_TMA=$mode timeout 300 python bench.py --steps 500 --warmup 30 --no-extras > gpurun_out/r2d_bench_tma$mode.log 2>&1
done
timeout 400 python tools/gpu_sweep.py r2pgs > gpurun_out/r2d_sweep_pgs.log 2>&1
timeout 900 python bench.py --steps 300 --warmup 20 > gpurun_out/r2d_bench.log 2>gpurun_out/r2d_bench.err
grep -E "passed|failed|^FAILED" gpurun_out/r2d_pytest.log | head; for m in 0 1 2; do echo "TMA mode $m"; cat gpurun_out/r2d_tma_$m.log; python -c "
import json,sys
d=json.loads(open('gpurun_out/r2d_bench_tma$m.log').read().strip().splitlines()[-1]); print('bench flushed us', d['ms_per_step']*1e3, 'warm', d['ms_per_step_l2_warm']*1e3, 'e2e us', d['e2e']['ms_per_step']*1e3)"; done; cat gpurun_out/r2d_sweep_pgs.log | head -12
