#!/bin/bash
# session Y: A/B of the kernel-parameter prefetch at entry (libdartb_pf.so: hopper_f, walker_f)
export DARTB_NO_REBUILD=1
mkdir -p gpurun_out
{
for sfx in "" _pf "" _pf; do
  echo "--- suffix '$sfx'"
  DARTB_SO_SUFFIX=$sfx timeout 300 python bench.py --steps 1000 --warmup 50 --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('flushed us %.2f warm %.2f e2e us %.1f' % (d['ms_per_step']*1e3, d['ms_per_step_l2_warm']*1e3, d['e2e']['ms_per_step']*1e3), d['e2e']['us_per_call_rank0'])"
done
for sfx in "" _pf; do
  echo "--- sweep suffix '$sfx'"
  DARTB_SO_SUFFIX=$sfx timeout 300 python tools/gpu_sweep.py r2y
done
} > gpurun_out/r2y_ab.log 2>&1
cat gpurun_out/r2y_ab.log
