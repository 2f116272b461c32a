#!/bin/bash
# session W: does the nvidia-smi poll explain the run-to-run spread of the end-to-end figure?
export DARTB_NO_REBUILD=1
mkdir -p gpurun_out
{
for a in 0 0 0 1 1; do
  echo "--- BENCH_SAMPLE_DURING_E2E=$a"
  BENCH_SAMPLE_DURING_E2E=$a timeout 300 python bench.py --steps 1000 --warmup 50 --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('flushed us %.1f warm %.1f e2e us %.1f' % (d['ms_per_step']*1e3, d['ms_per_step_l2_warm']*1e3, d['e2e']['ms_per_step']*1e3), d['e2e']['us_per_call_rank0'], d['clocks'])"
done
} > gpurun_out/r2w_probe.log 2>&1
cat gpurun_out/r2w_probe.log
