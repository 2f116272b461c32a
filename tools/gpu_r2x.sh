#!/bin/bash
# session X: prologue load order — GPU suite, bench (flushed / warm / end to end), twice
export DARTB_NO_REBUILD=1
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.log
timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/r2x_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2x_pytest.log
{
for a in 1 2 3; do
  timeout 300 python bench.py --steps 1000 --warmup 50 --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('flushed us %.1f warm %.1f e2e us %.1f' % (d['ms_per_step']*1e3, d['ms_per_step_l2_warm']*1e3, d['e2e']['ms_per_step']*1e3), d['e2e']['us_per_call_rank0'], d['clocks'])"
done
} > gpurun_out/r2x_bench.log 2>&1
tail -4 gpurun_out/r2x_pytest.log; cat gpurun_out/r2x_bench.log
