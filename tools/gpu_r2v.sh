#!/bin/bash
# session V: is the end-to-end step sensitive to where the process runs?  topology, launch floor, e2e with / without NVML affinity
export DARTB_NO_REBUILD=1
mkdir -p gpurun_out
{
nvidia-smi topo -m; lscpu | grep -i "numa\|model name\|socket\|^cpu(s)"; python -c "import os; print('affinity', len(os.sched_getaffinity(0)))"
echo "--- launch probe (unbound)"; tools/launch_probe
for a in 0 1 0 1; do
  echo "--- bench BENCH_CPU_AFFINITY=$a"
  BENCH_CPU_AFFINITY=$a timeout 300 python bench.py --steps 300 --warmup 20 --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('flushed us', d['ms_per_step']*1e3, 'warm', d['ms_per_step_l2_warm']*1e3, 'e2e us', d['e2e']['ms_per_step']*1e3, d['config'].get('cpu_affinity'), 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])"
done
} > gpurun_out/r2v_probe.log 2>&1
cat gpurun_out/r2v_probe.log
