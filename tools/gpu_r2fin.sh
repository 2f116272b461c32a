#!/bin/bash
# last build of round 2: smoke(), both bench arms, the launch list (the GPU suite runs in gpu_r2z.sh)
export DARTB_NO_REBUILD=1
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2fin_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2fin_smoke.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > gpurun_out/r2fin_bench_ref.log 2>gpurun_out/r2fin_bench_ref.err
timeout 900 python bench.py --gpus 1 --steps 1000 --warmup 50 > gpurun_out/r2fin_bench.log 2>gpurun_out/r2fin_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_hopper_launches.csv python bench.py --steps 40 --warmup 10 --no-extras > gpurun_out/r2fin_launches.log 2>&1
tail -8 gpurun_out/r2fin_smoke.log; python - <<PY
import json
r=json.loads([l for l in open('gpurun_out/r2fin_bench_ref.log') if l.startswith('{')][-1])
d=json.loads([l for l in open('gpurun_out/r2fin_bench.log') if l.startswith('{')][-1])
print('reference', r['value'], r['config']['workload'])
print('ours     ', d['value'], d['config']['workload'], d['kernel'])
print('us flushed', d['ms_per_step']*1e3, 'warm', d['ms_per_step_l2_warm']*1e3, 'e2e us', d['e2e']['ms_per_step']*1e3, 'e2e value', d['e2e']['value'], 'cpu_baseline', d['cpu_baseline']['value'])
print('ratio device', d['value']/r['value'], 'e2e', d['e2e']['value']/r['value'], 'same_config', r['config']['workload']==d['config']['workload'])
print('roofline', {k:d['roofline'][k] for k in ('achieved','peak','frac','traffic','kernel','issue_active_pct','avg_active_lanes') if k in d['roofline']}, d['clocks'], d['gpu_launches'])
for cid,c in d['configs'].items():
    print(cid, c['env'], [(x['lcp'], round(x['ms_per_step']*1e3,1), round(x['ms_per_step_l2_warm']*1e3,1), '%.3e'%x['value'], x['roofline']['kernel'], x['roofline'].get('issue_active_pct')) for x in c['runs']], '%.3e'%c['cpu_baseline']['value'])
print([ (p['pgs_iters'], round(p['us_per_step_l2_warm'],1)) for p in d['configs']['5']['pgs_sweep']])
PY
