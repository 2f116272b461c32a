#!/bin/bash
# round-2 session I: full suite on the three-form build, crossover sweep, bench line, ncu captures
export DARTB_NO_REBUILD=1
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2i_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
timeout 600 python tools/gpu_sweep.py r2xover > gpurun_out/r2i_xover.log 2>&1
timeout 900 python bench.py --steps 300 --warmup 20 > gpurun_out/r2i_bench.log 2>gpurun_out/r2i_bench.err
NCU="ncu --set full --clock-control none --import-source on -k regex:k_env_step -c 1"
timeout 600 $NCU --launch-skip 30 -o gpurun_out/r2_hopper_quad -f python bench.py --steps 40 --warmup 10 --no-extras > gpurun_out/r2i_ncu_hopper.log 2>&1
timeout 600 $NCU --launch-skip 30 -o gpurun_out/r2_snake_quad -f python bench.py --config 5 --steps 40 --warmup 10 --no-extras > gpurun_out/r2i_ncu_snake.log 2>&1
grep -E "passed|failed|^FAILED" gpurun_out/r2i_pytest.log | head -20; cat gpurun_out/r2i_xover.log; python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2i_bench.log') if l.startswith('{')][-1])
print('headline', d['kernel'], 'us', d['ms_per_step']*1e3, 'warm', d['ms_per_step_l2_warm']*1e3, 'e2e us', d['e2e']['ms_per_step']*1e3, 'cpu', d['cpu_baseline']['value'])
for cid,c in d['configs'].items():
    print(cid, c['env'], [(r['lcp'], round(r['ms_per_step']*1e3,1), round(r['ms_per_step_l2_warm']*1e3,1), r['roofline']['kernel']) for r in c['runs']])
print(d['configs']['5'].get('pgs_sweep'))
PY
