#!/bin/bash
# round-2 session F (N GPUs): the driver's launch line for the scaling bench, both arms
export DARTB_NO_REBUILD=1
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 20 > gpurun_out/r2f_bench_${N}gpu.log 2>gpurun_out/r2f_bench_${N}gpu.err
echo "rc=$?"; tail -c 600 gpurun_out/r2f_bench_${N}gpu.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2f_bench_${N}gpu.log') if l.startswith('{')][-1])
print('N', d['n_gpus'], 'value', d['value'], 'us', d['ms_per_step']*1e3, 'e2e', d['e2e']['value'], d['config']['parallelism'])
for cid,c in d['configs'].items():
    print(cid, c['env'], [(r['lcp'], round(r['ms_per_step']*1e3,1), '%.3e'%r['value'], r['parallelism']) for r in c['runs']])
print(d['configs']['5'].get('pgs_sweep'))
PY
