#!/bin/bash
export DARTB_NO_REBUILD=1
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/fused_gather_check.py DartHalfCheetah-v1 16384 > gpurun_out/r2m_fused_${N}gpu.log 2>&1
echo "rc=$?"; tail -4 gpurun_out/r2m_fused_${N}gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 tools/fused_gather_check.py DartHopper-v1 4096 >> gpurun_out/r2m_fused_${N}gpu.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/r2m_fused_${N}gpu.log
bash tools/gpu_r2f.sh $N
