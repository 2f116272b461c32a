#!/bin/bash
export DARTB_NO_REBUILD=1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/symm_probe.py > gpurun_out/r2l_symm.log 2>&1
tail -30 gpurun_out/r2l_symm.log
timeout 600 python -m pytest tests/test_gpu_quad.py tests/test_gpu_parity.py -m gpu -q --timeout 600 > gpurun_out/r2l_pytest.log 2>&1; tail -3 gpurun_out/r2l_pytest.log
timeout 300 python tools/gpu_sweep.py r2q4mid > gpurun_out/r2l_sweep.log 2>&1; cat gpurun_out/r2l_sweep.log
