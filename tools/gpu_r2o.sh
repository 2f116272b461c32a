#!/bin/bash
export DARTB_NO_REBUILD=1
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2o_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log
timeout 900 python tools/gpu_sweep.py r2final > gpurun_out/r2o_sweep.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -k regex:k_env_step -c 1"
timeout 600 $NCU --launch-skip 60 -o gpurun_out/r2_cheetah16k_v2 -f python bench.py --config 4 --steps 40 --warmup 10 --no-extras > gpurun_out/r2o_ncu_cheetah.log 2>&1
tail -4 gpurun_out/r2o_pytest.log; cat gpurun_out/r2o_sweep.log
