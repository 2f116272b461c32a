#!/bin/bash
# session P: GPU suite with the per-world dynamics parameters, crossover sweep after the branch-free solvers,
# steady-state ncu captures (launch inside the device-timed region) of the final build
export DARTB_NO_REBUILD=1
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2p_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log
timeout 1200 python tools/gpu_sweep.py r2p > gpurun_out/r2p_sweep.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -k regex:k_env_step -c 1 --launch-skip 25"
timeout 400 $NCU -o gpurun_out/r2_hopper_quad_v2 -f python bench.py --steps 40 --warmup 10 --no-extras > gpurun_out/r2p_ncu_hopper.log 2>&1
timeout 400 $NCU -o gpurun_out/r2_walker16k_pgs_v2 -f python bench.py --config 3 --steps 40 --warmup 10 --no-extras > gpurun_out/r2p_ncu_walker.log 2>&1
tail -4 gpurun_out/r2p_pytest.log; cat gpurun_out/r2p_sweep.log
