#!/bin/bash
# round-2 session C: suite, PGS-path sweep, e2e breakdown (zero-copy vs copies), ncu captures for profiles/r2_*
export DARTB_NO_REBUILD=1
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
timeout 400 python tools/gpu_sweep.py r2pgs > gpurun_out/r2c_sweep_pgs.log 2>&1
timeout 300 python tools/e2e_probe.py > gpurun_out/r2c_e2e_probe.log 2>&1
DARTB_ZEROCOPY=0 timeout 300 python tools/e2e_probe.py > gpurun_out/r2c_e2e_probe_copies.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -k regex:k_env_step -c 1"
timeout 600 $NCU --launch-skip 30 -o gpurun_out/r2_hopper_coop -f python bench.py --steps 40 --warmup 10 --no-extras > gpurun_out/r2c_ncu_hopper.log 2>&1
timeout 600 $NCU --launch-skip 60 -o gpurun_out/r2_walker16k_pgs -f python bench.py --config 3 --steps 40 --warmup 10 --no-extras > gpurun_out/r2c_ncu_walker.log 2>&1
timeout 600 $NCU --launch-skip 60 -o gpurun_out/r2_cheetah16k -f python bench.py --config 4 --steps 40 --warmup 10 --no-extras > gpurun_out/r2c_ncu_cheetah.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_hopper_launches.csv python bench.py --steps 40 --warmup 10 --no-extras > gpurun_out/r2c_launches.log 2>&1
grep -E "passed|failed|^FAILED" gpurun_out/r2c_pytest.log | head; cat gpurun_out/r2c_sweep_pgs.log gpurun_out/r2c_e2e_probe.log gpurun_out/r2c_e2e_probe_copies.log; ls -la gpurun_out/*.ncu-rep | tail -5
