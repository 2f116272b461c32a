#!/bin/bash
# One gpurun session: GPU tests, smoke, bench (both arms), optional ncu launch list + ONE full capture.
export DART_ENV_NO_REFERENCE=1 DARTB_NO_REBUILD=1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --steps 500 --warmup 50 > gpurun_out/bench.log 2>&1
timeout 600 python bench.py --impl reference --steps 200 --warmup 10 > gpurun_out/bench_reference.log 2>&1
if [ "$1" == "ncu" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 60 --warmup 10 > gpurun_out/bench_under_ncu.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_env_step -s 40 -c 1 -o gpurun_out/prof_hopper -f python bench.py --steps 60 --warmup 10 > gpurun_out/ncu_full.log 2>&1
fi
tail -3 gpurun_out/pytest_gpu.log; tail -4 gpurun_out/smoke.log; tail -1 gpurun_out/bench.log; tail -1 gpurun_out/bench_reference.log
