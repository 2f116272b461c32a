#!/bin/bash
export DARTB_NO_REBUILD=1
mkdir -p gpurun_out
for sfx in "" _q4; do
  DARTB_SO_SUFFIX=$sfx timeout 400 python tools/gpu_sweep.py r2quadonly > gpurun_out/r2j_sweep$sfx.log 2>&1
  DARTB_SO_SUFFIX=$sfx timeout 400 python tools/gpu_sweep.py r2q4mid >> gpurun_out/r2j_sweep$sfx.log 2>&1
done
for sfx in "" _q4; do echo "suffix '$sfx'"; cat gpurun_out/r2j_sweep$sfx.log; done
