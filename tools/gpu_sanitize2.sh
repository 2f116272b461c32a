#!/bin/bash
# compute-sanitizer racecheck (shared-memory hazards) and initcheck (uninitialised global reads) over smoke()
export DARTB_NO_REBUILD=1
mkdir -p gpurun_out
for tool in racecheck initcheck; do
  timeout 150 compute-sanitizer --tool $tool --error-exitcode 9 --log-file gpurun_out/sanitize_${tool}.txt \
    python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_${tool}.log 2>&1
  echo "$tool rc=$?"; tail -2 gpurun_out/sanitize_${tool}.log; tail -6 gpurun_out/sanitize_${tool}.txt
done
