#!/bin/bash
# compute-sanitizer over smoke() plus the round-2 kernels (stage-state kernels after a reconfiguration, streaming
# decode with every branch, queued Rx path); logs under gpurun_out/
cd "$(dirname "$0")/../.." || exit 1
for tool in memcheck racecheck; do
  compute-sanitizer --tool $tool --error-exitcode 9 python tests/sanitizer_cases.py > gpurun_out/sanitizer_${tool}_r2.log 2>&1
  echo "$tool rc=$?"; tail -3 gpurun_out/sanitizer_${tool}_r2.log
done
