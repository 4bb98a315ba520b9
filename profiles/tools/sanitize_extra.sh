#!/bin/bash
# compute-sanitizer initcheck + synccheck over the same cases as sanitize.sh (tests/sanitizer_cases.py)
cd "$(dirname "$0")/../.." || exit 1
for tool in initcheck synccheck; do
  compute-sanitizer --tool $tool --error-exitcode 9 python tests/sanitizer_cases.py > gpurun_out/sanitizer_${tool}_r2.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|sanitizer cases ok" gpurun_out/sanitizer_${tool}_r2.log | tail -2
done
