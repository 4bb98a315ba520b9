#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
(cd tests && python -c "import host_cases; print(host_cases.build('gpu'))") >/dev/null
for cfg in "4 16" "6 32" "5 32" "6 0" "4 0" "6 1" "4 1"; do set -- $cfg; ./tests/host/host_pipeline_gpu blocksq 20000 $1 $2 65536 16 2000; done
