#!/bin/bash
# queued small-block path, long runs (20000 blocks after 2000 warm-up blocks: the device's clocks have settled)
# columns: decim fecblk min_chain_blocks staging_threads
cd "$(dirname "$0")/../.." || exit 1
(cd tests && python -c "import host_cases; print(host_cases.build('gpu'))") >/dev/null
for cfg in "4 16 0 0" "4 16 16 0" "4 16 16 1" "4 16 16 2" "4 16 16 3" "4 16 0 2" "4 16 8 2" "6 32 16 2" "4 16 16 2"; do
  set -- $cfg; ./tests/host/host_pipeline_gpu blocksq 20000 $1 $2 65536 $3 2000 $4
done
