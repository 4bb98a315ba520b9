#!/bin/bash
# kernel durations of the queued small-block path (launch list, serialised: durations only)
cd "$(dirname "$0")/../.." || exit 1
for cfg in "4 16" "6 32"; do
  set -- $cfg
  ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/smallblock_ncu_M$1.csv ./tests/host/host_pipeline_gpu blocksq 200 $1 $2 65536 16 32 > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/smallblock_ncu_M$1.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[:12]: print(r[4][:60], r[-1], r[-2])
PY
done
