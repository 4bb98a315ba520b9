#!/bin/bash
# kernel durations inside one queued chain of 1 .. 32 blocks (launch list: durations only)
cd "$(dirname "$0")/../.." || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/chain_ncu.csv ./tests/host/host_pipeline_gpu chain 4 16 65536 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/chain_ncu.csv')) if len(r)>10 and r[0].isdigit()]
print(len(rows))
# last repetitions of each size: print unique (kernel, grid, duration) tail
seen={}
for r in rows:
    key=(r[4][:50], r[7] if len(r)>7 else '')
    seen.setdefault(key, []).append(float(r[-1]))
for k,v in seen.items():
    v=sorted(v); print(k, len(v), 'median', v[len(v)//2], 'min', v[0])
PY
