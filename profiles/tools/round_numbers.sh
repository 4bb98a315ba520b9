#!/bin/bash
# every single-GPU bench line of the round, in one go (run through gpurun): results under gpurun_out/<tag>_*.json
tag=${1:-r2}
cd "$(dirname "$0")/../.." || exit 1
python bench.py 2>gpurun_out/${tag}_bench.err > gpurun_out/${tag}_bench.json
python bench.py --impl reference --steps 5 --warmup 1 2>>gpurun_out/${tag}_bench.err > gpurun_out/${tag}_bench_reference.json
for c in 3 5; do python bench.py --config $c --steps 200 --no-cpu 2>>gpurun_out/${tag}_bench.err > gpurun_out/${tag}_bench_config$c.json; done
python bench.py --config 4 --steps 300 2>>gpurun_out/${tag}_bench.err > gpurun_out/${tag}_bench_config4.json
python bench.py --config 6 --steps 300 2>>gpurun_out/${tag}_bench.err > gpurun_out/${tag}_bench_config6.json
for f in gpurun_out/${tag}_bench*.json; do echo "== $f"; cut -c1-400 $f; done
