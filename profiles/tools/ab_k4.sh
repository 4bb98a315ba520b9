#!/bin/bash
# usage: ab_k4.sh name [lib] [log2_interp]  -> one config-6 bench line (K4 ms, value) for a variant build of the library
name=$1; lib=$2; L=${3:-4}
if [ -n "$lib" ]; then export SDRD_B200_LIB=$lib; fi
python bench.py --config 6 --log2-decim $L --steps 300 --no-cpu --no-e2e > gpurun_out/abk4_$name.json 2> gpurun_out/abk4_$name.err
python - <<PY
import json
d=json.loads(open("gpurun_out/abk4_$name.json").read().strip().splitlines()[-1])
print("$name", "x", 1 << $L, "value", d["value"], "ms_per_step", d["ms_per_step"], "frac", d["roofline"]["frac"], d["config"].get("parity"))
PY
