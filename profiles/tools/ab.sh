#!/bin/bash
# usage: ab.sh name [lib]  -> one bench line (K1 ms, value)
name=$1; lib=$2
if [ -n "$lib" ]; then export SDRD_B200_LIB=$lib; fi
python bench.py --steps 300 --no-cpu --no-e2e > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
python - <<PY
import json
d=json.load(open("gpurun_out/ab_$name.json"))
print("$name", "value", d["value"], "k1_ms", d["roofline"]["k1_ms_per_launch"], "step_ms", d["ms_per_step"], d["config"]["parity"])
PY
