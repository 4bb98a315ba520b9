#!/bin/bash
# usage: abm.sh name lib M frames
name=$1; lib=$2; M=$3; fr=$4
if [ -n "$lib" ] && [ "$lib" != "-" ]; then export SDRD_B200_LIB=$lib; fi
python bench.py --steps 200 --no-cpu --no-e2e --log2-decim $M --frames $fr > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
python - <<PY
import json
d=json.load(open("gpurun_out/ab_$name.json"))
print("$name", "value", d["value"], "k1_ms", d["roofline"]["k1_ms_per_launch"], "step_ms", d["ms_per_step"], d["config"]["parity"])
PY
