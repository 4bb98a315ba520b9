#!/bin/bash
# compute-sanitizer over smoke() plus the round-2 kernels (stage-state kernels after a reconfiguration, streaming
# decode with every branch, queued Rx path); logs under gpurun_out/
cd "$(dirname "$0")/../.." || exit 1
cat > /tmp/san_cases.py <<'PY'
import sys
import os
sys.path.insert(0, os.getcwd())
sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import __graft_entry__ as g
import cases
from oracle import bindings as ob
from sdrdaemon_b200 import capi
g.smoke()
lib = capi.load()
rng = np.random.default_rng(5)
plan = [(4, 2, 5000), (2, 2, 9000), (6, 0, 70000), (3, 1, 4096 + 128)]
cases.check_decimator_reconfigure(lib, ob, 1, cases.rand_iq(rng, (2, sum(k for _, _, k in plan))), plan)
cases.check_interpolator_reconfigure(lib, ob, cases.rand_iq(rng, (2, 900)), [(4, 300), (2, 100), (5, 400), (6, 100)])
x, frames = cases.make_frames(ob, rng, 13, 40)
sb, nb = cases.pack_received(frames, cases.erasure_cases(rng, frames, 40))
cases.check_decode(lib, ob, sb, nb)
cases.check_rx_queued(lib, ob, M=4, F=16, S=2, blk=65536, n_blk=12, max_blocks=4)
cases.check_cm256_blocks(lib, ob)
print("sanitizer cases ok")
PY
for tool in memcheck racecheck; do
  compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san_cases.py > gpurun_out/sanitizer_${tool}_r2.log 2>&1
  echo "$tool rc=$?"; tail -3 gpurun_out/sanitizer_${tool}_r2.log
done
