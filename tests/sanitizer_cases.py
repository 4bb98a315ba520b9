"""The cases compute-sanitizer runs over (profiles/tools/sanitize.sh, sanitize_extra.sh): smoke() plus the round-2
kernels -- stage-state kernels after a reconfiguration, streaming decode with every branch, the queued Rx path, the
cm256 descriptor entry points -- each checked against the oracle as in the tests.  Test infrastructure: lives here
because it uses oracle/."""
import sys
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
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
