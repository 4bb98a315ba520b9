#!/usr/bin/env python
"""Golden vectors for mid-stream reconfiguration, from the REFERENCE's own code (oracle/_ref: Downsampler.cpp +
Decimators.cpp, Upsampler.cpp + Interpolators.cpp compiled where they lie, EO1 and DB builds).

Downsampler::configure / Upsampler::configure change only decim / fcpos / interp; the six stage objects persist
(include/Decimators.h:57-62, include/Interpolators.h:52-58), so the first outputs after a change depend on what each
stage saw under earlier configurations.  Run in the build container:

    python tests/golden/make_golden_reconfigure.py        ->  tests/golden/reconfigure_ref.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import bindings as ob  # noqa: E402
import cases  # noqa: E402  (the plans are shared with the parity cases)


def main():
    ob.build(ref=True)
    assert ob.ref_available(0) and ob.ref_available(1), "reference build missing"
    rng = np.random.default_rng(20261018)
    out = {}
    dplans = cases.dec_reconfigure_plans()
    for variant in (0, 1):
        for pi in ((0, 1, 6) if variant == 0 else (3, 6)):  # a subset keeps the file small
            plan = dplans[pi]
            n = sum(k for _, _, k in plan)
            x = rng.integers(-32768, 32768, size=(n, 2), dtype=np.int16)
            r = ob.RefDownsampler(plan[0][0], plan[0][1], variant)
            ys, pos = [], 0
            for M, fc, k in plan:
                r.configure(M, fc)
                ys.append(r.process(x[pos:pos + k])[0])
                pos += k
            out[f"dec_v{variant}_p{pi}_in"] = x
            out[f"dec_v{variant}_p{pi}_plan"] = np.array(plan, dtype=np.int64)
            out[f"dec_v{variant}_p{pi}_out"] = np.concatenate(ys)
    for pi, plan in enumerate(cases.int_reconfigure_plans()):
        n = sum(k for _, k in plan)
        x = rng.integers(-32768, 32768, size=(n, 2), dtype=np.int16)
        r = ob.RefUpsampler(plan[0][0], pi & 1)
        ys, pos = [], 0
        for M, k in plan:
            r.configure(M)
            ys.append(r.process(x[pos:pos + k]))
            pos += k
        out[f"int_p{pi}_in"] = x
        out[f"int_p{pi}_plan"] = np.array(plan, dtype=np.int64)
        out[f"int_p{pi}_out"] = np.concatenate(ys)
    np.savez_compressed(os.path.join(HERE, "reconfigure_ref.npz"), **out)
    print("wrote reconfigure_ref.npz:", len(out) // 3, "sequences,", os.path.getsize(os.path.join(HERE, "reconfigure_ref.npz")), "bytes")


if __name__ == "__main__":
    main()
