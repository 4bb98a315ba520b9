#!/usr/bin/env python
"""Golden vectors for the Tx interpolation cascade, from the REFERENCE's own code.

Run in the build container (needs /root/reference; `make -C oracle ref` compiles the reference's
Upsampler.cpp + Interpolators.cpp where they lie, EO1 and DB builds).  Output: interpolator_ref.npz next
to this script; tests compare the oracle (CPU) and the CUDA library (GPU) with it.

    python tests/golden/make_golden_interp.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import bindings as ob  # noqa: E402


def main():
    ob.build(ref=True)
    assert ob.ref_available(0) and ob.ref_available(1), "reference build missing"
    rng = np.random.default_rng(20261018)
    n = 1200
    t = np.arange(n)
    inputs = {
        "random": rng.integers(-32768, 32768, size=(n, 2), dtype=np.int16),
        "tone": np.stack([np.round(16000 * np.cos(2 * np.pi * t / 37.0)), np.round(16000 * np.sin(2 * np.pi * t / 37.0))],
                         axis=1).astype(np.int16),
        "all_min": np.full((n, 2), -32768, np.int16),
    }
    imp = np.zeros((n, 2), np.int16)
    imp[100] = (32767, -32768)
    inputs["impulse"] = imp
    g = {}
    for name, x in inputs.items():
        g[f"in_{name}"] = x
        for M in range(0, 7):
            m = n if M <= 4 else 400
            outs = []
            for variant in (0, 1):
                u = ob.RefUpsampler(M, variant)
                # two calls: state carried across a ragged boundary
                outs.append(np.concatenate([u.process(x[:333]), u.process(x[333:m])]))
            assert np.array_equal(outs[0], outs[1]), "EO1 and DB builds of the reference disagree"
            g[f"out_{name}_M{M}"] = outs[0]
    np.savez_compressed(os.path.join(HERE, "interpolator_ref.npz"), **g)
    print("interpolator_ref.npz:", os.path.getsize(os.path.join(HERE, "interpolator_ref.npz")), "bytes")


if __name__ == "__main__":
    main()
