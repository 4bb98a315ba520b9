#!/usr/bin/env python
"""Generate the golden vectors in tests/golden/ from the REFERENCE's own code.

Run in the build container (needs /root/reference; `make -C oracle ref` compiles the reference's
Downsampler/Decimators (EO1 and DB builds), UDPSinkFEC and SDRdaemonFECBuffer from the sources where
they lie -- cm256cc, absent from the reference tree, is replaced by the restated CM256).  The outputs
are small .npz files committed next to this script; tests compare the oracle (CPU) and the CUDA
library (GPU) with them, so that parity does not depend on /root/reference at test time.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import bindings as ob  # noqa: E402

FRAME = 127 * 127


def tone(n, srate, dfp, amp):
    """What TestSource::read_samples produces (TestSource.cpp:395-416), taken from the reference itself."""
    buf, _ = ob.ref_testsource(n, srate, 2.0 * np.pi * dfp / srate, amp)
    return buf


def main():
    ob.build(ref=True)
    assert ob.ref_available(0) and ob.ref_available(1), "reference build missing"
    rng = np.random.default_rng(20261017)

    # ---- decimator: reference Downsampler::process, EO1 (x86/SSE4.1) and DB builds -----------------
    n = 8192
    inputs = {
        "random": rng.integers(-32768, 32768, size=(n, 2), dtype=np.int16),
        "tone": tone(n, 8000, 333.0, 0.5),      # 8 kS/s keeps TestSource's real-time usleep short
        "all_min": np.full((n, 2), -32768, np.int16),
    }
    imp = np.zeros((n, 2), np.int16)
    imp[1000] = (32767, -32768)
    inputs["impulse"] = imp
    dec = {}
    for name, x in inputs.items():
        dec[f"in_{name}"] = x
        for variant in (0, 1):
            for fcpos in (0, 1, 2):
                for M in range(0, 7):
                    if name != "random" and fcpos != 2:
                        continue
                    d = ob.RefDownsampler(M, fcpos, variant)
                    # two calls: state carried across a non-power-of-two boundary
                    y = np.concatenate([d.process(x[:3000])[0], d.process(x[3000:])[0]])
                    dec[f"out_{name}_v{variant}_fc{fcpos}_M{M}"] = y
    # 8- and 12-bit sources (shift rule of Decimators.cpp:408-409)
    for bits in (8, 12):
        x = (inputs["random"] >> (16 - bits)).astype(np.int16)
        dec[f"in_random_b{bits}"] = x
        for M in range(0, 7):
            d = ob.RefDownsampler(M, 2, 0)
            y, ss = d.process(x, bits)
            dec[f"out_random_b{bits}_M{M}"] = y
            dec[f"ss_random_b{bits}_M{M}"] = np.array([ss])
    np.savez_compressed(os.path.join(HERE, "decimator_ref.npz"), **dec)

    # ---- sender: reference UDPSinkFEC over loop-back UDP (restated CM256 inside) -------------------
    # block 0 carries gettimeofday(): bytes 16..27 of block 0 (tv_sec, tv_usec, crc32) are wall-clock
    # dependent and so are the recovery blocks' first 28 payload bytes; tests mask/recompute them.
    snd = {}
    for F in (4, 16):
        x = rng.integers(-32768, 32768, size=(FRAME * 3, 2), dtype=np.int16)
        dg = ob.ref_sink_run(x, F, 2, port=19100 + F, chunk=4096)
        assert dg.shape == (2 * (128 + F), 512), dg.shape
        snd[f"in_F{F}"] = x
        snd[f"dgrams_F{F}"] = dg
    np.savez_compressed(os.path.join(HERE, "sink_ref.npz"), **snd)

    # ---- receiver: reference SDRdaemonFECBuffer::writeAndRead ---------------------------------------
    F = 16
    sk = ob.Sink(n_fec=F)
    x = rng.integers(-32768, 32768, size=(FRAME * 7, 2), dtype=np.int16)
    sk.write(x)
    frames = np.stack(sk.frames)
    patterns = []
    for f in range(7):
        if f == 0:
            sel = list(range(128))
        elif f == 1:
            sel = [i for i in range(128) if i != 101] + [128]          # SDRDAEMON_PUNCTURE 101
        elif f == 2:
            er = set(rng.choice(128, 12, replace=False).tolist())
            sel = [i for i in range(128) if i not in er] + list(range(128, 140))
        elif f == 3:
            er = set(rng.choice(128, 16, replace=False).tolist())
            sel = [i for i in range(128) if i not in er] + list(range(128, 144))
        elif f == 4:
            sel = list(range(90))                                      # incomplete: zeros for the rest
        elif f == 5:
            er = set(rng.choice(128, 10, replace=False).tolist())
            sel = [i for i in range(128) if i not in er] + rng.choice(np.arange(128, 144), 10, replace=False).tolist()
            rng.shuffle(sel)                                           # out of order arrival (copy-back quirk)
            sel = [int(v) for v in sel]
        else:
            sel = list(range(128))
        patterns.append(sel)
    buf = ob.RefFecBuffer()
    outs = []
    stream = []
    for f, sel in enumerate(patterns):
        for i in sel:
            stream.append(frames[f][i])
    stream.append(frames[6][0] ^ np.uint8(0))  # already in; a new frame index flushes the last frame
    last = frames[6][0].copy()
    last[0] = 99  # frame index 99: forces the roll-over that emits frame 6
    stream.append(last)
    for sb in stream:
        r = buf.write_and_read(sb)
        if r is not None:
            outs.append(r.copy())
    # the first emission is the reference's uninitialised slot (SURVEY H4b iii): drop it
    outs = outs[1:]
    assert len(outs) == 7, len(outs)
    rcv = {"frames": frames, "payload": np.stack(outs)}
    for f, sel in enumerate(patterns):
        rcv[f"sel_{f}"] = np.array(sel, np.int32)
    np.savez_compressed(os.path.join(HERE, "fecbuffer_ref.npz"), **rcv)
    for fn in ("decimator_ref.npz", "sink_ref.npz", "fecbuffer_ref.npz"):
        print(fn, os.path.getsize(os.path.join(HERE, fn)), "bytes")


if __name__ == "__main__":
    main()
