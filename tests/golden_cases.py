"""Checks against the committed golden vectors (tests/golden/*.npz, produced by the reference's own
code through tests/golden/make_golden.py).  `dec_factory(M, fcpos, variant)` returns an object with
.process(x, bits) -> (y, ss); used with the oracle (CPU tests) and with the CUDA library (GPU tests)."""
import os
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
FRAME = 127 * 127


def load(name):
    return np.load(os.path.join(HERE, "golden", name))


def check_decimator_golden(dec_factory):
    g = load("decimator_ref.npz")
    n_checked = 0
    for key in g.files:
        if not key.startswith("out_"):
            continue
        parts = key.split("_")
        if parts[-1].startswith("M") and parts[-2].startswith("fc"):
            name = "_".join(parts[1:-3])
            variant, fcpos, M = int(parts[-3][1:]), int(parts[-2][2:]), int(parts[-1][1:])
            x = g[f"in_{name}"]
            d = dec_factory(M, fcpos, variant)
            y = np.concatenate([d.process(x[:3000], 16)[0], d.process(x[3000:], 16)[0]])
            assert y.shape == g[key].shape and np.array_equal(y, g[key]), key
        else:  # out_random_b{bits}_M{M}
            bits, M = int(parts[2][1:]), int(parts[3][1:])
            x = g[f"in_random_b{bits}"]
            y, ss = dec_factory(M, 2, 0).process(x, bits)
            assert np.array_equal(y, g[key]), key
            assert ss == int(g[f"ss_random_b{bits}_M{M}"][0]), key
        n_checked += 1
    assert n_checked >= 90


def check_interpolator_golden(int_factory):
    """int_factory(M) -> object with .process(x) -> y (state carried across calls)."""
    g = load("interpolator_ref.npz")
    n_checked = 0
    for key in g.files:
        if not key.startswith("out_"):
            continue
        name, M = key[4:key.rindex("_M")], int(key[key.rindex("_M") + 2:])
        x = g[f"in_{name}"]
        m = len(x) if M <= 4 else 400
        u = int_factory(M)
        y = np.concatenate([u.process(x[:333]), u.process(x[333:m])])
        assert y.shape == g[key].shape and np.array_equal(y, g[key]), key
        n_checked += 1
    assert n_checked == 28


def check_reconfigure_golden(dec_factory, int_factory):
    """Mid-stream Downsampler::configure / Upsampler::configure sequences recorded from the reference build
    (make_golden_reconfigure.py).  dec_factory(M, fcpos, variant) / int_factory(M) return objects with
    .configure(...) and .process(x)."""
    g = load("reconfigure_ref.npz")
    n_checked = 0
    for key in g.files:
        if not key.endswith("_plan"):
            continue
        base = key[:-5]
        x, plan, want = g[base + "_in"], g[key], g[base + "_out"]
        ys, pos = [], 0
        if base.startswith("dec_"):
            variant = int(base.split("_")[1][1:])
            d = dec_factory(int(plan[0][0]), int(plan[0][1]), variant)
            for M, fc, k in plan:
                d.configure(int(M), int(fc))
                ys.append(d.process(x[pos:pos + k], 16)[0])
                pos += int(k)
        else:
            u = int_factory(int(plan[0][0]))
            for M, k in plan:
                u.configure(int(M))
                ys.append(u.process(x[pos:pos + k]))
                pos += int(k)
        y = np.concatenate(ys)
        assert y.shape == want.shape, (base, y.shape, want.shape)
        if not np.array_equal(y, want):
            bad = np.nonzero((y != want).any(axis=1))[0]
            raise AssertionError(f"{base}: {len(bad)} samples differ from the reference build, first at {bad[:5]}")
        n_checked += 1
    assert n_checked == 8


def check_sink_golden(sink_factory):
    """sink_factory(F, tv_sec, tv_usec) -> object with .write(x) -> (n_frames, 128+F, 512)."""
    g = load("sink_ref.npz")
    for F in (4, 16):
        x, want = g[f"in_F{F}"], g[f"dgrams_F{F}"].reshape(2, 128 + F, 512)
        # the reference stamped block 0 with gettimeofday(): replay its time stamps
        got = []
        pos = 0
        for f in range(2):
            tv_sec = int.from_bytes(want[f, 0, 16:20].tobytes(), "little")
            tv_usec = int.from_bytes(want[f, 0, 20:24].tobytes(), "little")
            sk = sink_factory(F, tv_sec, tv_usec) if f == 0 else sk
            sk.set_time(tv_sec, tv_usec)
            out = sk.write(x[pos:pos + FRAME])
            pos += FRAME
            got.append(out)
        got = np.concatenate(got, axis=0)
        assert got.shape == want.shape
        # header filler byte of the recovery blocks is uninitialised memory in the reference
        # (UDPSinkFEC.cpp:233-243 never writes it): compare with it masked
        a, b = got.copy(), want.copy()
        a[:, 128:, 3] = 0
        b[:, 128:, 3] = 0
        assert np.array_equal(a, b), f"F={F}"
        for f in range(2):
            assert int.from_bytes(want[f, 0, 24:28].tobytes(), "little") == zlib.crc32(want[f, 0, 4:24].tobytes())


def check_fecbuffer_golden(decode):
    """decode(superblocks (n,512)) -> payload (127,508)."""
    g = load("fecbuffer_ref.npz")
    frames, payload = g["frames"], g["payload"]
    for f in range(7):
        sel = g[f"sel_{f}"]
        got = decode(frames[f][sel])
        assert np.array_equal(got.reshape(-1), payload[f]), f"frame {f}"
