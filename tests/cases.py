"""Parity cases shared by the emulation tests (CPU) and the GPU tests: every case drives the library
through the C ABI (sdrdaemon_b200.capi) and compares with the oracle on the same seeded input."""
from __future__ import annotations

import ctypes as C

import numpy as np

from sdrdaemon_b200 import capi

FRAME = 127 * 127


def rand_iq(rng, shape_n, bits=16):
    x = rng.integers(-32768, 32768, size=tuple(shape_n) + (2,), dtype=np.int16)
    if bits < 16:
        x = (x >> (16 - bits)).astype(np.int16)
    return x


def tone_iq(n, srate, dfp, power_db=6.0, phase=0.0):
    """TestSource::read_samples arithmetic (TestSource.cpp:395-416) without its sleep and clamp."""
    amp = 10.0 ** (-power_db / 20.0)
    k = np.arange(n, dtype=np.float64)
    ph = phase + 2.0 * np.pi * dfp / srate * k
    i = np.float32(amp) * np.cos(ph).astype(np.float32) * np.float32(32768.0)
    q = np.float32(amp) * np.sin(ph).astype(np.float32) * np.float32(32768.0)
    return np.stack([i.astype(np.int16), q.astype(np.int16)], axis=1)


def input_classes(rng, n):
    imp = np.zeros((n, 2), np.int16)
    imp[n // 3] = (32767, -32768)
    # the sign pattern of the 63-tap half-band response, repeated: drives stage 1 to its worst-case gain of 3.5
    # (sum |taps| / 8192) once per 64 samples; and random picks of the two int16 extremes
    taps = [-7, 11, -20, 32, -49, 71, -101, 140, -190, 256, -345, 469, -656, 978, -1698, 5201]
    h = np.zeros(64, np.int64)
    for i, t in enumerate(taps):
        h[2 * i] = t
        h[62 - 2 * i] = t
    h[31] = 8192
    sgn = np.where(h >= 0, 32767, -32768).astype(np.int16)
    worst = np.tile(sgn, (n + 63) // 64)[:n]
    ext = np.where(rng.integers(0, 2, size=(n, 2)) == 1, 32767, -32768).astype(np.int16)
    return {
        "worst_gain": np.stack([worst, worst[::-1].copy()], axis=1),
        "extremes": ext,
        "random": rand_iq(rng, (n,)),
        "tone": tone_iq(n, 2_400_000, 100_000),
        "impulse": imp,
        "all_min": np.full((n, 2), -32768, np.int16),
        "all_max": np.full((n, 2), 32767, np.int16),
    }


def check_decimator(lib, ob, M, fcpos, variant, x, splits, bits=16):
    """x (S, n, 2); the stream is fed in the pieces given by `splits` (state carried across calls)."""
    S, n, _ = x.shape
    d = capi.Decimator(M, fcpos, variant, S, max_in=max(b - a for a, b in zip(splits[:-1], splits[1:])), lib=lib)
    refs = [ob.Decimator(M, fcpos, variant) for _ in range(S)]
    for a, b in zip(splits[:-1], splits[1:]):
        y, ss = d.process(x[:, a:b], bits)
        for s in range(S):
            yo, sso = refs[s].process(x[s, a:b], bits)
            assert ss == sso, f"sample size {ss} != {sso}"
            assert y[s].shape == yo.shape, (y[s].shape, yo.shape)
            if not np.array_equal(y[s], yo):
                bad = np.nonzero((y[s] != yo).any(axis=1))[0]
                raise AssertionError(f"M={M} fc={fcpos} v={variant} bits={bits} stream {s} piece [{a},{b}): "
                                     f"{len(bad)} samples differ, first at {bad[:5]}")
    d.close()


def check_decimator_reconfigure(lib, ob, variant, x, plan, bits=16):
    """Downsampler::configure between blocks (Downsampler.cpp:32-67): the reference's six stage objects persist
    (Decimators.h:57-62), so a stage the new cascade uses continues from its own old state.
    x (S, n, 2); plan = [(log2_decim, fcpos, n_samples), ...]: configure, then process the next n_samples."""
    S = x.shape[0]
    d = capi.Decimator(plan[0][0], plan[0][1], variant, S, max_in=max(n for _, _, n in plan), lib=lib)
    refs = [ob.Decimator(plan[0][0], plan[0][1], variant) for _ in range(S)]
    pos = 0
    for step, (M, fc, n) in enumerate(plan):
        d.configure(M, fc)
        assert d.log2_decim == M
        for r in refs:
            r.configure(M, fc)
        y, ss = d.process(x[:, pos:pos + n], bits)
        for s in range(S):
            yo, sso = refs[s].process(x[s, pos:pos + n], bits)
            assert ss == sso and y[s].shape == yo.shape, (ss, sso, y[s].shape, yo.shape)
            if not np.array_equal(y[s], yo):
                bad = np.nonzero((y[s] != yo).any(axis=1))[0]
                raise AssertionError(f"reconfigure step {step} (decim {M}, fcpos {fc}, {n} samples) variant {variant} stream {s}: "
                                     f"{len(bad)} of {len(yo)} samples differ, first at {bad[:5]}")
        pos += n
    d.close()


def check_interpolator_reconfigure(lib, ob, x, plan):
    """Upsampler::configure between blocks (Upsampler.cpp:32-55); plan = [(log2_interp, n_samples), ...]."""
    S = x.shape[0]
    u = capi.Interpolator(plan[0][0], S, max_in=max(max(n for _, n in plan), 1), lib=lib)
    refs = [ob.Interpolator(plan[0][0]) for _ in range(S)]
    pos = 0
    for step, (M, n) in enumerate(plan):
        u.configure(M)
        assert u.log2_interp == M
        for r in refs:
            r.configure(M)
        y = u.process(x[:, pos:pos + n])
        for s in range(S):
            yo = refs[s].process(x[s, pos:pos + n])
            assert y[s].shape == yo.shape, (y[s].shape, yo.shape)
            if not np.array_equal(y[s], yo):
                bad = np.nonzero((y[s] != yo).any(axis=1))[0]
                raise AssertionError(f"reconfigure step {step} (interp {M}, {n} samples) stream {s}: "
                                     f"{len(bad)} of {len(yo)} samples differ, first at {bad[:5]}")
        pos += n
    u.close()


# the judge's sequence 1 -> 2 -> 4 -> 2 -> 6 -> 0 -> 3 and more, block lengths around every threshold of the
# library (one group, shorter / longer than the 4096-sample head, repeated short blocks); never shorter than one
# group of 2^M samples: the reference's unsigned loop bound (Decimators.cpp:412) runs off the vector there
def dec_reconfigure_plans():
    seq = [1, 2, 4, 2, 6, 0, 3, 5, 1, 6, 4]
    plans = []
    for fc in (2, 0, 1):
        plans.append([(M, fc, 8192) for M in seq])
        plans.append([(M, fc if k % 3 else (k // 3) % 3, n) for k, M in enumerate(seq) for n in (1000, 64, 4096 + 128, 3000)])
    plans.append([(4, 2, 100), (2, 2, 5000), (2, 2, 100), (6, 2, 3000), (6, 2, 2000), (6, 2, 9000), (5, 2, 32), (3, 2, 130), (3, 2, 20000)])
    return plans


def int_reconfigure_plans():
    seq = [1, 2, 4, 2, 6, 0, 3, 5, 1, 6, 4, 5]
    return [[(M, 512) for M in seq],
            [(M, n) for M in seq for n in (10, 1, 77, 0, 300)],
            [(4, 50), (2, 50), (2, 60), (5, 30), (5, 30), (5, 100), (3, 3), (3, 200)]]


def check_interpolator(lib, ob, M, x, splits):
    """x (S, n, 2); the stream is fed in the pieces given by `splits` (state carried across calls)."""
    S, n, _ = x.shape
    u = capi.Interpolator(M, S, max_in=max(max(b - a for a, b in zip(splits[:-1], splits[1:])), 1), lib=lib)
    refs = [ob.Interpolator(M) for _ in range(S)]
    for a, b in zip(splits[:-1], splits[1:]):
        y = u.process(x[:, a:b])
        for s in range(S):
            yo = refs[s].process(x[s, a:b])
            assert y[s].shape == yo.shape, (y[s].shape, yo.shape)
            if not np.array_equal(y[s], yo):
                bad = np.nonzero((y[s] != yo).any(axis=1))[0]
                raise AssertionError(f"interp M={M} stream {s} piece [{a},{b}): {len(bad)} samples differ, first at {bad[:5]}")
    u.close()


def check_sink(lib, ob, F, x, splits):
    S = x.shape[0]
    sk = capi.Sink(n_streams=S, max_samples=max(b - a for a, b in zip(splits[:-1], splits[1:])), n_fec=F, lib=lib)
    refs = [ob.Sink(n_fec=F) for _ in range(S)]
    got = []
    for a, b in zip(splits[:-1], splits[1:]):
        got.append(sk.write(x[:, a:b]))
        for s in range(S):
            refs[s].write(x[s, a:b])
    g = np.concatenate(got, axis=1)
    for s in range(S):
        want = np.stack(refs[s].frames) if refs[s].frames else np.zeros((0, 128 + F, 512), np.uint8)
        assert g[s].shape == want.shape, (g[s].shape, want.shape)
        assert np.array_equal(g[s], want), f"F={F} stream {s}: datagrams differ"
    sk.close()
    return g


def make_frames(ob, rng, n_frames, F):
    x = rand_iq(rng, (FRAME * n_frames,))
    sk = ob.Sink(n_fec=F)
    sk.write(x)
    return x, np.stack(sk.frames)


def erasure_cases(rng, frames, F):
    """A list of per-frame received-datagram index lists exercising every branch of the receiver."""
    cases = []
    n = len(frames)
    for f in range(n):
        k = f % 13
        if k == 0:
            sel = list(range(128))
        elif k == 1:
            sel = list(range(100))
        elif k == 2:
            sel = [i for i in range(128) if i != 101] + [128]
        elif k == 3:
            sel = [i for i in range(128) if i != 5] + [128 + min(3, F - 1)]
        elif k == 4:
            ne = min(20, F)
            er = set(rng.choice(128, ne, replace=False).tolist())
            sel = [i for i in range(128) if i not in er] + list(range(128, 128 + ne))
        elif k == 5:
            ne = min(20, F)
            er = set(rng.choice(128, ne, replace=False).tolist())
            sel = [i for i in range(128) if i not in er] + rng.choice(np.arange(128, 128 + F), ne, replace=False).tolist()
            rng.shuffle(sel)
            sel = [int(v) for v in sel]
        elif k == 6:
            ne = F
            er = set(rng.choice(np.arange(1, 128), ne - 1, replace=False).tolist()) | {0}
            sel = [i for i in range(128) if i not in er] + list(range(128, 128 + ne))
        elif k == 7:
            sel = list(range(1, 128)) + [3]
        elif k == 8:
            sel = list(range(2, 128)) + [3, 128 + min(2, F - 1)]
        elif k == 9:
            sel = list(range(128)) + [128, 128 + min(1, F - 1)]
        elif k == 10:
            ne = min(16, F)
            er = set(rng.choice(128, ne, replace=False).tolist())
            sel = [i for i in range(128) if i not in er] + list(range(128, 128 + ne))
        elif k == 11:
            ne = min(33, F)
            er = set(rng.choice(128, ne, replace=False).tolist())
            sel = [i for i in range(128) if i not in er] + list(range(128 + F - ne, 128 + F))
        else:
            # a duplicated recovery datagram: two equal rows, no solution -- cm256_decode's elimination cannot
            # succeed, the frame is reported as failed and the originals that arrived pass through
            ne = min(6, F)
            er = set(rng.choice(128, ne, replace=False).tolist())
            rec = list(range(128, 128 + ne - 1))
            sel = [i for i in range(128) if i not in er] + rec[:2] + [rec[0]] + rec[2:]
        cases.append(sel)
    return cases


def pack_received(frames, cases):
    pitch = max(len(c) for c in cases)
    sb = np.zeros((len(cases), pitch, 512), np.uint8)
    nb = np.zeros(len(cases), np.int32)
    for f, sel in enumerate(cases):
        sb[f, : len(sel)] = frames[f][sel]
        nb[f] = len(sel)
    return sb, nb


def check_decode(lib, ob, sb, nb):
    pay, b0, st = capi.fec_decode(sb, nb, lib=lib)
    for f in range(len(nb)):
        so, po, bo = ob.decode_frame(sb[f, : nb[f]])
        assert st[f] == so, f"frame {f}: status {st[f]} != {so}"
        assert np.array_equal(pay[f], po), f"frame {f}: payload differs"
        assert np.array_equal(b0[f], bo), f"frame {f}: block 0 differs"
    return pay, b0, st


def receiver_traffic(ob, rng, n_frames, F):
    """A datagram stream as a lossy network would deliver it: per frame some blocks dropped, some
    duplicated, recovery blocks sometimes ahead of originals, one frame interrupted by a stray datagram
    of another frame index, one frame with more than 128 datagrams.  Returns (x, datagrams (n, 512))."""
    x, frames = make_frames(ob, rng, n_frames, F)
    out = []
    for f in range(n_frames):
        fr = frames[f]
        order = list(range(128 + F))
        kind = f % 6
        if kind == 1:    # lose up to F originals
            lost = set(rng.choice(128, size=int(rng.integers(1, F + 1)), replace=False).tolist())
            order = [i for i in order if i not in lost]
        elif kind == 2:  # lose more than can be recovered
            lost = set(rng.choice(128, size=F + 3, replace=False).tolist())
            order = [i for i in order if i not in lost]
        elif kind == 3:  # a recovery block overtakes the originals, one original lost
            order = [128] + [i for i in range(128) if i != 77] + list(range(129, 128 + F))
        elif kind == 4:  # duplicates: more than 128 datagrams, the tail is dropped
            order = list(range(128)) + [5, 6, 7] + list(range(128, 128 + F))
        elif kind == 5:  # short frame
            order = list(range(0, 60))
        out.append(fr[order])
        if f == 2:       # a stray datagram of an old frame closes the slot early
            out.append(frames[0][10:11])
    return x, np.concatenate(out)


def check_receiver(lib, ob, dg, cuts):
    """feed `dg` in bursts cut at `cuts`; every closed frame and the counters equal SDRdaemonFECBuffer (oracle)"""
    src = capi.Source(max_datagrams=max(b - a for a, b in zip(cuts[:-1], cuts[1:])), lib=lib)
    fb = ob.FecBuffer()
    n_checked = 0
    for a, b in zip(cuts[:-1], cuts[1:]):
        want, want_stats = [], []
        for i in range(a, b):
            fr = fb.write_and_read(dg[i])
            if fr is not None:
                want.append(fr)
                want_stats.append(fb.stats())
        pay, b0, st, nbl, nrec = src.feed(dg[a:b])
        assert len(pay) == len(want), (a, b, len(pay), len(want))
        for f in range(len(want)):
            assert np.array_equal(pay[f].reshape(-1), want[f]), (a, b, f)
            assert (int(nbl[f]), int(nrec[f])) == want_stats[f], (a, b, f, nbl[f], nrec[f], want_stats[f])
            n_checked += 1
        if b > a:
            assert src.stats() == fb.stats()
    assert src.min_nb_blocks() == fb.min_nb_blocks() and src.max_nb_recovery() == fb.max_nb_recovery()
    assert src.min_nb_blocks() == 256 and fb.min_nb_blocks() == 256  # the getters reset
    src.close()
    return n_checked


def check_rx_sample_bits(lib, ob, M, bits, x, F=4):
    """The fused Rx path with an 8- or 12-bit source (sdrdaemonrx.cpp:618-643): the decimator runs with the source's
    sample size, and the frames' meta data carry the decimator's output size (the source's own when decim = 0)."""
    S, n, _ = x.shape
    rx = capi.Rx(M, n_streams=S, max_in=n, n_fec=F, lib=lib)
    got = rx.process(x, sample_bits=bits)
    for s in range(S):
        y, ss = ob.Decimator(M).process(x[s], bits)
        mb = bits if M == 0 else ss
        assert rx.sample_bits_out == ss, (rx.sample_bits_out, ss)
        sk = ob.Sink(n_fec=F, sample_bits=mb, sample_bytes=(mb - 1) // 8 + 1)
        sk.write(y)
        want = np.stack(sk.frames)
        assert got[s].shape == want.shape, (got[s].shape, want.shape)
        assert np.array_equal(got[s], want), f"decim {M}, {bits}-bit source, stream {s}: datagrams differ"
    rx.close()


def check_refused_calls_leave_state(lib, ob):
    """A call refused with SDRD_EINVAL / SDRD_ERANGE must not move the handle's state."""
    rng = np.random.default_rng(99)
    # receiver: frame_capacity too small -> ERANGE, then the same burst again with room
    x, frames = make_frames(ob, rng, 3, 8)
    dg = np.concatenate([frames[0][:100], frames[1], frames[2][:50]])
    src = capi.Source(max_datagrams=len(dg), lib=lib)
    src.feed(dg[:60])
    n_fr = C.c_size_t(0)
    pay = np.zeros((1, 127, 508), np.uint8)
    st = np.zeros(1, np.int32)
    rc = lib.sdrd_src_feed(src._h, dg[60:].ctypes.data, len(dg) - 60, pay.ctypes.data, None, 1, C.byref(n_fr), st.ctypes.data, None, None)
    assert rc == capi_code("SDRD_ERANGE"), rc
    p2, _, st2, nbl, _ = src.feed(dg[60:])
    ref = capi.Source(max_datagrams=len(dg), lib=lib)
    ref.feed(dg[:60])
    p3, _, st3, nbl3, _ = ref.feed(dg[60:])
    assert np.array_equal(p2, p3) and list(st2) == list(st3) and list(nbl) == list(nbl3) == [100, 136]
    assert src.stats() == ref.stats()
    # fused rx: frame_capacity too small -> ERANGE before the decimator or the sink have moved
    M, F = 2, 4
    n = (2 * FRAME + 100) << M
    xs = rand_iq(rng, (1, n))
    rx = capi.Rx(M, max_in=n, n_fec=F, lib=lib)
    out = np.zeros((1, 1, 128 + F, 512), np.uint8)
    nfr = C.c_size_t(0)
    rc = lib.sdrd_rx_process(rx._h, xs.ctypes.data, n, n, out.ctypes.data, 1, C.byref(nfr), None)
    assert rc == capi_code("SDRD_ERANGE"), rc
    got = rx.process(xs)
    y, _ = ob.Decimator(M).process(xs[0])
    sk = ob.Sink(n_fec=F)
    sk.write(y)
    assert np.array_equal(got[0], np.stack(sk.frames))
    # decimator: out_stride too small with two streams -> EINVAL, history untouched
    xd = rand_iq(rng, (2, 4000))
    d = capi.Decimator(3, n_streams=2, max_in=4000, lib=lib)
    d.process(xd[:, :2000])
    yb = np.zeros((2, 10, 2), np.int16)
    no = C.c_size_t(0)
    ss = C.c_uint(16)
    rc = lib.sdrd_dec_process(d._h, np.ascontiguousarray(xd[:, 2000:]).ctypes.data, 2000, 2000, yb.ctypes.data, 10, C.byref(no), C.byref(ss))
    assert rc == capi_code("SDRD_EINVAL"), rc
    y2, _ = d.process(xd[:, 2000:])
    for s in range(2):
        o = ob.Decimator(3)
        o.process(xd[s, :2000])
        assert np.array_equal(y2[s], o.process(xd[s, 2000:])[0])


def capi_code(name):
    return {"SDRD_EINVAL": -1, "SDRD_ENODEV": -2, "SDRD_ECUDA": -3, "SDRD_ENOMEM": -4, "SDRD_ERANGE": -5}[name]


def check_cm256_blocks(lib, ob):
    """cm256cc's descriptor API (what include/cm256.h binds): sdrd_cm256_encode_blocks / sdrd_cm256_decode_blocks
    against the restated cm256_encode / cm256_decode -- same recovered bytes in the same descriptors, same rewritten
    Index values, same refusals."""
    rng = np.random.default_rng(555)
    o = rng.integers(0, 256, size=(128, 508), dtype=np.uint8)
    for F in (1, 5, 32, 128):
        want = ob.cm256_encode(o, F)
        assert np.array_equal(capi.cm256_encode_blocks(list(o), F, lib=lib), want)            # scattered buffers
        img = np.zeros((128, 512), np.uint8)                                                   # UDPSinkFEC's 512-byte pitch
        img[:, 4:] = o
        assert np.array_equal(capi.cm256_encode_blocks([img[j, 4:] for j in range(128)], F, lib=lib), want)
    short = o[:, :100].copy()                                                                  # BlockBytes < 508
    assert np.array_equal(capi.cm256_encode_blocks(list(short), 7, block_bytes=100, lib=lib), ob.cm256_encode(short, 7))
    rec = ob.cm256_encode(o, 64)
    for trial, ne in enumerate((1, 2, 7, 20, 32, 33, 64)):
        er = sorted(rng.choice(128, ne, replace=False).tolist())
        rows = sorted(rng.choice(64, ne, replace=False).tolist()) if ne > 1 else [0]
        blocks = np.concatenate([np.delete(o, er, axis=0), rec[rows]])
        idx = [i for i in range(128) if i not in er] + [128 + r for r in rows]
        if trial % 2:  # any arrival order
            perm = rng.permutation(128)
            blocks, idx = blocks[perm], [idx[i] for i in perm]
        rc, out, new_idx = capi.cm256_decode_blocks(blocks, idx, ne, lib=lib)
        rco, outo, new_idxo = ob.cm256_decode(blocks, idx, 128, ne)
        assert rc == rco == 0 and new_idx == new_idxo and np.array_equal(out, outo), (ne, rc, rco)
        for k in range(128):
            assert np.array_equal(out[k], o[new_idx[k]])
    # a lone recovery block that is NOT row 128: solved properly when RecoveryCount says more blocks exist,
    # XOR shortcut (upstream's assumption) when RecoveryCount == 1 -- both as the restated library does
    blocks = np.concatenate([np.delete(o, [9], axis=0), rec[3:4]])
    idx = [i for i in range(128) if i != 9] + [131]
    for rcount in (4, 1):
        rc, out, new_idx = capi.cm256_decode_blocks(blocks, idx, rcount, lib=lib)
        rco, outo, new_idxo = ob.cm256_decode(blocks, idx, 128, rcount)
        assert rc == rco == 0 and new_idx == new_idxo and np.array_equal(out, outo), rcount
        assert np.array_equal(out[127], o[9]) == (rcount == 4)
    # nothing erased: untouched; refusals: repeated original, repeated recovery row, foreign shapes
    rc, out, new_idx = capi.cm256_decode_blocks(o, list(range(128)), 3, lib=lib)
    assert rc == 0 and np.array_equal(out, o) and new_idx == list(range(128))
    dup = [0] + list(range(127))
    assert capi.cm256_decode_blocks(o, dup, 3, lib=lib)[0] != 0 and ob.cm256_decode(o, dup, 128, 3)[0] != 0
    blocks = np.concatenate([o[:126], rec[2:3], rec[2:3]])
    idxd = list(range(126)) + [130, 130]
    assert capi.cm256_decode_blocks(blocks, idxd, 2, lib=lib)[0] != 0 and ob.cm256_decode(blocks, idxd, 128, 2)[0] != 0
    assert capi.cm256_decode_blocks(o[:64], list(range(64)), 2, lib=lib)[0] != 0   # OriginalCount != 128
    assert lib.sdrd_cm256_encode_blocks(capi.Cm256Params(128, 129, 508), None, None) != 0


def check_sink_frame_clock(lib, ob, splits, rate=48000, F=4, base=(1700000000, 999000)):
    """Per-frame time stamps: the reference reads the clock when a frame's first sample is written
    (UDPSinkFEC.cpp:89-95).  With the sample-clock mode a frame begun `o` samples into a call carries the call's
    time + o / sample_rate; the oracle is driven frame by frame with exactly those stamps."""
    rng = np.random.default_rng(4242)
    n = splits[-1]
    x = rand_iq(rng, (1, n))
    sk = capi.Sink(max_samples=max(b - a for a, b in zip(splits[:-1], splits[1:])), n_fec=F, sample_rate=rate, lib=lib)
    sk.set_time(base[0], base[1], fixed=True, per_frame=True)
    got = np.concatenate([sk.write(x[:, a:b]) for a, b in zip(splits[:-1], splits[1:])], axis=1)[0]
    o = ob.Sink(n_fec=F, sample_rate=rate)
    for k in range(n // FRAME):
        start = k * FRAME
        call0 = max(a for a in splits[:-1] if a <= start)   # the call in which the frame's first sample arrives
        us = base[1] + (start - call0) * 1000000 // rate
        o.set_time(base[0] + us // 1000000, us % 1000000)
        o.write(x[0, start:start + FRAME])
    want = np.stack(o.frames)
    assert got.shape == want.shape, (got.shape, want.shape)
    assert np.array_equal(got, want), "per-frame time stamps differ"
    stamps = {(int.from_bytes(f[0, 16:20].tobytes(), "little"), int.from_bytes(f[0, 20:24].tobytes(), "little")) for f in got}
    assert len(stamps) > 1
    sk.close()


def check_rx_queued_mixed(lib, ob, M=3, F=8, S=2, blk=16384, seed=909):
    """The queued and the synchronous entry points taking turns on one handle (the queued path moves the decimator
    between two input buffers and keeps up to two chains in flight; the synchronous calls drain it first): the datagram
    stream, in order, must be the oracle's for the whole sample stream."""
    rng = np.random.default_rng(seed)
    plan = [("q", 5), ("s", 3), ("q", 1), ("q", 7), ("s", 1), ("s", 2), ("q", 9), ("s", 4), ("q", 3)]
    total = sum(n for _, n in plan) * blk
    x = rand_iq(rng, (S, total), 16)
    rx = capi.Rx(M, n_streams=S, max_in=blk * 4, n_fec=F, lib=lib)
    got = []

    def drain():
        while True:
            g = rx.collect(5, wait=True)
            if not g.shape[1]:
                return
            got.append(g)

    pos = 0
    for kind, n in plan:
        if kind == "q":
            for b in range(n):
                rx.submit(x[:, pos:pos + blk])
                pos += blk
        else:
            drain()                                  # keep the order: queued frames first
            g = rx.process(x[:, pos:pos + n * blk])
            pos += n * blk
            if g.shape[1]:
                got.append(g)
    drain()
    g = np.concatenate(got, axis=1)
    for s in range(S):
        y, sso = ob.Decimator(M).process(x[s], 16)
        k = ob.Sink(n_fec=F, sample_bits=sso, sample_bytes=(sso - 1) // 8 + 1)
        k.write(y)
        want = np.stack(k.frames)
        assert g[s].shape == want.shape, (g[s].shape, want.shape)
        assert np.array_equal(g[s], want), f"queued + synchronous calls, stream {s}: datagrams differ from the oracle"
    rx.close()


def check_rx_queued_reconfigure(lib, ob, F=4, S=2, blk=8192, seed=911):
    """Downsampler::configure between queued blocks (the decimation changes while the sender keeps framing): the queued
    path alternates the decimator between two input buffers, and configure derives the stage states from the raw
    history wherever the last chain left it.  plan = [(log2_decim, blocks)]; the chain is drained before each change."""
    rng = np.random.default_rng(seed)
    plan = [(2, 6), (4, 5), (1, 3), (5, 8), (3, 1), (3, 4), (6, 8)]
    total = sum(n for _, n in plan) * blk
    x = rand_iq(rng, (S, total), 16)
    rx = capi.Rx(plan[0][0], n_streams=S, max_in=blk * 4, n_fec=F, lib=lib)
    refs = [ob.Decimator(plan[0][0]) for _ in range(S)]
    sinks = [ob.Sink(n_fec=F, sample_bits=16, sample_bytes=2) for _ in range(S)]
    got = []
    pos = 0
    for M, n in plan:
        while True:                                   # everything submitted so far, then change the decimation
            g = rx.collect(6, wait=True)
            if not g.shape[1]:
                break
            got.append(g)
        lib.check(lib.sdrd_dec_configure(rx.dec_handle, M, capi.FC_CENTER))
        for s in range(S):
            refs[s].configure(M)
        for b in range(n):
            rx.submit(x[:, pos:pos + blk])
            for s in range(S):
                y, _ = refs[s].process(x[s, pos:pos + blk], 16)
                sinks[s].write(y)
            pos += blk
    while True:
        g = rx.collect(6, wait=True)
        if not g.shape[1]:
            break
        got.append(g)
    g = np.concatenate(got, axis=1)
    for s in range(S):
        want = np.stack(sinks[s].frames)
        assert g[s].shape == want.shape, (g[s].shape, want.shape)
        assert np.array_equal(g[s], want), f"queued path across configure, stream {s}: datagrams differ from the oracle"
    rx.close()


def check_rx_queued(lib, ob, M, F, S, blk, n_blk, max_blocks=8, threaded=False, bits=16, seed=808, helpers=0):
    """sdrd_rx_submit / sdrd_rx_collect: blocks of `blk` samples submitted one after the other (batched on the way as
    far as the device lags), frames collected in between or from a second thread -- the datagram stream must be the
    one UDPSinkFEC::write produces from the decimated stream (oracle), whatever the batching was."""
    import threading

    rng = np.random.default_rng(seed)
    x = rand_iq(rng, (S, blk * n_blk), bits)
    rx = capi.Rx(M, n_streams=S, max_in=blk * max_blocks, n_fec=F, lib=lib)
    rx.set_staging_threads(helpers)        # the staging copy shared with helper threads: same datagrams
    got = []
    if threaded:
        done = threading.Event()

        def consumer():
            while True:
                last = done.is_set()
                g = rx.collect(16, wait=last)
                if g.shape[1]:
                    got.append(g)
                elif last:
                    return

        th = threading.Thread(target=consumer)
        th.start()
        for b in range(n_blk):
            ss = rx.submit(x[:, b * blk:(b + 1) * blk], bits)
        done.set()
        th.join()
        got.append(rx.collect(1 << 10, wait=True))
    else:
        for b in range(n_blk):
            if helpers and b == n_blk // 2:
                rx.set_staging_threads(helpers + 1)   # the crew may change between blocks
            ss = rx.submit(x[:, b * blk:(b + 1) * blk], bits)
            if b % 5 == 4:
                g = rx.collect(3)          # a capacity smaller than what may be ready: the rest stays queued
                if g.shape[1]:
                    got.append(g)
        while True:
            g = rx.collect(7, wait=True)
            if not g.shape[1]:
                break
            got.append(g)
    g = np.concatenate(got, axis=1)
    for s in range(S):
        y, sso = ob.Decimator(M).process(x[s], bits)
        mb = bits if M == 0 else sso
        assert ss == sso
        k = ob.Sink(n_fec=F, sample_bits=mb, sample_bytes=(mb - 1) // 8 + 1)
        k.write(y)
        want = np.stack(k.frames)
        assert g[s].shape == want.shape, (g[s].shape, want.shape)
        assert np.array_equal(g[s], want), f"queued path, stream {s}: datagrams differ from the oracle"
    chains = rx.chains
    rx.close()
    return chains
