"""GPU parity tests: the product library (libsdrd_b200.so, sm_100a kernels) through its C ABI against
the oracle on the same seeded inputs -- bit-exact, integer/byte work throughout -- plus size-independent
properties at BASELINE.json's full sizes."""
import os
import zlib

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M", [1, 2, 3, 4, 5, 6])
def test_decimator_centre_eo1(gpu_lib, oracle, M):
    rng = np.random.default_rng(1000 + M)
    n = 1 << 20
    x = cases.rand_iq(rng, (3, n))
    cases.check_decimator(gpu_lib, oracle, M, 2, 0, x, [0, 65536, 65536 + 123456, 65536 + 123456 + 64 * 1001, n])


@pytest.mark.parametrize("M", [0, 1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("fcpos", [0, 1, 2])
@pytest.mark.parametrize("variant", [0, 1])
def test_decimator_all_modes(gpu_lib, oracle, M, fcpos, variant):
    rng = np.random.default_rng(2000 + 100 * M + 10 * fcpos + variant)
    for bits in (16, 12, 8):
        n = 150000
        x = cases.rand_iq(rng, (2, n), bits)
        cases.check_decimator(gpu_lib, oracle, M, fcpos, variant, x, [0, 777, 65536 + 777, n], bits)


def test_decimator_input_classes(gpu_lib, oracle):
    """SURVEY 8c pins: random full scale, TestSource CW, impulse, all-min (wrap-around) x M = 1..6 x split calls."""
    rng = np.random.default_rng(3000)
    for M in range(1, 7):
        for name, x in cases.input_classes(rng, 1 << 17).items():
            cases.check_decimator(gpu_lib, oracle, M, 2, 0, x[None], [0, 50000, 1 << 17])


def test_decimator_testsource_blocks(gpu_lib, oracle):
    """config 1: 2.4 Msps TestSource CW, decimate-by-2, fed in the reference's 65536-sample blocks."""
    x = cases.tone_iq(65536 * 6, 2_400_000, 100_000)
    cases.check_decimator(gpu_lib, oracle, 1, 2, 0, x[None], [65536 * k for k in range(7)])


def test_decimator_short_and_empty(gpu_lib, oracle):
    rng = np.random.default_rng(3001)
    x = cases.rand_iq(rng, (1, 1000))
    cases.check_decimator(gpu_lib, oracle, 6, 2, 0, x, [0, 10, 10, 70, 135, 1000])
    cases.check_decimator(gpu_lib, oracle, 1, 2, 0, x, [0, 1, 2, 3, 1000])


@pytest.mark.parametrize("variant", [0, 1])
def test_decimator_reconfigure(gpu_lib, oracle, variant):
    """Downsampler::configure mid-stream: the six stage objects persist (Decimators.h:57-62); sequences
    1 -> 2 -> 4 -> 2 -> 6 -> 0 -> 3 ... with all three fcpos, blocks shorter and longer than the stateful head."""
    rng = np.random.default_rng(3200 + variant)
    for plan in cases.dec_reconfigure_plans():
        n = sum(k for _, _, k in plan)
        cases.check_decimator_reconfigure(gpu_lib, oracle, variant, cases.rand_iq(rng, (3, n)), plan)
    # the reference's block size, many streams, the warp kernel taking over behind the head inside one call
    plan = [(4, 2, 65536), (2, 2, 65536), (5, 0, 65536), (5, 0, 65536), (6, 2, 65536), (1, 2, 65536), (3, 1, 65536)]
    cases.check_decimator_reconfigure(gpu_lib, oracle, variant, cases.rand_iq(rng, (40, 7 * 65536)), plan)
    for bits in (8, 12):
        plan = [(3, 2, 5000), (5, 2, 70000), (2, 2, 3000), (4, 1, 9000)]
        cases.check_decimator_reconfigure(gpu_lib, oracle, variant, cases.rand_iq(rng, (2, 87000), bits), plan, bits)


def test_interpolator_reconfigure(gpu_lib, oracle):
    rng = np.random.default_rng(3210)
    for plan in cases.int_reconfigure_plans():
        n = sum(k for _, k in plan)
        cases.check_interpolator_reconfigure(gpu_lib, oracle, cases.rand_iq(rng, (3, n)), plan)
    plan = [(4, 4096), (2, 4096), (5, 4096), (6, 1000), (1, 4096), (3, 130), (3, 4096)]
    cases.check_interpolator_reconfigure(gpu_lib, oracle, cases.rand_iq(rng, (40, sum(k for _, k in plan))), plan)


def test_decimator_many_streams_many_segments(gpu_lib, oracle):
    """64 streams laid end to end on the global event axis, the warps' shares cross stream boundaries; every
    stream against the oracle."""
    from sdrdaemon_b200 import capi

    rng = np.random.default_rng(3002)
    S, n, M = 64, 1 << 19, 5
    x = cases.rand_iq(rng, (S, n))
    d = capi.Decimator(M, n_streams=S, max_in=n, lib=gpu_lib)
    y, _ = d.process(x)
    for s in range(S):
        yo, _ = oracle.Decimator(M).process(x[s])
        assert np.array_equal(y[s], yo), f"stream {s}"
    # linearity-free property at full size: splitting the call must not change a single bit
    d.reset()
    y2 = np.concatenate([d.process(x[:, : n // 2 + 64])[0], d.process(x[:, n // 2 + 64:])[0]], axis=1)
    assert np.array_equal(y, y2)


@pytest.mark.parametrize("M,S,n", [(4, 7, 300_007), (5, 13, 250_003), (6, 5, 700_001), (1, 3, 100_001), (6, 300, 5000)])
def test_decimator_shares_cross_streams(gpu_lib, oracle, M, S, n):
    """ragged stream lengths (the last event of a stream is partial) with warps whose share of the event axis
    ends one stream and starts the next; (6, 300, 5000): more streams than one warp per stream would need."""
    rng = np.random.default_rng(3100 + M)
    x = cases.rand_iq(rng, (S, n))
    cases.check_decimator(gpu_lib, oracle, M, 2, M & 1, x, [0, n // 3, n])


def test_decimator_large_single_stream(gpu_lib, oracle):
    """config 2 shape: one 10 Msps stream, decimate-by-16, 64 superframes in one call."""
    from sdrdaemon_b200 import capi

    rng = np.random.default_rng(3003)
    n = 64 * cases.FRAME * 16
    x = cases.rand_iq(rng, (n,))
    d = capi.Decimator(4, max_in=n, lib=gpu_lib)
    y, _ = d.process(x)
    yo, _ = oracle.Decimator(4).process(x)
    assert np.array_equal(y, yo)


def test_full_bench_sizes_against_the_oracle(gpu_lib, oracle):
    """BASELINE config 2 at bench.py's full size (592 superframes, 152.8 M input samples), every sample against
    the oracle: decimate by 16 on the GPU, then the Tx cascade back up by 16 (152.8 M output samples)."""
    from sdrdaemon_b200 import capi

    rng = np.random.default_rng(3004)
    n = 592 * cases.FRAME * 16
    x = rng.integers(-32768, 32768, size=(n, 2), dtype=np.int16)
    d = capi.Decimator(4, max_in=n, lib=gpu_lib)
    y, _ = d.process(x)
    d.close()
    yo, _ = oracle.Decimator(4).process(x)
    assert y.shape == yo.shape == (592 * cases.FRAME, 2) and np.array_equal(y, yo)
    del x
    # all 592 superframes, 128 + 16 datagrams each, against UDPSinkFEC::write + cm256_encode (oracle)
    sk = capi.Sink(max_samples=len(y), n_fec=16, lib=gpu_lib)
    dg = sk.write(y)
    sk.close()
    osk = oracle.Sink(n_fec=16)
    osk.write(yo)
    assert dg.shape == (592, 144, 512) and np.array_equal(dg, np.stack(osk.frames))
    del dg, osk
    u = capi.Interpolator(4, max_in=len(y), lib=gpu_lib)
    z = u.process(y)
    u.close()
    zo = oracle.Interpolator(4).process(yo)
    assert z.shape == (n, 2) and np.array_equal(z, zo)


@pytest.mark.parametrize("M", [0, 1, 2, 3, 4, 5, 6])
def test_interpolator(gpu_lib, oracle, M):
    rng = np.random.default_rng(900 + M)
    n = 40000 if M <= 4 else 9000
    x = cases.rand_iq(rng, (3, n))
    cases.check_interpolator(gpu_lib, oracle, M, x, [0, 1, 1, 40, 4096 + 17, n])  # single sample, empty, ragged tiles


@pytest.mark.parametrize("M,S,n", [(4, 300, 1000), (5, 40, 64 * 37 + 1), (1, 7, 64 * 600 - 1), (3, 2000, 130)])
def test_interpolator_many_streams_segments(gpu_lib, oracle, M, S, n):
    """the warp-private K4: warps take contiguous ranges of 64-sample steps of one stream (each with a warm-up
    step); few long, many short and ragged streams, state carried over two calls"""
    rng = np.random.default_rng(950 + M)
    x = cases.rand_iq(rng, (S, n))
    cases.check_interpolator(gpu_lib, oracle, M, x, [0, n // 3, n])


def test_interpolator_input_classes_and_reconfigure(gpu_lib, oracle):
    rng = np.random.default_rng(901)
    for name, x in cases.input_classes(rng, 8192).items():
        cases.check_interpolator(gpu_lib, oracle, 4, x[None], [0, 3333, 8192])
    # decimate by 16 then interpolate by 16: the Rx and Tx cascades back to back keep a slow tone
    from sdrdaemon_b200 import capi

    n = 1 << 16
    t = np.arange(n)
    tone = np.stack([np.round(8000 * np.cos(2 * np.pi * t / 4096.0)), np.round(8000 * np.sin(2 * np.pi * t / 4096.0))],
                    axis=1).astype(np.int16)
    y, _ = capi.Decimator(4, max_in=n, lib=gpu_lib).process(tone)
    z = capi.Interpolator(4, max_in=len(y), lib=gpu_lib).process(y)
    zo = oracle.Interpolator(4).process(oracle.Decimator(4).process(tone)[0])
    assert np.array_equal(z, zo)
    d = 794  # group delay of the two cascades in input samples (measured with the oracle)
    err = np.abs(z[4096:n - 4096].astype(np.int32) - tone[4096 - d:n - 4096 - d].astype(np.int32)).max()
    assert err < 100, err


@pytest.mark.parametrize("F", [0, 1, 16, 20, 32, 40, 128])
def test_sink_framing_and_encode(gpu_lib, oracle, F):
    rng = np.random.default_rng(4000 + F)
    n = cases.FRAME * 5 + 700
    x = cases.rand_iq(rng, (3, n))
    cases.check_sink(gpu_lib, oracle, F, x, [0, 1000, 1000 + cases.FRAME, 3 * cases.FRAME + 5, n])


@pytest.mark.parametrize("F", [1, 16, 33])
def test_sink_many_frames_per_call(gpu_lib, oracle, F):
    """More frames in one call than there are SMs: the two-CTAs-per-SM shape of the encode kernel (fec::EncShape<true>);
    149 .. 2 x 148 + 1 work items cover its grid rounding."""
    rng = np.random.default_rng(4770 + F)
    for S, nfr in ((149, 1), (99, 3), (150, 2)):
        x = cases.rand_iq(rng, (S, nfr * cases.FRAME + 300))
        cases.check_sink(gpu_lib, oracle, F, x, [0, 200, nfr * cases.FRAME + 300])


def test_sink_known_answers(gpu_lib):
    """Framing pins of SURVEY 8c(3): 16129 samples per frame, header bytes, CRC-32 of the meta data."""
    import zlib
    from sdrdaemon_b200 import capi

    x = np.arange(2 * cases.FRAME * 2, dtype=np.int16).reshape(-1, 2)
    sk = capi.Sink(n_fec=8, center_freq_khz=435000, sample_rate=625000, tv_sec=1700000000, tv_usec=12, lib=gpu_lib)
    fr = sk.write(x)
    assert fr.shape == (2, 136, 512)
    for f in range(2):
        assert (fr[f, :, 0].astype(int) | (fr[f, :, 1].astype(int) << 8) == f).all()  # frameIndex
        assert (fr[f, :, 2] == np.arange(136)).all() and (fr[f, :, 3] == 0).all()       # blockIndex, filler
        meta = fr[f, 0, 4:28].tobytes()
        assert int.from_bytes(meta[0:4], "little") == 435000 and int.from_bytes(meta[4:8], "little") == 625000
        assert meta[8:12] == bytes([2, 16, 128, 8])
        assert int.from_bytes(meta[12:16], "little") == 1700000000 and int.from_bytes(meta[16:20], "little") == 12
        assert int.from_bytes(meta[20:24], "little") == zlib.crc32(meta[:20])
        assert not fr[f, 0, 28:].any()
        assert np.array_equal(fr[f, 1:128, 4:].reshape(-1), x[f * cases.FRAME:(f + 1) * cases.FRAME].view(np.uint8).reshape(-1))
        # recovery row 0 of a Cauchy code with x_0 = 128 is the XOR parity of the 128 originals
        assert np.array_equal(fr[f, 128, 4:], np.bitwise_xor.reduce(fr[f, :128, 4:], axis=0))


def test_cm256_encode_raw_and_linearity(gpu_lib, oracle):
    from sdrdaemon_b200 import capi

    rng = np.random.default_rng(5000)
    a = rng.integers(0, 256, size=(6, 128, 508), dtype=np.uint8)
    b = rng.integers(0, 256, size=(6, 128, 508), dtype=np.uint8)
    for F in (1, 16, 32, 33):
        ra = capi.cm256_encode(a, F, lib=gpu_lib)
        assert np.array_equal(ra, np.stack([oracle.cm256_encode(a[f], F) for f in range(6)]))
        rb = capi.cm256_encode(b, F, lib=gpu_lib)
        assert np.array_equal(capi.cm256_encode(a ^ b, F, lib=gpu_lib), ra ^ rb)  # GF(2^8)-linear
    with pytest.raises(capi.SdrdError):
        capi.cm256_encode(a, 0, lib=gpu_lib)


@pytest.mark.parametrize("F", [4, 32, 40])
def test_decode_all_branches(gpu_lib, oracle, F):
    rng = np.random.default_rng(6000 + F)
    x, frames = cases.make_frames(oracle, rng, 24, F)
    sel = cases.erasure_cases(rng, frames, F)
    sb, nb = cases.pack_received(frames, sel)
    cases.check_decode(gpu_lib, oracle, sb, nb)


def test_decode_mds_property(gpu_lib, oracle):
    """Any 128 of the 128+F blocks (delivered originals first) give the frame back."""
    from sdrdaemon_b200 import capi

    rng = np.random.default_rng(6100)
    F = 32
    x, frames = cases.make_frames(oracle, rng, 64, F)
    sel = []
    for f in range(64):
        ne = int(rng.integers(1, F + 1))
        er = set(rng.choice(128, ne, replace=False).tolist())
        rec = sorted(rng.choice(np.arange(128, 128 + F), ne, replace=False).tolist())
        sel.append([i for i in range(128) if i not in er] + rec)
    sb, nb = cases.pack_received(frames, sel)
    pay, b0, st = capi.fec_decode(sb, nb, lib=gpu_lib)
    for f in range(64):
        ne = sum(1 for v in sel[f] if v >= 128)
        if ne == 1 and sel[f][-1] != 128:
            continue  # cm256's single-block shortcut assumes row 128 (documented quirk)
        assert st[f] == 2
        assert np.array_equal(pay[f], frames[f, 1:128, 4:]) and np.array_equal(b0[f], frames[f, 0, 4:])


def test_config4_decode_4096_frames(gpu_lib, oracle):
    """BASELINE config 4: 4096 superframes, 20 of 128 blocks erased, F = 32: encode on the GPU, erase,
    recover on the GPU, compare with what was sent (round trip) and spot-check frames against the oracle."""
    from sdrdaemon_b200 import capi

    rng = np.random.default_rng(0xFEC0)
    nf, F = 4096, 32
    x = cases.rand_iq(rng, (1, nf * cases.FRAME))
    sk = capi.Sink(max_samples=nf * cases.FRAME, n_fec=F, lib=gpu_lib)
    frames = sk.write(x)[0]
    assert frames.shape == (nf, 160, 512)
    sb = np.zeros((nf, 128, 512), np.uint8)
    for f in range(nf):
        er = rng.permutation(128)[:20]
        keep = np.ones(128, bool)
        keep[er] = False
        sb[f, :108] = frames[f, :128][keep]
        sb[f, 108:] = frames[f, 128:148]
    pay, b0, st = capi.fec_decode(sb, 128, lib=gpu_lib)
    assert (st == 2).all()
    assert np.array_equal(pay, frames[:, 1:128, 4:])
    assert np.array_equal(b0, frames[:, 0, 4:])
    # every one of the 4096 frames through the oracle's SDRdaemonFECBuffer + cm256_decode (threaded), not a sample
    po, bo, so = oracle.decode_frames(sb, 128, n_threads=os.cpu_count() or 1)
    assert (so == 2).all() and np.array_equal(po, pay) and np.array_equal(bo, b0)


def test_receiver_batched(gpu_lib, oracle):
    rng = np.random.default_rng(651)
    for F in (8, 40):
        x, dg = cases.receiver_traffic(oracle, rng, 13, F)
        n = len(dg)
        assert cases.check_receiver(gpu_lib, oracle, dg, [0, n]) >= 13
        assert cases.check_receiver(gpu_lib, oracle, dg, [0, 3, 3, 129, 500, 501, 1200, n]) >= 13


def test_rx_pipeline_sliced(gpu_lib, oracle):
    """the overlapped (sliced) form of sdrd_rx_process, forced on a small input, DB variant, 3 streams"""
    from sdrdaemon_b200 import capi

    rng = np.random.default_rng(710)
    M, F, S = 3, 8, 3
    n = (2 * cases.FRAME + 1234) << M
    x = cases.rand_iq(rng, (S, 2 * n))
    rx = capi.Rx(M, n_streams=S, max_in=n, n_fec=F, variant=1, lib=gpu_lib)
    rx.set_slice_bytes(1)
    got = np.concatenate([rx.process(x[:, :n]), rx.process(x[:, n:])], axis=1)
    for s in range(S):
        y, _ = oracle.Decimator(M, 2, 1).process(x[s])
        sk = oracle.Sink(n_fec=F)
        sk.write(y)
        assert np.array_equal(got[s], np.stack(sk.frames))


def test_rx_pipeline_config2(gpu_lib, oracle):
    """config 2 end to end through sdrd_rx_process: decimate-by-16 + 128+16 FEC, ragged call sizes."""
    from sdrdaemon_b200 import capi

    rng = np.random.default_rng(7000)
    M, F = 4, 16
    n = (3 * cases.FRAME + 1234) << M
    x = cases.rand_iq(rng, (n,))
    rx = capi.Rx(M, max_in=n, n_fec=F, lib=gpu_lib)
    cut = (cases.FRAME // 2) << M
    got = np.concatenate([rx.process(x[:cut]), rx.process(x[cut:])], axis=0)
    y, _ = oracle.Decimator(M).process(x)
    sk = oracle.Sink(n_fec=F)
    sk.write(y)
    assert np.array_equal(got, np.stack(sk.frames))


def test_rx_pipeline_config3_shape(gpu_lib, oracle):
    """config 3 shape at reduced length: 256 streams, decimate-by-32, 128+32 FEC; streams spot-checked."""
    from sdrdaemon_b200 import capi

    rng = np.random.default_rng(7001)
    M, F, S = 5, 32, 256
    n = cases.FRAME << M
    x = cases.rand_iq(rng, (S, n))
    rx = capi.Rx(M, n_streams=S, max_in=n, n_fec=F, lib=gpu_lib)
    got = rx.process(x)
    assert got.shape == (S, 1, 160, 512)
    for s in (0, 100, 255):
        y, _ = oracle.Decimator(M).process(x[s])
        sk = oracle.Sink(n_fec=F)
        sk.write(y)
        assert np.array_equal(got[s], np.stack(sk.frames))
    # encode -> erase 20 -> decode round trip on every stream
    sb = np.concatenate([got[:, 0, 20:128], got[:, 0, 128:148]], axis=1)
    pay, b0, st = capi.fec_decode(sb, 128, lib=gpu_lib)
    assert (st == 2).all() and np.array_equal(pay, got[:, 0, 1:128, 4:])


@pytest.mark.parametrize("M,bits", [(0, 8), (0, 12), (1, 8), (2, 12), (4, 8), (4, 12), (5, 8), (6, 12)])
def test_rx_pipeline_sample_bits(gpu_lib, oracle, M, bits):
    """8- / 12-bit sources (RTL-SDR, Airspy) through the fused path: the decimator runs with the source's sample size
    and the meta data carry its output size (sdrdaemonrx.cpp:618-643)"""
    rng = np.random.default_rng(7200 + M + bits)
    n = (2 * cases.FRAME + 50) << M
    cases.check_rx_sample_bits(gpu_lib, oracle, M, bits, cases.rand_iq(rng, (3, n), bits))


def test_refused_calls_leave_state(gpu_lib, oracle):
    cases.check_refused_calls_leave_state(gpu_lib, oracle)


def test_rescale_keeps_filter_state(gpu_lib, oracle):
    from sdrdaemon_b200 import capi

    rng = np.random.default_rng(7300)
    x = cases.rand_iq(rng, (1, 200000), 8)
    d = capi.Decimator(5, max_in=100000, lib=gpu_lib)
    o = oracle.Decimator(5)
    y1, _ = d.process(x[:, :100000], 8)
    z = x[0, :70000].copy()
    ss = cases.C.c_uint(8)
    gpu_lib.check(gpu_lib.sdrd_dec_rescale(d._h, z.ctypes.data, len(z), len(z), cases.C.byref(ss)))
    assert np.array_equal(z, (x[0, :70000].astype(np.int32) << 8).astype(np.int16)) and ss.value == 8
    y2, _ = d.process(x[:, 100000:], 8)
    assert np.array_equal(y1[0], o.process(x[0, :100000], 8)[0]) and np.array_equal(y2[0], o.process(x[0, 100000:], 8)[0])


def test_sink_per_frame_time_stamps(gpu_lib, oracle):
    """one time stamp per frame (UDPSinkFEC.cpp:89-95) for callers that batch many frames into one write"""
    F = cases.FRAME
    cases.check_sink_frame_clock(gpu_lib, oracle, [0, 100, 100 + 3 * F, 100 + 3 * F + 50, 5 * F + 7, 6 * F + 7])
    cases.check_sink_frame_clock(gpu_lib, oracle, [0, 40 * F + 11, 41 * F, 90 * F + 5], rate=625000, F=16)


def _full_size_rx(gpu_lib, oracle, M, F, S, frames, seed, batch):
    """S streams x `frames` superframes of full-scale random int16 I/Q through decimate + frame + encode on the GPU
    (`batch` streams per call, which keeps the host buffers bounded), EVERY stream compared with the oracle through
    the CRC-32 of its datagram bytes in send order; the oracle runs all streams on all host cores."""
    from sdrdaemon_b200 import capi

    n = (frames * cases.FRAME) << M
    threads = os.cpu_count() or 1
    rx = capi.Rx(M, n_streams=batch, max_in=n, n_fec=F, lib=gpu_lib)
    out = np.zeros((batch, frames, 128 + F, 512), np.uint8)
    checked = 0
    for s0 in range(0, S, batch):
        rng = np.random.default_rng(seed + s0)
        x = rng.integers(-32768, 32768, size=(batch, n, 2), dtype=np.int16)
        rx.reset()
        got = rx.process(x, out=out)
        assert got.shape == (batch, frames, 128 + F, 512)
        nfr, crc = oracle.rx_stream_crcs(x, M, F, threads)
        assert nfr == batch * frames
        for s in range(batch):
            assert zlib.crc32(got[s].tobytes()) == int(crc[s]), f"stream {s0 + s}: datagrams differ from the oracle"
            checked += 1
        # frame counter and sample content are stream-independent state: spot-check one stream byte for byte as well
        y, _ = oracle.Decimator(M).process(x[batch - 1])
        sk = oracle.Sink(n_fec=F)
        sk.write(y)
        assert np.array_equal(got[batch - 1], np.stack(sk.frames))
    assert checked == S
    rx.close()


def test_full_size_config3_every_stream(gpu_lib, oracle):
    """BASELINE config 3 at full size: 256 streams x 8 superframes, decimate-by-32, 128 + 32 FEC (1.06 G input
    samples), every stream against the oracle."""
    _full_size_rx(gpu_lib, oracle, M=5, F=32, S=256, frames=8, seed=30000, batch=64)


def test_full_size_config5_shard_every_stream(gpu_lib, oracle):
    """BASELINE config 5, one GPU's shard at full size: 256 streams x 2 superframes, decimate-by-64, 128 + 32 FEC
    (528 M input samples), every stream against the oracle."""
    _full_size_rx(gpu_lib, oracle, M=6, F=32, S=256, frames=2, seed=50000, batch=128)


def test_rx_queued(gpu_lib, oracle):
    """the queued form at the reference's block size: submit never waits for the device, blocks are batched on the way"""
    n_blk = 64
    chains = cases.check_rx_queued(gpu_lib, oracle, M=4, F=16, S=1, blk=65536, n_blk=n_blk, max_blocks=16)
    assert 1 <= chains <= n_blk + 1
    cases.check_rx_queued(gpu_lib, oracle, M=2, F=4, S=3, blk=4096, n_blk=100)
    cases.check_rx_queued(gpu_lib, oracle, M=5, F=32, S=2, blk=65536, n_blk=20, max_blocks=3, bits=8)
    cases.check_rx_queued(gpu_lib, oracle, M=4, F=16, S=1, blk=65536, n_blk=64, max_blocks=16, threaded=True)
    cases.check_rx_queued(gpu_lib, oracle, M=0, F=8, S=1, blk=16384, n_blk=30, bits=12)
    cases.check_rx_queued(gpu_lib, oracle, M=4, F=16, S=1, blk=65536, n_blk=96, max_blocks=16, helpers=2)
    cases.check_rx_queued(gpu_lib, oracle, M=3, F=8, S=2, blk=65536, n_blk=24, max_blocks=4, helpers=3, threaded=True)
    cases.check_rx_queued_mixed(gpu_lib, oracle)
    cases.check_rx_queued_reconfigure(gpu_lib, oracle)
    cases.check_rx_queued_mixed(gpu_lib, oracle, M=4, F=16, S=1, blk=65536, seed=910)
