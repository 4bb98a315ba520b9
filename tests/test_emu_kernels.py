"""Kernel logic and library host logic, checked on the CPU through the host emulation build of the
very same sources (tests/emu): parity with the oracle at small sizes.  The GPU tests repeat these
cases (and larger ones) on the product library."""
import numpy as np
import pytest

import cases


@pytest.mark.parametrize("M", [1, 2, 3, 4, 5, 6])
def test_decimator_centre_eo1(emu_lib, oracle, M):
    rng = np.random.default_rng(100 + M)
    n = 30000
    x = cases.rand_iq(rng, (2, n))
    cases.check_decimator(emu_lib, oracle, M, 2, 0, x, [0, 4096, 4096 + 12345, n])


@pytest.mark.parametrize("M,fcpos,variant,bits", [(0, 2, 0, 12), (0, 2, 0, 16), (1, 0, 0, 16), (1, 1, 1, 8), (2, 0, 0, 12),
                                                  (2, 1, 0, 16), (3, 0, 0, 16), (4, 1, 1, 12), (6, 0, 1, 8), (5, 2, 1, 16),
                                                  (4, 2, 1, 16), (6, 1, 0, 16)])
def test_decimator_variants(emu_lib, oracle, M, fcpos, variant, bits):
    rng = np.random.default_rng(200 + 10 * M + fcpos)
    n = 24000
    x = cases.rand_iq(rng, (1, n), bits)
    cases.check_decimator(emu_lib, oracle, M, fcpos, variant, x, [0, 777, 10000, n], bits)


@pytest.mark.parametrize("M,S,n", [(4, 3, 20011), (5, 3, 40003), (6, 3, 80001)])
def test_decimator_shares_cross_streams(emu_lib, oracle, M, S, n):
    """a warp's share of the global event axis ends one stream and starts the next (own warm-up per piece)"""
    rng = np.random.default_rng(250 + M)
    x = cases.rand_iq(rng, (S, n))
    cases.check_decimator(emu_lib, oracle, M, 2, M & 1, x, [0, n])


def test_decimator_input_classes(emu_lib, oracle):
    rng = np.random.default_rng(300)
    for name, x in cases.input_classes(rng, 16384).items():
        cases.check_decimator(emu_lib, oracle, 4, 2, 0, x[None], [0, 5000, 16384])


def test_decimator_short_and_empty(emu_lib, oracle):
    rng = np.random.default_rng(301)
    x = cases.rand_iq(rng, (1, 1000))
    cases.check_decimator(emu_lib, oracle, 6, 2, 0, x, [0, 10, 10, 70, 135, 1000])  # below one group, empty, ragged


@pytest.mark.parametrize("variant", [0, 1])
def test_decimator_reconfigure(emu_lib, oracle, variant):
    """stage states persist across Downsampler::configure (the verified round-1 mismatch)"""
    rng = np.random.default_rng(310 + variant)
    for plan in cases.dec_reconfigure_plans():
        n = sum(k for _, _, k in plan)
        cases.check_decimator_reconfigure(emu_lib, oracle, variant, cases.rand_iq(rng, (2, n)), plan)


def test_reconfigure_golden(emu_lib):
    """the sequences recorded from the reference build (tests/golden/reconfigure_ref.npz)"""
    import golden_cases
    from sdrdaemon_b200 import capi

    golden_cases.check_reconfigure_golden(lambda M, fc, v: capi.Decimator(M, fc, v, max_in=1 << 15, lib=emu_lib),
                                          lambda M: capi.Interpolator(M, max_in=1024, lib=emu_lib))


def test_interpolator_reconfigure(emu_lib, oracle):
    rng = np.random.default_rng(320)
    for plan in cases.int_reconfigure_plans():
        n = sum(k for _, k in plan)
        cases.check_interpolator_reconfigure(emu_lib, oracle, cases.rand_iq(rng, (2, n)), plan)


@pytest.mark.parametrize("M", [0, 1, 2, 3, 4, 5, 6])
def test_interpolator(emu_lib, oracle, M):
    rng = np.random.default_rng(800 + M)
    n = 3000 if M <= 4 else 700
    x = cases.rand_iq(rng, (2, n))
    cases.check_interpolator(emu_lib, oracle, M, x, [0, 1, 1, 40, 1300, n])  # single sample, empty, shorter than the history


def test_interpolator_many_streams_segments(emu_lib, oracle):
    rng = np.random.default_rng(850)
    for M, S, n in [(4, 300, 200), (5, 40, 64 * 5 + 1), (2, 1900, 70)]:
        cases.check_interpolator(emu_lib, oracle, M, cases.rand_iq(rng, (S, n)), [0, n // 3, n])


def test_interpolator_golden_and_classes(emu_lib, oracle):
    import golden_cases
    from sdrdaemon_b200 import capi

    golden_cases.check_interpolator_golden(lambda M: capi.Interpolator(M, max_in=2048, lib=emu_lib))
    rng = np.random.default_rng(801)
    for name, x in cases.input_classes(rng, 2048).items():
        cases.check_interpolator(emu_lib, oracle, 4, x[None], [0, 777, 2048])


@pytest.mark.parametrize("F", [0, 1, 16, 32, 40])
def test_sink_framing_and_encode(emu_lib, oracle, F):
    rng = np.random.default_rng(400 + F)
    n = cases.FRAME * 2 + 700
    x = cases.rand_iq(rng, (2, n))
    cases.check_sink(emu_lib, oracle, F, x, [0, 1000, 1000 + cases.FRAME, n])


def test_sink_many_frames_per_call(emu_lib, oracle):
    """More frames in one call than there are SMs: the encode kernel then runs in its two-CTAs-per-SM shape
    (fec::EncShape<true>: 256 threads, one image, 16 columns per warp) -- same datagrams."""
    rng = np.random.default_rng(477)
    S = 151
    x = cases.rand_iq(rng, (S, cases.FRAME + 300))
    cases.check_sink(emu_lib, oracle, 3, x, [0, 200, cases.FRAME + 300])


def test_cm256_encode_raw(emu_lib, oracle):
    from sdrdaemon_b200 import capi

    rng = np.random.default_rng(500)
    o = rng.integers(0, 256, size=(2, 128, 508), dtype=np.uint8)
    for F in (1, 7, 32):
        r = capi.cm256_encode(o, F, lib=emu_lib)
        want = np.stack([oracle.cm256_encode(o[f], F) for f in range(2)])
        assert np.array_equal(r, want)
    with pytest.raises(capi.SdrdError):
        capi.cm256_encode(o, 0, lib=emu_lib)
    with pytest.raises(capi.SdrdError):
        capi.cm256_encode(o, 129, lib=emu_lib)


def test_decode_all_branches(emu_lib, oracle):
    rng = np.random.default_rng(600)
    F = 40
    x, frames = cases.make_frames(oracle, rng, 12, F)
    sel = cases.erasure_cases(rng, frames, F)
    sb, nb = cases.pack_received(frames, sel)
    pay, b0, st = cases.check_decode(emu_lib, oracle, sb, nb)
    assert list(st) == [1, 0, 2, 2, 2, 2, 2, 1, -1, 1, 2, 2]
    # in-order recoveries return the transmitted samples
    for f in (0, 2, 4, 6, 10, 11):
        assert np.array_equal(pay[f].reshape(-1), x[f * cases.FRAME:(f + 1) * cases.FRAME].view(np.uint8).reshape(-1))


def test_receiver_batched(emu_lib, oracle):
    rng = np.random.default_rng(650)
    x, dg = cases.receiver_traffic(oracle, rng, 9, 8)
    n = len(dg)
    assert cases.check_receiver(emu_lib, oracle, dg, [0, n]) >= 10           # one burst
    assert cases.check_receiver(emu_lib, oracle, dg, [0, 1, 1, 50, 137, 300, 301, 700, n]) >= 10  # ragged bursts, empty burst


def test_golden_through_emulation(emu_lib):
    """the golden vectors through the emulated library as well (host logic + kernel indexing)"""
    import golden_cases
    from sdrdaemon_b200 import capi

    class D:
        def __init__(self, M, fc, v):
            self.d = capi.Decimator(M, fc, v, max_in=8192, lib=emu_lib)

        def process(self, x, bits):
            return self.d.process(x, bits)
    golden_cases.check_decimator_golden(D)
    golden_cases.check_fecbuffer_golden(lambda sb: capi.fec_decode(sb[None], [len(sb)], lib=emu_lib)[0][0])


def test_rx_pipeline_sliced(emu_lib, oracle):
    """sdrd_rx_process in 8 overlapped slices (what large calls do) gives the same datagrams"""
    test_rx_pipeline(emu_lib, oracle, slice_bytes=1)


def test_rx_pipeline(emu_lib, oracle, slice_bytes=0):
    from sdrdaemon_b200 import capi

    rng = np.random.default_rng(700)
    M, F, S = 2, 16, 2
    n = (cases.FRAME + 300) << M
    x = cases.rand_iq(rng, (S, 2 * n))
    rx = capi.Rx(M, n_streams=S, max_in=n, n_fec=F, lib=emu_lib)
    rx.set_slice_bytes(slice_bytes)
    got = np.concatenate([rx.process(x[:, :n]), rx.process(x[:, n:])], axis=1)
    for s in range(S):
        y, _ = oracle.Decimator(M).process(x[s])
        sk = oracle.Sink(n_fec=F)
        sk.write(y)
        assert np.array_equal(got[s], np.stack(sk.frames))


@pytest.mark.parametrize("M,bits", [(0, 8), (0, 12), (2, 8), (4, 8), (4, 12), (6, 12)])
def test_rx_pipeline_sample_bits(emu_lib, oracle, M, bits):
    """8- / 12-bit sources through the fused path (sample_bits was hard-coded to 16)"""
    rng = np.random.default_rng(720 + M + bits)
    n = (cases.FRAME + 50) << M
    cases.check_rx_sample_bits(emu_lib, oracle, M, bits, cases.rand_iq(rng, (2, n), bits))


def test_refused_calls_leave_state(emu_lib, oracle):
    cases.check_refused_calls_leave_state(emu_lib, oracle)


def test_rescale_keeps_filter_state(emu_lib, oracle):
    """Downsampler::rescale (static decimate1) between two process calls must not disturb the cascade"""
    from sdrdaemon_b200 import capi

    rng = np.random.default_rng(730)
    x = cases.rand_iq(rng, (1, 6000), 12)
    d = capi.Decimator(3, max_in=6000, lib=emu_lib)
    o = oracle.Decimator(3)
    y1, _ = d.process(x[:, :3000], 12)
    z = x[0, :500].copy()
    ss = cases.C.c_uint(12)
    emu_lib.check(emu_lib.sdrd_dec_rescale(d._h, z.ctypes.data, 500, 500, cases.C.byref(ss)))
    assert np.array_equal(z, (x[0, :500].astype(np.int32) << 4).astype(np.int16)) and ss.value == 12
    y2, _ = d.process(x[:, 3000:], 12)
    assert np.array_equal(y1[0], o.process(x[0, :3000], 12)[0]) and np.array_equal(y2[0], o.process(x[0, 3000:], 12)[0])


def test_sink_per_frame_time_stamps(emu_lib, oracle):
    F = cases.FRAME
    cases.check_sink_frame_clock(emu_lib, oracle, [0, 100, 100 + 3 * F, 100 + 3 * F + 50, 5 * F + 7, 6 * F + 7])


def test_rx_queued(emu_lib, oracle):
    cases.check_rx_queued(emu_lib, oracle, M=2, F=4, S=2, blk=4096, n_blk=24)
    cases.check_rx_queued(emu_lib, oracle, M=4, F=8, S=1, blk=65536, n_blk=9, max_blocks=2, bits=12)
    cases.check_rx_queued(emu_lib, oracle, M=1, F=0, S=1, blk=8192, n_blk=10, threaded=True)
    cases.check_rx_queued(emu_lib, oracle, M=3, F=8, S=2, blk=65536, n_blk=12, max_blocks=4, helpers=2)
    cases.check_rx_queued_mixed(emu_lib, oracle)
    cases.check_rx_queued_reconfigure(emu_lib, oracle)


def test_ipc_feed_entry_points(emu_lib, oracle):
    """sdrd_dec_ipc_export / sdrd_ipc_open / sdrd_ipc_copy_rows: another owner of the samples writes them straight into the
    handle's input buffer, then the device-resident form runs (in the emulation the 'peer' is this process)"""
    import ctypes as C

    from sdrdaemon_b200 import capi

    rng = np.random.default_rng(909)
    S, n, M = 3, 8192, 3
    x = cases.rand_iq(rng, (S, n))
    d = capi.Decimator(M, n_streams=S, max_in=n, lib=emu_lib)
    handle = (C.c_ubyte * 64)()
    off, stride = C.c_size_t(0), C.c_size_t(0)
    emu_lib.check(emu_lib.sdrd_dec_ipc_export(d._h, handle, C.byref(off), C.byref(stride)))
    peer = C.c_void_p()
    emu_lib.check(emu_lib.sdrd_ipc_open(handle, C.byref(peer)))
    emu_lib.check(emu_lib.sdrd_ipc_copy_rows(peer.value + off.value, stride.value * 4, x.ctypes.data, n * 4, n * 4, S, None))
    n_out, ss = d.process_dev(n)
    out_ptr, out_stride = d.dev_output()
    y = np.ctypeslib.as_array(C.cast(out_ptr, C.POINTER(C.c_int16)), shape=(S, out_stride, 2))[:, :n_out]
    for s in range(S):
        assert np.array_equal(y[s], oracle.Decimator(M).process(x[s])[0])
    emu_lib.check(emu_lib.sdrd_ipc_close(peer))
