"""The C++ host layer (sdrdaemon_b200/host/sdrd_host.hpp) driven by tests/host/host_pipeline.cpp, linked
against the emulation library (CPU tests) or the CUDA library (GPU tests), checked against the oracle."""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FRAME = 127 * 127


def build(kind: str) -> str:
    """kind: 'emu' or 'gpu' -> path of the test binary."""
    exe = os.path.join(ROOT, "tests", "host", f"host_pipeline_{kind}")
    src = os.path.join(ROOT, "tests", "host", "host_pipeline.cpp")
    hdr = os.path.join(ROOT, "sdrdaemon_b200", "host", "sdrd_host.hpp")
    if kind == "emu":
        libdir, lib = os.path.join(ROOT, "tests", "emu"), "sdrd_emu"
        subprocess.run(["make", "-s", "-C", libdir], check=True)
    else:
        libdir, lib = os.path.join(ROOT, "sdrdaemon_b200"), "sdrd_b200"
    so = os.path.join(libdir, f"lib{lib}.so")
    if (not os.path.exists(exe)) or any(os.path.getmtime(p) > os.path.getmtime(exe) for p in (src, hdr, so)):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([cxx, "-std=c++17", "-O2", "-pthread", "-o", exe, src, f"-L{libdir}", f"-l{lib}",
                        f"-Wl,-rpath,{libdir}"], check=True)
    return exe


def tone(n, srate, dfp, power_db, exe, tmp):
    out = os.path.join(tmp, "tone.raw")
    subprocess.run([exe, "testsource", str(n), str(srate), str(dfp), str(power_db), out], check=True)
    return np.fromfile(out, dtype=np.int16).reshape(-1, 2)


def check_testsource(exe, ob, tmp):
    """TestSource::read_samples mirror vs the reference's own function (needs oracle/_ref)."""
    srate, dfp, power = 48000, 1234.5, 6.0
    x = tone(4096, srate, dfp, power, exe, tmp)
    amp = np.float32(10.0 ** (-power / 20.0))
    dphi = np.float32(2.0 * np.pi * dfp / srate)
    ref, _ = ob.ref_testsource(4096, srate, float(dphi), float(amp))
    assert np.array_equal(x, ref)


def check_pipeline(exe, ob, tmp, port, decim=2, fecblk=8, n_blocks=5, blklen=65536, puncture=-1, srate=2400000):
    sdriq = os.path.join(tmp, "out.sdriq")
    dgbin = os.path.join(tmp, "dgrams.bin")
    cfg = f"srate={srate},decim={decim},fecblk={fecblk},dfp=100000,power=6,blklen={blklen}"
    cmd = [exe, "pipeline", str(port), cfg, str(n_blocks), sdriq, dgbin] + ([str(puncture)] if puncture >= 0 else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    # what the reference would have produced from the same source samples
    x = tone(n_blocks * blklen, srate, 100000, 6.0, exe, tmp)
    dec = ob.Decimator(decim)
    outs = []
    for b in range(n_blocks):
        y, _ = dec.process(x[b * blklen:(b + 1) * blklen])
        if b > 0:  # sdrdaemonrx.cpp:646-648 discards the first block
            outs.append(y)
    y = np.concatenate(outs)
    pad = np.zeros(((FRAME - len(y) % FRAME) + FRAME, 2), np.int16)
    sk = ob.Sink(center_freq_khz=435000, sample_rate=srate >> decim, n_fec=fecblk, tv_sec=1700000000, tv_usec=0)
    sk.write(y)
    sk.write(pad)
    want = np.stack(sk.frames)  # (n_frames, 128+F, 512)
    got = np.fromfile(dgbin, dtype=np.uint8).reshape(-1, 512)
    if puncture >= 0:
        keep = [i for i in range(128 + fecblk) if i != puncture]
        want_sent = want[:, keep].reshape(-1, 512)
    else:
        want_sent = want.reshape(-1, 512)
    assert got.shape == want_sent.shape, (got.shape, want_sent.shape)
    assert np.array_equal(got, want_sent), "datagrams differ from UDPSinkFEC (oracle)"
    # the .sdriq file: 20-byte header, then every frame the receiver handed over
    raw = np.fromfile(sdriq, dtype=np.uint8)
    assert int.from_bytes(raw[0:4].tobytes(), "little") == 150000
    assert int.from_bytes(raw[4:12].tobytes(), "little") == 435000000
    assert int.from_bytes(raw[12:20].tobytes(), "little") == 1700000000
    samples = raw[20:].view(np.int16).reshape(-1, 2)
    n_frames_rx = len(samples) // FRAME
    n_full = len(y) // FRAME
    assert n_frames_rx >= n_full, (n_frames_rx, n_full, r.stdout)
    assert np.array_equal(samples[: n_full * FRAME], y[: n_full * FRAME]), "received stream differs from the decimated stream"
    return r.stdout


def check_upsampler(exe, ob, tmp, interp=4, block=3000, n=10000):
    """host Upsampler (configure + process, block by block) vs the oracle"""
    rng = np.random.default_rng(55 + interp)
    x = rng.integers(-32768, 32768, size=(n, 2), dtype=np.int16)
    fin, fout = os.path.join(tmp, "up_in.raw"), os.path.join(tmp, "up_out.raw")
    x.tofile(fin)
    r = subprocess.run([exe, "upsample", str(interp), str(block), fin, fout], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stdout + r.stderr)
    got = np.fromfile(fout, dtype=np.int16).reshape(-1, 2)
    want = ob.Interpolator(interp).process(x)
    assert got.shape == want.shape and np.array_equal(got, want)
