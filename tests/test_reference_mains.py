"""The reference's own main programs, compiled UNMODIFIED (sdrdaemonrx.cpp, sdrdaemontx.cpp) against the B200 host layer
(sdrdaemon_b200/host/compat/ in front of the reference's include/; glue types are the reference's, compute classes this
library's) and run as processes:

    sdrdaemonrx -t test -c srate=..,decim=..,fecblk=..  -I 127.0.0.1 -D port     TestSource -> Downsampler -> UDPSinkFEC
    sdrdaemontx -t file -c file=..,interp=..            -I 127.0.0.1 -D port     UDPSourceFEC -> Upsampler -> FileSink

The datagrams / the .sdriq file are compared with what the oracle computes for the same samples.  Binaries are built by
`make -C oracle mains` where /root/reference exists (oracle/_ref/, git-ignored) and travel to the GPU box prebuilt."""
import os
import signal
import socket
import subprocess
import time

import numpy as np
import pytest

import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
FRAME = 127 * 127


def exe(name, kind):
    return os.path.join(REFDIR, f"{name}_{kind}")


def have(kind):
    return os.path.exists(exe("sdrdaemonrx", kind)) and os.path.exists(exe("sdrdaemontx", kind))


def tone_blocks(n_blocks, blklen, srate, dfp, power_db):
    """TestSource::read_samples block after block: float phasor carried across blocks (TestSource.cpp:395-416)"""
    amp = np.float32(10.0 ** (-power_db / 20.0))
    dphi = np.float32(2.0 * np.pi * dfp / srate)
    out = np.zeros((n_blocks * blklen, 2), np.int16)
    ph = np.float32(0.0)
    two_pi = 2.0 * np.pi
    hw = np.float32(32768.0)
    for i in range(n_blocks * blklen):
        out[i, 0] = np.int16(np.float32(amp * np.float32(np.cos(np.float64(ph)))) * hw)
        out[i, 1] = np.int16(np.float32(amp * np.float32(np.sin(np.float64(ph)))) * hw)
        ph = np.float32(ph + dphi)
        if ph > two_pi:
            ph = np.float32(ph - two_pi)
        elif ph < two_pi:
            ph = np.float32(ph + two_pi)
    return out


def check_rx_main(kind, oracle, port):
    decim, fecblk, srate, blklen, n_frames = 2, 8, 1_000_000, 16384, 5
    sock = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    sock.setsockopt(socket.SOL_SOCKET, socket.SO_RCVBUF, 32 << 20)
    sock.bind(("127.0.0.1", port))
    sock.settimeout(20.0)
    cfg = f"srate={srate},freq=435000000,decim={decim},fecblk={fecblk},txdelay=0,dfp=100000,power=6,blklen={blklen}"
    p = subprocess.Popen([exe("sdrdaemonrx", kind), "-t", "test", "-c", cfg, "-I", "127.0.0.1", "-D", str(port)],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    got = {}
    try:
        while True:
            dg = sock.recv(2048)
            if len(dg) != 512:
                continue
            fi, bi = dg[0] | (dg[1] << 8), dg[2]
            if fi >= n_frames:
                break
            got[(fi, bi)] = np.frombuffer(dg, np.uint8)
    finally:
        p.send_signal(signal.SIGINT)
        try:
            _, err = p.communicate(timeout=30)
        except subprocess.TimeoutExpired:
            p.kill()
            _, err = p.communicate()
        sock.close()
    assert p.returncode == 0, err.decode()[-2000:]
    assert len(got) >= 0.9 * n_frames * (128 + fecblk), (len(got), err.decode()[-1000:])
    # what the reference computes from the same source: blocks of blklen, the first block thrown away (sdrdaemonrx.cpp:646-648)
    need_out = n_frames * FRAME
    n_blocks = need_out // (blklen >> decim) + 3
    from oracle import bindings as ob
    if ob.ref_available(0):
        x, _ = ob.ref_testsource(n_blocks * blklen, srate, float(np.float32(2.0 * np.pi * 100000 / srate)), float(np.float32(10.0 ** (-6 / 20.0))))
    else:
        x = tone_blocks(n_blocks, blklen, srate, 100000, 6.0)
    dec = oracle.Decimator(decim)
    ys = [dec.process(x[b * blklen:(b + 1) * blklen])[0] for b in range(n_blocks)][1:]
    sk = oracle.Sink(center_freq_khz=435000, sample_rate=srate >> decim, n_fec=fecblk)
    sk.write(np.concatenate(ys))
    want = np.stack(sk.frames)
    bad = 0
    for (fi, bi), dg in got.items():
        w = want[fi, bi].copy()
        g = dg.copy()
        if bi == 0 or bi >= 128:  # wall-clock time stamp + its CRC in block 0 (datagram bytes 16..27), and what they encode to
            w[16:28] = 0
            g[16:28] = 0
        bad += not np.array_equal(g, w)
    assert bad == 0, f"{bad} of {len(got)} datagrams differ from Downsampler + UDPSinkFEC (oracle)"


def check_tx_main(kind, oracle, port, tmp_path):
    interp, n_fec, n_frames = 2, 8, 4
    out = str(tmp_path / "tx.sdriq")
    cfg = f"file={out},srate=250000,freq=435000000,interp={interp},stamp=1700000000"
    p = subprocess.Popen([exe("sdrdaemontx", kind), "-t", "file", "-c", cfg, "-I", "127.0.0.1", "-D", str(port)],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    rng = np.random.default_rng(91)
    x, frames = cases.make_frames(oracle, rng, n_frames + 1, n_fec)
    sock = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    # the program binds its socket, then FileSink::start of the reference mirror sleeps: wait for the port, then for that
    hexport = f":{port:04X} "
    for _ in range(1200):
        try:
            if any(hexport in line for line in open("/proc/net/udp")):
                break
        except OSError:
            break
        if p.poll() is not None:
            break
        time.sleep(0.05)
    time.sleep(2.5)
    try:
        for f in range(n_frames + 1):
            for b in range(128 + n_fec):
                if b == 77:  # one original lost per frame: recovered through cm256_decode
                    continue
                sock.sendto(frames[f][b].tobytes(), ("127.0.0.1", port))
                if f == 0 and b == 0:
                    # the first datagram makes the receiver emit its (empty) initial slot: the first GPU work of the
                    # process -- context creation and module load, seconds on a freshly started box.  Nothing else is
                    # sent until that has reached the file, so no datagram waits in a socket buffer meanwhile.
                    for _ in range(1800):
                        if (os.path.exists(out) and os.path.getsize(out) > 20) or p.poll() is not None:
                            break
                        time.sleep(0.05)
                if b % 16 == 0:
                    time.sleep(0.002)
        # wait until the frames have gone through (the emulation library decodes slowly)
        want_bytes = 20 + (((n_frames + 1) * FRAME * 4) << interp) - 65536  # the stream buffer holds the tail back
        for _ in range(600):
            if os.path.exists(out) and os.path.getsize(out) >= want_bytes:
                break
            time.sleep(0.05)
        p.send_signal(signal.SIGINT)
        for k in range(20):  # UDPSourceFEC::read blocks in recv: datagrams of a further frame let the loop see the flag
            d = frames[0][k % 128].copy()
            d[0:2] = (0xFF, 0x7F)
            sock.sendto(d.tobytes(), ("127.0.0.1", port))
            time.sleep(0.01)
            if p.poll() is not None:
                break
        _, err = p.communicate(timeout=30)
    finally:
        if p.poll() is None:
            p.kill()
        sock.close()
    assert p.returncode == 0, err.decode()[-2000:]
    raw = np.fromfile(out, dtype=np.uint8)
    assert int.from_bytes(raw[0:4].tobytes(), "little") == 250000 and int.from_bytes(raw[4:12].tobytes(), "little") == 435000000
    got = raw[20:].view(np.int16).reshape(-1, 2)
    # the receiver's first emission is its empty initial slot (zeros), then the frames in order
    up = oracle.Interpolator(interp)
    want = np.concatenate([up.process(np.zeros((FRAME, 2), np.int16))] + [up.process(x[f * FRAME:(f + 1) * FRAME]) for f in range(n_frames)])
    assert len(got) >= len(want), (len(got), len(want), err.decode()[-1500:])
    assert np.array_equal(got[:len(want)], want), "the .sdriq stream differs from UDPSourceFEC + Upsampler (oracle)"


@pytest.mark.skipif(not have("emu"), reason="oracle/_ref mains not built (make -C oracle mains)")
def test_reference_rx_main_emulation(emu_lib, oracle):
    check_rx_main("emu", oracle, 21000 + os.getpid() % 5000)


@pytest.mark.skipif(not have("emu"), reason="oracle/_ref mains not built (make -C oracle mains)")
def test_reference_tx_main_emulation(emu_lib, oracle, tmp_path):
    check_tx_main("emu", oracle, 27000 + os.getpid() % 5000, tmp_path)


@pytest.mark.gpu
def test_reference_rx_main_gpu(gpu_lib, oracle):
    assert have("gpu"), "oracle/_ref mains missing: built where /root/reference exists, they travel prebuilt"
    check_rx_main("gpu", oracle, 33000 + os.getpid() % 5000)


@pytest.mark.gpu
def test_reference_tx_main_gpu(gpu_lib, oracle, tmp_path):
    assert have("gpu"), "oracle/_ref mains missing: built where /root/reference exists, they travel prebuilt"
    check_tx_main("gpu", oracle, 39000 + os.getpid() % 5000, tmp_path)
