"""The drop-in claim at the library seam: the reference's UNMODIFIED sender and receiver sources
(sdmnbase/UDPSinkFEC.cpp, sdmnbase/SDRdaemonFECBuffer.cpp, gr-sdrdaemon/lib/SDRdaemonFECBuffer.cpp), compiled against
include/cm256.h and linked with the library under test, run end to end over loop-back UDP with the reference's own
erasure injection (-DSDRDAEMON_PUNCTURE=101) plus extra dropped blocks, and recover every frame -- with the same
output, byte for byte, as the same programs built against the CPU restatement of cm256.

The programs are built by `make -C oracle seam` (oracle/_ref/, where /root/reference exists) and travel to the GPU
box prebuilt, like the other reference-derived objects."""
import json
import os
import subprocess

import numpy as np
import pytest

import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")

# (n_fec, extra blocks dropped per frame): N = 1 + extra recovery blocks are needed per frame --
# cm256's single-block shortcut, a small system, BASELINE config 4's 20 erasures, and more than 32 (K3's large build)
COMBOS = [(1, 0), (8, 3), (32, 19), (40, 35)]


def have(kind):
    return all(os.path.exists(os.path.join(REFDIR, f"ref_seam_{p}_{kind}")) for p in ("loopback", "gr"))


def run_pair(kind, n_fec, extra, tmp_path, n_frames=5, seed=11):
    port = 20000 + (os.getpid() * 7 + n_fec * 13 + extra + {"oracle": 0, "emu": 1, "gpu": 2}[kind] * 101) % 20000
    cap = str(tmp_path / f"cap_{kind}_{n_fec}_{extra}.bin")
    r = subprocess.run([os.path.join(REFDIR, f"ref_seam_loopback_{kind}"), str(port), str(n_fec), str(n_frames), str(extra), str(seed), cap],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (kind, n_fec, extra, r.stdout, r.stderr[-2000:])
    a = json.loads(r.stdout.strip().splitlines()[-1])
    r = subprocess.run([os.path.join(REFDIR, f"ref_seam_gr_{kind}"), cap, str(n_frames), str(extra), str(seed), "101"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (kind, "gr", n_fec, extra, r.stdout, r.stderr[-2000:])
    b = json.loads(r.stdout.strip().splitlines()[-1])
    return a, b


def check(kind, tmp_path):
    for n_fec, extra in COMBOS:
        a, b = run_pair(kind, n_fec, extra, tmp_path)
        o, og = run_pair("oracle", n_fec, extra, tmp_path)
        for got, want in ((a, o), (b, og)):
            assert got["frames_checked"] >= 5 and got["frames_ok"] == got["frames_checked"], got
            assert got["decode_error"] == 0 and got["decode_success"] >= got["frames_checked"], got  # every frame went through cm256_decode
            assert got["digest"] == want["digest"], (got, want)
        assert a["datagrams"] == o["datagrams"] and a["datagrams"] >= 5 * (127 + n_fec)  # recovery blocks were produced and sent


@pytest.mark.skipif(not (have("oracle") and have("emu")), reason="oracle/_ref seam programs not built (make -C oracle seam)")
def test_reference_sources_over_the_seam_emulation(emu_lib, tmp_path):
    check("emu", tmp_path)


@pytest.mark.gpu
def test_reference_sources_over_the_seam_gpu(gpu_lib, tmp_path):
    assert have("gpu") and have("oracle"), "oracle/_ref seam programs missing: they are built where /root/reference exists and travel prebuilt"
    check("gpu", tmp_path)


def test_cm256_descriptor_api_emulation(emu_lib, oracle):
    cases.check_cm256_blocks(emu_lib, oracle)


@pytest.mark.gpu
def test_cm256_descriptor_api_gpu(gpu_lib, oracle):
    cases.check_cm256_blocks(gpu_lib, oracle)
