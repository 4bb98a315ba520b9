"""C++ host layer: CPU check through the emulation library, GPU check through libsdrd_b200.so."""
import pytest

import host_cases
from oracle import bindings as ob


def test_testsource_matches_reference(tmp_path):
    if not ob.ref_available(0):
        pytest.skip("reference build (oracle/_ref) not present")
    host_cases.check_testsource(host_cases.build("emu"), ob, str(tmp_path))


def test_host_pipeline_emulated(tmp_path):
    out = host_cases.check_pipeline(host_cases.build("emu"), ob, str(tmp_path), port=19311, decim=2, fecblk=8, n_blocks=4,
                                    blklen=32768)
    assert "frames_received=" in out


def test_host_upsampler_emulated(tmp_path):
    host_cases.check_upsampler(host_cases.build("emu"), ob, str(tmp_path), interp=3, block=700, n=3000)


@pytest.mark.gpu
def test_host_upsampler_gpu(tmp_path):
    for interp in (1, 4, 6):
        host_cases.check_upsampler(host_cases.build("gpu"), ob, str(tmp_path), interp=interp)


@pytest.mark.gpu
def test_host_pipeline_gpu(tmp_path):
    out = host_cases.check_pipeline(host_cases.build("gpu"), ob, str(tmp_path), port=19312, decim=4, fecblk=16, n_blocks=20)
    assert "frames_recovered=0" in out


@pytest.mark.gpu
def test_host_pipeline_gpu_punctured(tmp_path):
    """SDRDAEMON_PUNCTURE 101 (UDPSinkFEC.cpp:27,261-265): block 101 never sent, every frame recovered."""
    out = host_cases.check_pipeline(host_cases.build("gpu"), ob, str(tmp_path), port=19313, decim=3, fecblk=8, n_blocks=12,
                                    puncture=101)
    assert "frames_recovered=0" not in out
