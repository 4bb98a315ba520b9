"""The C-ABI boundary on a CPU-only box: the product library loads, exports every symbol the header
declares, and refuses to run without a device (no CPU fallback)."""
import os
import re
import subprocess

import pytest

from sdrdaemon_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "sdrd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sdrd_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(n for n, _, _ in capi.SYMBOLS)


def test_product_library_exports_every_symbol():
    from sdrdaemon_b200 import build

    so = build.build_library()
    out = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (sdrd_[a-z0-9_]+)", out))
    missing = [s for s in header_symbols() if s not in exported]
    assert not missing, missing
    lib = capi.load()
    assert b"sm_100a" in lib.sdrd_version()


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = capi.load()
    assert lib.sdrd_device_count() == 0
    with pytest.raises(capi.SdrdError) as e:
        capi.Decimator(4, lib=lib)
    assert e.value.code == -2  # SDRD_ENODEV
    with pytest.raises(capi.SdrdError):
        capi.Sink(lib=lib)


def test_argument_errors(emu_lib):
    for bad in [dict(log2_decim=7), dict(log2_decim=-1), dict(log2_decim=2, fcpos=3), dict(log2_decim=2, variant=2)]:
        with pytest.raises(capi.SdrdError) as e:
            capi.Decimator(lib=emu_lib, **bad)
        assert e.value.code == -1
    d = capi.Decimator(2, max_in=100, lib=emu_lib)
    import numpy as np

    with pytest.raises(capi.SdrdError) as e:
        d.process(np.zeros((101, 2), np.int16))
    assert e.value.code == -5
    with pytest.raises(capi.SdrdError) as e:
        d.configure(9, 2)
    assert "Invalid log2 decimation factor" in e.value.msg  # the reference's message, Downsampler.cpp:41
    s = capi.Sink(lib=emu_lib)
    with pytest.raises(capi.SdrdError):
        s.set_nb_fec(129)
    for bad in (-1, 7):
        with pytest.raises(capi.SdrdError) as e:
            capi.Interpolator(bad, lib=emu_lib)
        assert e.value.code == -1 and "Invalid log2 interpolation factor" in e.value.msg  # Upsampler.cpp:40
    u = capi.Interpolator(3, max_in=10, lib=emu_lib)
    with pytest.raises(capi.SdrdError) as e:
        u.process(np.zeros((11, 2), np.int16))
    assert e.value.code == -5
