"""pytest configuration: the `gpu` marker and the two libraries the parity cases run against.

  - `-m "not gpu"`: the oracle against the reference build / golden vectors, the host logic, the
    C-ABI export check, and the kernels' logic through the host emulation (tests/emu).
  - `-m gpu`: the parity tests proper, through the C ABI of libsdrd_b200.so on a B200.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) device")


@pytest.fixture(scope="session")
def oracle():
    from oracle import bindings as ob

    ob.lib()
    return ob


@pytest.fixture(scope="session")
def emu_lib():
    """The C-ABI library compiled for the host (sdrd_platform.cuh, -DSDRD_EMU): test infrastructure."""
    from sdrdaemon_b200 import capi

    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")], check=True)
    return capi.load(os.path.join(ROOT, "tests", "emu", "libsdrd_emu.so"))


@pytest.fixture(scope="session")
def gpu_lib():
    """The product library; fails (does not skip) when it is missing or no sm_100 device is usable."""
    from sdrdaemon_b200 import capi

    lib = capi.load()
    assert lib.sdrd_device_count() >= 1, "no sm_100 device: the library has no CPU path"
    return lib
