"""The CUDA library against the golden vectors produced by the reference's own code."""
import numpy as np
import pytest

import golden_cases
from sdrdaemon_b200 import capi

pytestmark = pytest.mark.gpu


def test_decimator_golden(gpu_lib):
    class D:
        def __init__(self, M, fc, v):
            self.d = capi.Decimator(M, fc, v, max_in=8192, lib=gpu_lib)

        def process(self, x, bits):
            return self.d.process(x, bits)
    golden_cases.check_decimator_golden(D)


def test_interpolator_golden(gpu_lib):
    golden_cases.check_interpolator_golden(lambda M: capi.Interpolator(M, max_in=2048, lib=gpu_lib))


def test_reconfigure_golden(gpu_lib):
    """mid-stream Downsampler::configure / Upsampler::configure sequences recorded from the reference build"""
    golden_cases.check_reconfigure_golden(lambda M, fc, v: capi.Decimator(M, fc, v, max_in=1 << 15, lib=gpu_lib),
                                          lambda M: capi.Interpolator(M, max_in=1024, lib=gpu_lib))


def test_sink_golden(gpu_lib):
    def factory(F, tv_sec, tv_usec):
        class S:
            def __init__(self):
                self.k = capi.Sink(max_samples=3 * golden_cases.FRAME, n_fec=F, tv_sec=tv_sec, tv_usec=tv_usec, lib=gpu_lib)

            def set_time(self, a, b):
                self.k.set_time(a, b)

            def write(self, x):
                return self.k.write(x)
        return S()
    golden_cases.check_sink_golden(factory)


def test_fecbuffer_golden(gpu_lib):
    golden_cases.check_fecbuffer_golden(lambda sb: capi.fec_decode(sb[None], [len(sb)], lib=gpu_lib)[0][0])
