"""ctypes bindings for the CPU checker (oracle/liboracle.so) and, when built, the reference's own
code (oracle/_ref/libsdrd_ref_{eo1,db}.so).

TEST INFRASTRUCTURE ONLY.  Import this from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never from sdrdaemon_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Callable, List, Optional, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = {0: os.path.join(HERE, "_ref", "libsdrd_ref_eo1.so"), 1: os.path.join(HERE, "_ref", "libsdrd_ref_db.so")}

FC_INFRA, FC_SUPRA, FC_CENTER = 0, 1, 2
HB_EO1, HB_DB = 0, 1
UDPSIZE, NB_ORIGINAL, BLOCK_BYTES, SAMPLES_PER_BLOCK = 512, 128, 508, 127
FRAME_SAMPLES = 127 * 127


def build(ref: bool = True) -> None:
    """Compile the checker (and the reference build when /root/reference is present)."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if ref and os.path.isdir("/root/reference/sdmnbase"):
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)
        # the reference's sender / receiver sources over include/cm256.h and its two main programs over the host
        # layer (both link the product library: it has to be built first)
        if os.path.exists(os.path.join(HERE, "..", "sdrdaemon_b200", "libsdrd_b200.so")):
            subprocess.run(["make", "-s", "-C", HERE, "seam", "mains"], check=True)


class _Block(C.Structure):
    _fields_ = [("Block", C.c_void_p), ("Index", C.c_uint8)]


class _Params(C.Structure):
    _fields_ = [("OriginalCount", C.c_int), ("RecoveryCount", C.c_int), ("BlockBytes", C.c_int)]


_FRAME_CB = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_uint8), C.c_int, C.c_uint16)

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        L = C.CDLL(ORACLE_SO)
        L.sdro_dec_create.restype = C.c_void_p
        L.sdro_dec_create.argtypes = [C.c_int, C.c_int, C.c_int]
        L.sdro_dec_destroy.argtypes = [C.c_void_p]
        L.sdro_dec_reset.argtypes = [C.c_void_p]
        L.sdro_dec_configure.restype = C.c_int
        L.sdro_dec_configure.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.sdro_int_configure.restype = C.c_int
        L.sdro_int_configure.argtypes = [C.c_void_p, C.c_int]
        L.sdro_int_create.restype = C.c_void_p
        L.sdro_int_create.argtypes = [C.c_int]
        L.sdro_int_destroy.argtypes = [C.c_void_p]
        L.sdro_int_reset.argtypes = [C.c_void_p]
        L.sdro_int_process.restype = C.c_size_t
        L.sdro_int_process.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.sdro_dec_process.restype = C.c_size_t
        L.sdro_dec_process.argtypes = [C.c_void_p, C.POINTER(C.c_uint), C.c_void_p, C.c_size_t, C.c_void_p]
        for f in ("sdro_gf_mul", "sdro_gf_div"):
            getattr(L, f).restype = C.c_uint8
            getattr(L, f).argtypes = [C.c_uint8, C.c_uint8]
        L.sdro_gf_exp.restype = C.c_uint8
        L.sdro_gf_exp.argtypes = [C.c_int]
        L.sdro_gf_log.restype = C.c_uint8
        L.sdro_gf_log.argtypes = [C.c_uint8]
        L.sdro_cm256_matrix_element.restype = C.c_uint8
        L.sdro_cm256_matrix_element.argtypes = [C.c_uint8, C.c_uint8, C.c_uint8]
        L.sdro_cm256_encode.restype = C.c_int
        L.sdro_cm256_encode.argtypes = [_Params, C.POINTER(_Block), C.c_void_p]
        L.sdro_cm256_decode.restype = C.c_int
        L.sdro_cm256_decode.argtypes = [_Params, C.POINTER(_Block)]
        L.sdro_set_simd.argtypes = [C.c_int]
        L.sdro_simd.restype = C.c_int
        L.sdro_crc32.restype = C.c_uint32
        L.sdro_crc32.argtypes = [C.c_void_p, C.c_size_t]
        L.sdro_sink_create.restype = C.c_void_p
        L.sdro_sink_create.argtypes = [_FRAME_CB, C.c_void_p]
        L.sdro_sink_destroy.argtypes = [C.c_void_p]
        L.sdro_sink_set_meta.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint8, C.c_uint8]
        L.sdro_sink_set_nb_fec.argtypes = [C.c_void_p, C.c_int]
        L.sdro_sink_set_time.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.sdro_sink_write.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.sdro_fecbuf_create.restype = C.c_void_p
        L.sdro_fecbuf_destroy.argtypes = [C.c_void_p]
        L.sdro_fecbuf_write_and_read.restype = C.c_int
        L.sdro_fecbuf_write_and_read.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t)]
        for f in ("sdro_fecbuf_cur_nb_blocks", "sdro_fecbuf_cur_nb_recovery", "sdro_fecbuf_min_nb_blocks",
                  "sdro_fecbuf_max_nb_recovery"):
            getattr(L, f).restype = C.c_int
            getattr(L, f).argtypes = [C.c_void_p]
        L.sdro_fecbuf_current_meta.argtypes = [C.c_void_p, C.c_void_p]
        L.sdro_decode_frame.restype = C.c_int
        L.sdro_decode_frame.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _iq(a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.int16)
    assert a.ndim == 2 and a.shape[1] == 2, "IQ arrays are (n, 2) int16"
    return a


class Decimator:
    """sdro_dec_*: one Downsampler + Decimators state."""

    def __init__(self, log2_decim: int, fcpos: int = FC_CENTER, variant: int = HB_EO1):
        self._h = lib().sdro_dec_create(log2_decim, fcpos, variant)
        if not self._h:
            raise ValueError("Invalid log2 decimation factor / Fc position index")
        self.log2_decim = log2_decim

    def __del__(self):
        if getattr(self, "_h", None):
            lib().sdro_dec_destroy(self._h)
            self._h = None

    def reset(self) -> None:
        lib().sdro_dec_reset(self._h)

    def configure(self, log2_decim: int, fcpos: int = FC_CENTER) -> None:
        """Downsampler::configure between blocks: the six stage objects keep their state."""
        if lib().sdro_dec_configure(self._h, log2_decim, fcpos):
            raise ValueError("Invalid log2 decimation factor / Fc position index")
        self.log2_decim = log2_decim

    def process(self, iq: np.ndarray, sample_bits: int = 16) -> Tuple[np.ndarray, int]:
        iq = _iq(iq)
        out = np.zeros((max(len(iq) >> self.log2_decim, 1) + 1, 2), dtype=np.int16)
        ss = C.c_uint(sample_bits)
        n = lib().sdro_dec_process(self._h, C.byref(ss), iq.ctypes.data, len(iq), out.ctypes.data)
        return out[:n].copy(), ss.value


class Interpolator:
    """sdro_int_*: one Upsampler + Interpolators state."""

    def __init__(self, log2_interp: int):
        self._h = lib().sdro_int_create(log2_interp)
        if not self._h:
            raise ValueError("Invalid log2 interpolation factor")
        self.log2_interp = log2_interp

    def __del__(self):
        if getattr(self, "_h", None):
            lib().sdro_int_destroy(self._h)
            self._h = None

    def reset(self) -> None:
        lib().sdro_int_reset(self._h)

    def configure(self, log2_interp: int) -> None:
        """Upsampler::configure between blocks: the stage objects keep their state."""
        if lib().sdro_int_configure(self._h, log2_interp):
            raise ValueError("Invalid log2 interpolation factor")
        self.log2_interp = log2_interp

    def process(self, iq: np.ndarray) -> np.ndarray:
        iq = _iq(iq)
        out = np.zeros(((len(iq) << self.log2_interp) + 1, 2), dtype=np.int16)
        n = lib().sdro_int_process(self._h, iq.ctypes.data, len(iq), out.ctypes.data)
        return out[:n].copy()


def cm256_encode(originals: np.ndarray, n_fec: int) -> np.ndarray:
    """originals: (K, B) uint8 -> (n_fec, B) recovery blocks."""
    originals = np.ascontiguousarray(originals, dtype=np.uint8)
    k, b = originals.shape
    blocks = (_Block * k)()
    for i in range(k):
        blocks[i].Block = originals[i].ctypes.data
        blocks[i].Index = i
    out = np.zeros((n_fec, b), dtype=np.uint8)
    rc = lib().sdro_cm256_encode(_Params(k, n_fec, b), blocks, out.ctypes.data)
    if rc:
        raise RuntimeError(f"cm256_encode failed: {rc}")
    return out


def cm256_decode(blocks: np.ndarray, indices, original_count: int, recovery_count: int) -> Tuple[int, np.ndarray, List[int]]:
    """In-place decode of `original_count` received blocks; returns (rc, blocks, rewritten indices)."""
    blocks = np.ascontiguousarray(blocks, dtype=np.uint8).copy()
    k, b = blocks.shape
    desc = (_Block * k)()
    for i in range(k):
        desc[i].Block = blocks[i].ctypes.data
        desc[i].Index = int(indices[i])
    rc = lib().sdro_cm256_decode(_Params(original_count, recovery_count, b), desc)
    return rc, blocks, [desc[i].Index for i in range(k)]


def crc32(data: bytes) -> int:
    return lib().sdro_crc32(data, len(data))


class Sink:
    """sdro_sink_*: UDPSinkFEC::write + the encode half of transmitUDP; collects datagram images."""

    def __init__(self, center_freq_khz=435000, sample_rate=625000, n_fec=16, tv_sec=1700000000, tv_usec=0,
                 sample_bytes=2, sample_bits=16):
        self.frames: List[np.ndarray] = []

        def _cb(_user, data, n_blocks, _frame_index):
            self.frames.append(np.ctypeslib.as_array(data, shape=(n_blocks, UDPSIZE)).copy())

        self._cb = _FRAME_CB(_cb)
        self._h = lib().sdro_sink_create(self._cb, None)
        lib().sdro_sink_set_meta(self._h, center_freq_khz, sample_rate, sample_bytes, sample_bits)
        lib().sdro_sink_set_nb_fec(self._h, n_fec)
        lib().sdro_sink_set_time(self._h, tv_sec, tv_usec)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().sdro_sink_destroy(self._h)
            self._h = None

    def set_time(self, tv_sec: int, tv_usec: int) -> None:
        lib().sdro_sink_set_time(self._h, tv_sec, tv_usec)

    def set_nb_fec(self, n_fec: int) -> None:
        lib().sdro_sink_set_nb_fec(self._h, n_fec)

    def write(self, iq: np.ndarray) -> None:
        iq = _iq(iq)
        lib().sdro_sink_write(self._h, iq.ctypes.data, len(iq))


class FecBuffer:
    """sdro_fecbuf_*: SDRdaemonFECBuffer::writeAndRead."""

    def __init__(self):
        self._h = lib().sdro_fecbuf_create()

    def __del__(self):
        if getattr(self, "_h", None):
            lib().sdro_fecbuf_destroy(self._h)
            self._h = None

    def write_and_read(self, superblock: np.ndarray) -> Optional[np.ndarray]:
        sb = np.ascontiguousarray(superblock, dtype=np.uint8)
        assert sb.size == UDPSIZE
        out = np.zeros(127 * BLOCK_BYTES, dtype=np.uint8)
        n = C.c_size_t(0)
        if lib().sdro_fecbuf_write_and_read(self._h, sb.ctypes.data, out.ctypes.data, C.byref(n)):
            return out[: n.value]
        return None

    def stats(self) -> Tuple[int, int]:
        return lib().sdro_fecbuf_cur_nb_blocks(self._h), lib().sdro_fecbuf_cur_nb_recovery(self._h)

    def min_nb_blocks(self) -> int:
        return lib().sdro_fecbuf_min_nb_blocks(self._h)

    def max_nb_recovery(self) -> int:
        return lib().sdro_fecbuf_max_nb_recovery(self._h)


def decode_frame(superblocks: np.ndarray) -> Tuple[int, np.ndarray, np.ndarray]:
    """First <=128 received superblocks of one frame -> (status, payload 127x508, block0 508)."""
    sb = np.ascontiguousarray(superblocks, dtype=np.uint8).reshape(-1, UDPSIZE)
    payload = np.zeros((127, BLOCK_BYTES), dtype=np.uint8)
    block0 = np.zeros(BLOCK_BYTES, dtype=np.uint8)
    st = lib().sdro_decode_frame(sb.ctypes.data, len(sb), payload.ctypes.data, block0.ctypes.data)
    return st, payload, block0


# ----------------------------------------------------------------------------------------------
# The reference's own code (oracle/_ref).  Present in the build container and, prebuilt, on the
# GPU box; absent otherwise.
# ----------------------------------------------------------------------------------------------

_ref = {}


def ref_available(variant: int = HB_EO1) -> bool:
    return os.path.exists(REF_SO[variant])


def ref(variant: int = HB_EO1) -> C.CDLL:
    if variant not in _ref:
        L = C.CDLL(REF_SO[variant])
        L.ref_variant.restype = C.c_int
        L.ref_ds_create.restype = C.c_void_p
        L.ref_ds_create.argtypes = [C.c_int, C.c_int]
        L.ref_ds_destroy.argtypes = [C.c_void_p]
        L.ref_ds_process.restype = C.c_size_t
        L.ref_ds_process.argtypes = [C.c_void_p, C.POINTER(C.c_uint), C.c_void_p, C.c_size_t, C.c_void_p]
        L.ref_ds_configure.restype = C.c_int
        L.ref_ds_configure.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ref_us_configure.restype = C.c_int
        L.ref_us_configure.argtypes = [C.c_void_p, C.c_int]
        L.ref_ds_process_streams.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p,
                                             C.c_size_t]
        L.ref_us_create.restype = C.c_void_p
        L.ref_us_create.argtypes = [C.c_int]
        L.ref_us_destroy.argtypes = [C.c_void_p]
        L.ref_us_process.restype = C.c_size_t
        L.ref_us_process.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.ref_fecbuf_create.restype = C.c_void_p
        L.ref_fecbuf_destroy.argtypes = [C.c_void_p]
        L.ref_fecbuf_write_and_read.restype = C.c_int
        L.ref_fecbuf_write_and_read.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t)]
        for f in ("ref_fecbuf_cur_nb_blocks", "ref_fecbuf_cur_nb_recovery", "ref_fecbuf_min_nb_blocks",
                  "ref_fecbuf_max_nb_recovery"):
            getattr(L, f).restype = C.c_int
            getattr(L, f).argtypes = [C.c_void_p]
        L.ref_sink_run.restype = C.c_int
        L.ref_sink_run.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t,
                                   C.c_void_p, C.c_int, C.c_int]
        L.ref_testsource_read.restype = C.c_int
        L.ref_testsource_read.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.c_int, C.c_float, C.c_float]
        assert L.ref_variant() == variant
        _ref[variant] = L
    return _ref[variant]


class RefDownsampler:
    """The reference's Downsampler (+Decimators), EO1 or DB build."""

    def __init__(self, log2_decim: int, fcpos: int = FC_CENTER, variant: int = HB_EO1):
        self._L = ref(variant)
        self._h = self._L.ref_ds_create(log2_decim, fcpos)
        self.log2_decim = log2_decim

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.ref_ds_destroy(self._h)
            self._h = None

    def configure(self, log2_decim: int, fcpos: int = FC_CENTER) -> None:
        if not self._L.ref_ds_configure(self._h, log2_decim, fcpos):
            raise ValueError("the reference refused decim/fcpos")
        self.log2_decim = log2_decim

    def process(self, iq: np.ndarray, sample_bits: int = 16) -> Tuple[np.ndarray, int]:
        iq = _iq(iq)
        out = np.zeros((max(len(iq) >> self.log2_decim, 1) + 1, 2), dtype=np.int16)
        ss = C.c_uint(sample_bits)
        n = self._L.ref_ds_process(self._h, C.byref(ss), iq.ctypes.data, len(iq), out.ctypes.data)
        return out[:n].copy(), ss.value


class RefUpsampler:
    """The reference's Upsampler (+Interpolators), EO1 or DB build."""

    def __init__(self, log2_interp: int, variant: int = HB_EO1):
        self._L = ref(variant)
        self._h = self._L.ref_us_create(log2_interp)
        self.log2_interp = log2_interp

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.ref_us_destroy(self._h)
            self._h = None

    def configure(self, log2_interp: int) -> None:
        if not self._L.ref_us_configure(self._h, log2_interp):
            raise ValueError("the reference refused interp")
        self.log2_interp = log2_interp

    def process(self, iq: np.ndarray) -> np.ndarray:
        iq = _iq(iq)
        out = np.zeros(((len(iq) << self.log2_interp) + 1, 2), dtype=np.int16)
        n = self._L.ref_us_process(self._h, iq.ctypes.data, len(iq), out.ctypes.data)
        return out[:n].copy()


class RefFecBuffer:
    def __init__(self):
        self._L = ref(HB_EO1)
        self._h = self._L.ref_fecbuf_create()

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.ref_fecbuf_destroy(self._h)
            self._h = None

    def write_and_read(self, superblock: np.ndarray) -> Optional[np.ndarray]:
        sb = np.ascontiguousarray(superblock, dtype=np.uint8)
        out = np.zeros(128 * 512, dtype=np.uint8)
        n = C.c_size_t(0)
        if self._L.ref_fecbuf_write_and_read(self._h, sb.ctypes.data, out.ctypes.data, C.byref(n)):
            return out[: n.value]
        return None

    def stats(self) -> Tuple[int, int]:
        return self._L.ref_fecbuf_cur_nb_blocks(self._h), self._L.ref_fecbuf_cur_nb_recovery(self._h)

    def min_nb_blocks(self) -> int:
        return self._L.ref_fecbuf_min_nb_blocks(self._h)

    def max_nb_recovery(self) -> int:
        return self._L.ref_fecbuf_max_nb_recovery(self._h)


def ref_sink_run(iq: np.ndarray, n_fec: int, n_frames_out: int, port: int = 19090, chunk: int = 4096,
                 center_freq_khz: int = 435000, sample_rate: int = 625000) -> np.ndarray:
    """Push iq through the reference UDPSinkFEC over loop-back; returns (n, 512) captured datagrams."""
    iq = _iq(iq)
    per_frame = NB_ORIGINAL + n_fec
    expect = per_frame * n_frames_out
    out = np.zeros((expect + 2 * per_frame, UDPSIZE), dtype=np.uint8)
    got = ref(HB_EO1).ref_sink_run(port, center_freq_khz, sample_rate, n_fec, iq.ctypes.data, len(iq), chunk,
                                    out.ctypes.data, len(out), expect)
    if got < 0:
        raise RuntimeError(f"ref_sink_run failed: {got}")
    return out[:got]


def ref_testsource(n_samples: int, sample_rate: int, delta_phase: float, amplitude: float, phase: float = 0.0):
    buf = np.zeros((n_samples, 2), dtype=np.int16)
    ph = C.c_float(phase)
    n = ref(HB_EO1).ref_testsource_read(buf.ctypes.data, n_samples, C.byref(ph), sample_rate, delta_phase, amplitude)
    assert n == n_samples
    return buf, ph.value


# ----------------------------------------------------------------------------------------------
# bench.py CPU legs (the only place outside tests/ and smoke() that may execute the oracle)
# ----------------------------------------------------------------------------------------------

def cpu_rx_streams(iq: np.ndarray, log2_decim: int, n_fec: int, n_threads: int, fcpos: int = FC_CENTER,
                   block: int = 65536, prefer_reference: bool = True) -> Tuple[int, int, str]:
    """Decimate -> pack -> encode every stream of iq (S, n, 2) on n_threads host threads.

    Returns (superframes, digest, kind): kind "reference" when the reference's own Downsampler code
    (oracle/_ref, EO1 build) did the decimation, "port" when the C restatement did."""
    a = np.ascontiguousarray(iq, dtype=np.int16)
    if a.ndim == 2:
        a = a[None]
    s, n, _ = a.shape
    dig = C.c_uint32(0)
    if prefer_reference and ref_available(HB_EO1):
        L = ref(HB_EO1)
        L.ref_rx_streams.restype = C.c_longlong
        L.ref_rx_streams.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t,
                                     C.c_size_t, C.POINTER(C.c_uint32)]
        fr = L.ref_rx_streams(log2_decim, fcpos, n_fec, s, n_threads, a.ctypes.data, n, n, block, C.byref(dig))
        return int(fr), int(dig.value), "reference"
    L = lib()
    L.sdro_rx_streams.restype = C.c_longlong
    L.sdro_rx_streams.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                  C.c_size_t, C.c_size_t, C.POINTER(C.c_uint32)]
    fr = L.sdro_rx_streams(log2_decim, fcpos, HB_EO1, n_fec, s, n_threads, a.ctypes.data, n, n, block, C.byref(dig))
    return int(fr), int(dig.value), "port"


def rx_stream_crcs(iq: np.ndarray, log2_decim: int, n_fec: int, n_threads: int, fcpos: int = FC_CENTER,
                   variant: int = HB_EO1, block: int = 65536) -> Tuple[int, np.ndarray]:
    """Decimate -> pack -> encode every stream of iq (S, n, 2) with the oracle; returns (superframes, crc[S]) with
    crc[s] = zlib.crc32 of stream s's datagram bytes in send order."""
    a = np.ascontiguousarray(iq, dtype=np.int16)
    s, n, _ = a.shape
    L = lib()
    L.sdro_rx_streams_crc.restype = C.c_longlong
    L.sdro_rx_streams_crc.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                      C.c_size_t, C.c_size_t, C.POINTER(C.c_uint32), C.c_void_p]
    crc = np.zeros(s, dtype=np.uint32)
    dig = C.c_uint32(0)
    fr = L.sdro_rx_streams_crc(log2_decim, fcpos, variant, n_fec, s, n_threads, a.ctypes.data, n, n, block, C.byref(dig),
                               crc.ctypes.data)
    return int(fr), crc


def decode_frames(superblocks: np.ndarray, n_blocks, n_threads: int = 1):
    """(n_frames, pitch, 512) received datagrams -> (payload (n, 127, 508), block0 (n, 508), status (n,)), threaded."""
    sb = np.ascontiguousarray(superblocks, dtype=np.uint8)
    nf, pitch, _ = sb.shape
    nb = np.ascontiguousarray(np.broadcast_to(np.asarray(n_blocks, dtype=np.int32), (nf,)))
    pay = np.zeros((nf, 127, BLOCK_BYTES), np.uint8)
    b0 = np.zeros((nf, BLOCK_BYTES), np.uint8)
    st = np.zeros(nf, np.int32)
    L = lib()
    L.sdro_decode_frames.restype = None
    L.sdro_decode_frames.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.sdro_decode_frames(sb.ctypes.data, pitch, nb.ctypes.data, nf, n_threads, pay.ctypes.data, b0.ctypes.data, st.ctypes.data)
    return pay, b0, st


def set_simd(on: bool) -> None:
    """force the scalar (False) or allow the SSSE3 (True) block multiply of the restated CM256"""
    lib().sdro_set_simd(1 if on else 0)


def simd() -> bool:
    return bool(lib().sdro_simd())


def ref_decode_frames(superblocks: np.ndarray, n_blocks, n_threads: int = 1) -> np.ndarray:
    """(n_frames, pitch, 512) received datagrams through the REFERENCE's SDRdaemonFECBuffer (oracle/_ref, restated
    CM256 inside), n_threads buffers in parallel -> payload (n_frames, 127, 508)."""
    sb = np.ascontiguousarray(superblocks, dtype=np.uint8)
    nf, pitch, _ = sb.shape
    nb = np.ascontiguousarray(np.broadcast_to(np.asarray(n_blocks, dtype=np.int32), (nf,)))
    pay = np.zeros((nf, 127, BLOCK_BYTES), np.uint8)
    L = ref(HB_EO1)
    L.ref_fecbuf_decode_frames.restype = None
    L.ref_fecbuf_decode_frames.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.ref_fecbuf_decode_frames(sb.ctypes.data, pitch, nb.ctypes.data, nf, n_threads, pay.ctypes.data)
    return pay
