"""ctypes binding of the C ABI in include/sdrd_b200.h (libsdrd_b200.so).

This is the Python face of the library used by the tests and by bench.py; the arithmetic lives in
the CUDA kernels behind the C ABI.  There is no fallback: if the shared library is missing or no
sm_100 device is present, loading / handle creation raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PRODUCT_SO = os.path.join(HERE, "libsdrd_b200.so")

FC_INFRA, FC_SUPRA, FC_CENTER = 0, 1, 2
HB_EO1, HB_DB = 0, 1
UDPSIZE, NB_ORIGINAL, BLOCK_BYTES, SAMPLES_PER_BLOCK = 512, 128, 508, 127
FRAME_SAMPLES = 127 * 127
FRAME_INCOMPLETE, FRAME_COMPLETE, FRAME_RECOVERED, FRAME_FAILED = 0, 1, 2, -1

# every symbol include/sdrd_b200.h declares: (name, restype, argtypes)
_P = C.c_void_p
_SZ = C.c_size_t
_SZP = C.POINTER(C.c_size_t)
_UP = C.POINTER(C.c_uint)
SYMBOLS = [
    ("sdrd_last_error", C.c_char_p, []),
    ("sdrd_version", C.c_char_p, []),
    ("sdrd_device_count", C.c_int, []),
    ("sdrd_set_device", C.c_int, [C.c_int]),
    ("sdrd_dec_create", C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.c_int, C.c_int, _SZ]),
    ("sdrd_dec_destroy", None, [_P]),
    ("sdrd_dec_reset", C.c_int, [_P]),
    ("sdrd_dec_configure", C.c_int, [_P, C.c_int, C.c_int]),
    ("sdrd_dec_log2_decim", C.c_int, [_P]),
    ("sdrd_dec_process", C.c_int, [_P, _P, _SZ, _SZ, _P, _SZ, _SZP, _UP]),
    ("sdrd_dec_rescale", C.c_int, [_P, _P, _SZ, _SZ, _UP]),
    ("sdrd_dec_dev_input", _P, [_P, _SZP]),
    ("sdrd_dec_dev_output", _P, [_P, _SZP]),
    ("sdrd_dec_process_dev", C.c_int, [_P, _SZ, _SZP, _UP, _P]),
    ("sdrd_dec_launches", C.c_longlong, [_P]),
    ("sdrd_dec_ipc_export", C.c_int, [_P, _P, _SZP, _SZP]),
    ("sdrd_ipc_open", C.c_int, [_P, C.POINTER(_P)]),
    ("sdrd_ipc_close", C.c_int, [_P]),
    ("sdrd_ipc_copy_rows", C.c_int, [_P, _SZ, _P, _SZ, _SZ, _SZ, _P]),
    ("sdrd_src_create", C.c_int, [C.POINTER(_P), _SZ]),
    ("sdrd_src_destroy", None, [_P]),
    ("sdrd_src_reset", C.c_int, [_P]),
    ("sdrd_src_feed", C.c_int, [_P, _P, _SZ, _P, _P, _SZ, _SZP, _P, _P, _P]),
    ("sdrd_src_cur_nb_blocks", C.c_int, [_P]),
    ("sdrd_src_cur_nb_recovery", C.c_int, [_P]),
    ("sdrd_src_min_nb_blocks", C.c_int, [_P]),
    ("sdrd_src_max_nb_recovery", C.c_int, [_P]),
    ("sdrd_src_launches", C.c_longlong, [_P]),
    ("sdrd_int_create", C.c_int, [C.POINTER(_P), C.c_int, C.c_int, _SZ]),
    ("sdrd_int_destroy", None, [_P]),
    ("sdrd_int_reset", C.c_int, [_P]),
    ("sdrd_int_configure", C.c_int, [_P, C.c_int]),
    ("sdrd_int_log2_interp", C.c_int, [_P]),
    ("sdrd_int_process", C.c_int, [_P, _P, _SZ, _SZ, _P, _SZ, _SZP]),
    ("sdrd_int_dev_input", _P, [_P, _SZP]),
    ("sdrd_int_dev_output", _P, [_P, _SZP]),
    ("sdrd_int_process_dev", C.c_int, [_P, _SZ, _SZP, _P]),
    ("sdrd_int_launches", C.c_longlong, [_P]),
    ("sdrd_cm256_encode", C.c_int, [_P, _SZ, C.c_int, C.c_int, _P]),
    ("sdrd_cm256_encode_dev", C.c_int, [_P, _SZ, C.c_int, C.c_int, _P, _P]),
    ("sdrd_sink_create", C.c_int, [C.POINTER(_P), C.c_int, _SZ]),
    ("sdrd_sink_destroy", None, [_P]),
    ("sdrd_sink_reset", C.c_int, [_P]),
    ("sdrd_sink_set_meta", C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_uint8, C.c_uint8]),
    ("sdrd_sink_set_nb_fec", C.c_int, [_P, C.c_int]),
    ("sdrd_sink_set_time", C.c_int, [_P, C.c_int, C.c_uint32, C.c_uint32]),
    ("sdrd_sink_blocks_per_frame", C.c_int, [_P]),
    ("sdrd_sink_frames_for", _SZ, [_P, _SZ]),
    ("sdrd_sink_write", C.c_int, [_P, _P, _SZ, _SZ, _P, _SZ, _SZP]),
    ("sdrd_sink_write_dev", C.c_int, [_P, _P, _SZ, _SZ, _SZP, _P]),
    ("sdrd_sink_dev_datagrams", _P, [_P, _SZP]),
    ("sdrd_sink_launches", C.c_longlong, [_P]),
    ("sdrd_rx_create", C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.c_int, C.c_int, _SZ]),
    ("sdrd_rx_destroy", None, [_P]),
    ("sdrd_rx_reset", C.c_int, [_P]),
    ("sdrd_rx_dec", _P, [_P]),
    ("sdrd_rx_sink", _P, [_P]),
    ("sdrd_rx_process", C.c_int, [_P, _P, _SZ, _SZ, _P, _SZ, _SZP, _UP]),
    ("sdrd_rx_set_slice_bytes", C.c_int, [_P, _SZ]),
    ("sdrd_rx_submit", C.c_int, [_P, _P, _SZ, _SZ, _UP]),
    ("sdrd_rx_collect", C.c_int, [_P, _P, _SZ, _SZP, C.POINTER(C.c_int), C.c_int]),
    ("sdrd_rx_chains", C.c_longlong, [_P]),
    ("sdrd_rx_set_min_chain", C.c_int, [_P, _SZ]),
    ("sdrd_rx_set_staging_threads", C.c_int, [_P, C.c_int]),
    ("sdrd_rx_dev_datagrams", _P, [_P, _SZP]),
    ("sdrd_rx_process_dev", C.c_int, [_P, _SZ, _SZP, _UP, _P]),
    ("sdrd_rx_launches", C.c_longlong, [_P]),
    ("sdrd_fec_decode", C.c_int, [_P, _SZ, _P, C.c_int, _P, _P, _P]),
    ("sdrd_fec_decode_dev", C.c_int, [_P, _SZ, _P, C.c_int, _P, _P, _P, _P]),
]


class Cm256Block(C.Structure):
    _fields_ = [("Block", C.c_void_p), ("Index", C.c_ubyte)]


class Cm256Params(C.Structure):
    _fields_ = [("OriginalCount", C.c_int), ("RecoveryCount", C.c_int), ("BlockBytes", C.c_int)]


SYMBOLS += [
    ("sdrd_cm256_encode_blocks", C.c_int, [Cm256Params, C.POINTER(Cm256Block), _P]),
    ("sdrd_cm256_decode_blocks", C.c_int, [Cm256Params, C.POINTER(Cm256Block)]),
]


class SdrdError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"sdrd error {code}: {msg}")
        self.code = code
        self.msg = msg


class Library:
    """A loaded libsdrd_b200.so with typed entry points."""

    def __init__(self, path: Optional[str] = None):
        path = path or PRODUCT_SO
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        self.path = path
        self.dll = C.CDLL(path)
        for name, res, args in SYMBOLS:
            fn = getattr(self.dll, name)  # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args

    def __getattr__(self, name):
        return getattr(self.dll, name)

    def check(self, rc: int) -> None:
        if rc != 0:
            raise SdrdError(rc, (self.dll.sdrd_last_error() or b"").decode())


_default: Optional[Library] = None


def load(path: Optional[str] = None) -> Library:
    global _default
    if path is not None:
        return Library(path)
    if _default is None:
        # SDRD_B200_LIB: alternative build of the same library (kernel tuning experiments)
        _default = Library(os.environ.get("SDRD_B200_LIB") or None)
    return _default


def _iq3(a: np.ndarray) -> np.ndarray:
    """(n, 2) or (S, n, 2) int16 -> contiguous (S, n, 2)."""
    a = np.asarray(a)
    if a.dtype != np.int16:
        raise TypeError("IQ arrays are int16")
    if a.ndim == 2:
        a = a[None]
    if a.ndim != 3 or a.shape[2] != 2:
        raise ValueError("IQ arrays are (n, 2) or (streams, n, 2)")
    return np.ascontiguousarray(a)


class Decimator:
    """Downsampler + Decimators for n_streams independent streams (sdrd_dec_*)."""

    def __init__(self, log2_decim: int, fcpos: int = FC_CENTER, variant: int = HB_EO1, n_streams: int = 1,
                 max_in: int = 1 << 20, lib: Optional[Library] = None):
        self.lib = lib or load()
        self._h = _P()
        self.n_streams = n_streams
        self.max_in = max_in
        self.lib.check(self.lib.sdrd_dec_create(C.byref(self._h), log2_decim, fcpos, variant, n_streams, max_in))

    def close(self):
        if getattr(self, "_h", None):
            self.lib.sdrd_dec_destroy(self._h)
            self._h = None

    __del__ = close

    @property
    def handle(self):
        return self._h

    def reset(self):
        self.lib.check(self.lib.sdrd_dec_reset(self._h))

    def configure(self, log2_decim: int, fcpos: int):
        self.lib.check(self.lib.sdrd_dec_configure(self._h, log2_decim, fcpos))

    @property
    def log2_decim(self) -> int:
        return self.lib.sdrd_dec_log2_decim(self._h)

    @property
    def launches(self) -> int:
        return self.lib.sdrd_dec_launches(self._h)

    def process(self, iq: np.ndarray, sample_bits: int = 16) -> Tuple[np.ndarray, int]:
        """Downsampler::process.  iq (n,2) or (S,n,2) int16 -> (out with the same leading shape, sample_bits)."""
        single = np.asarray(iq).ndim == 2
        a = _iq3(iq)
        s, n, _ = a.shape
        if s != self.n_streams:
            raise ValueError(f"expected {self.n_streams} streams, got {s}")
        out = np.zeros((s, max(n, 1), 2), dtype=np.int16)
        n_out = C.c_size_t(0)
        ss = C.c_uint(sample_bits)
        self.lib.check(self.lib.sdrd_dec_process(self._h, a.ctypes.data, n, n, out.ctypes.data, out.shape[1],
                                                 C.byref(n_out), C.byref(ss)))
        out = out[:, : n_out.value].copy()
        return (out[0] if single else out), ss.value

    # device-resident form -------------------------------------------------------------------
    def dev_input(self) -> Tuple[int, int]:
        st = C.c_size_t(0)
        p = self.lib.sdrd_dec_dev_input(self._h, C.byref(st))
        return p, st.value

    def dev_output(self) -> Tuple[int, int]:
        st = C.c_size_t(0)
        p = self.lib.sdrd_dec_dev_output(self._h, C.byref(st))
        return p, st.value

    def process_dev(self, n_in: int, sample_bits: int = 16, stream: int = 0) -> Tuple[int, int]:
        n_out = C.c_size_t(0)
        ss = C.c_uint(sample_bits)
        self.lib.check(self.lib.sdrd_dec_process_dev(self._h, n_in, C.byref(n_out), C.byref(ss), _P(stream)))
        return n_out.value, ss.value


def cm256_encode(originals: np.ndarray, n_fec: int, lib: Optional[Library] = None) -> np.ndarray:
    """originals (n_frames, 128, 508|512) uint8 -> (n_frames, n_fec, 508) recovery blocks (sdrd_cm256_encode).

    With a last dimension of 512 the rows are datagram images and the payload starts 4 bytes in."""
    lib = lib or load()
    o = np.ascontiguousarray(originals, dtype=np.uint8)
    if o.ndim == 2:
        o = o[None]
    nf, k, b = o.shape
    if k != NB_ORIGINAL or b not in (BLOCK_BYTES, UDPSIZE):
        raise ValueError("originals must be (n_frames, 128, 508) or (n_frames, 128, 512)")
    rec = np.zeros((nf, max(n_fec, 0), BLOCK_BYTES), dtype=np.uint8)
    base = o.ctypes.data + (4 if b == UDPSIZE else 0)
    if b == UDPSIZE:  # keep the last row's payload inside the buffer the library copies
        o = np.concatenate([o.reshape(-1), np.zeros(16, np.uint8)])
        base = o.ctypes.data + 4
    lib.check(lib.sdrd_cm256_encode(base, b, nf, n_fec, rec.ctypes.data))
    return rec


def cm256_encode_blocks(blocks, n_fec: int, block_bytes: int = BLOCK_BYTES, lib: Optional[Library] = None) -> np.ndarray:
    """cm256_encode through the descriptor API: blocks = list of 128 uint8 arrays (any memory layout)."""
    lib = lib or load()
    desc = (Cm256Block * 128)()
    keep = [np.ascontiguousarray(b, dtype=np.uint8) for b in blocks]
    for j, b in enumerate(keep):
        desc[j].Block = b.ctypes.data
        desc[j].Index = j
    out = np.zeros((n_fec, block_bytes), dtype=np.uint8)
    lib.check(lib.sdrd_cm256_encode_blocks(Cm256Params(128, n_fec, block_bytes), desc, out.ctypes.data))
    return out


def cm256_decode_blocks(blocks: np.ndarray, indices, recovery_count: int, lib: Optional[Library] = None):
    """cm256_decode through the descriptor API, in place on a copy: returns (rc, blocks, rewritten indices)."""
    lib = lib or load()
    blocks = np.ascontiguousarray(blocks, dtype=np.uint8).copy()
    k, b = blocks.shape
    desc = (Cm256Block * k)()
    for i in range(k):
        desc[i].Block = blocks[i].ctypes.data
        desc[i].Index = int(indices[i])
    rc = lib.sdrd_cm256_decode_blocks(Cm256Params(k, recovery_count, b), desc)
    return rc, blocks, [desc[i].Index for i in range(k)]


def fec_decode(superblocks: np.ndarray, n_blocks, lib: Optional[Library] = None):
    """superblocks (n_frames, pitch, 512) uint8 received datagrams in arrival order, n_blocks (n_frames,)
    -> (payload (n_frames,127,508), block0 (n_frames,508), status (n_frames,))  (sdrd_fec_decode)."""
    lib = lib or load()
    sb = np.ascontiguousarray(superblocks, dtype=np.uint8)
    if sb.ndim == 2:
        sb = sb[None]
    nf, pitch, w = sb.shape
    if w != UDPSIZE:
        raise ValueError("superblocks are 512-byte datagrams")
    nb = np.ascontiguousarray(np.broadcast_to(np.asarray(n_blocks, dtype=np.int32), (nf,)))
    payload = np.zeros((nf, 127, BLOCK_BYTES), dtype=np.uint8)
    block0 = np.zeros((nf, BLOCK_BYTES), dtype=np.uint8)
    status = np.zeros(nf, dtype=np.int32)
    lib.check(lib.sdrd_fec_decode(sb.ctypes.data, pitch, nb.ctypes.data, nf, payload.ctypes.data, block0.ctypes.data,
                                  status.ctypes.data))
    return payload, block0, status



class Interpolator:
    """Upsampler + Interpolators for n_streams independent streams (sdrd_int_*)."""

    def __init__(self, log2_interp: int, n_streams: int = 1, max_in: int = 1 << 16, lib: Optional[Library] = None):
        self.lib = lib or load()
        self._h = _P()
        self.n_streams = n_streams
        self.max_in = max_in
        self.lib.check(self.lib.sdrd_int_create(C.byref(self._h), log2_interp, n_streams, max_in))

    def close(self):
        if getattr(self, "_h", None):
            self.lib.sdrd_int_destroy(self._h)
            self._h = None

    __del__ = close

    def reset(self):
        self.lib.check(self.lib.sdrd_int_reset(self._h))

    def configure(self, log2_interp: int):
        self.lib.check(self.lib.sdrd_int_configure(self._h, log2_interp))

    @property
    def log2_interp(self) -> int:
        return self.lib.sdrd_int_log2_interp(self._h)

    @property
    def launches(self) -> int:
        return self.lib.sdrd_int_launches(self._h)

    def process(self, iq: np.ndarray) -> np.ndarray:
        """Upsampler::process.  iq (n,2) or (S,n,2) int16 -> (n << log2_interp) samples per stream."""
        single = np.asarray(iq).ndim == 2
        a = _iq3(iq)
        s, n, _ = a.shape
        if s != self.n_streams:
            raise ValueError(f"expected {self.n_streams} streams, got {s}")
        out = np.zeros((s, max(n << self.log2_interp, 1), 2), dtype=np.int16)
        n_out = C.c_size_t(0)
        self.lib.check(self.lib.sdrd_int_process(self._h, a.ctypes.data, n, n, out.ctypes.data, out.shape[1], C.byref(n_out)))
        out = out[:, : n_out.value].copy()
        return out[0] if single else out

    def dev_input(self) -> Tuple[int, int]:
        st = C.c_size_t(0)
        p = self.lib.sdrd_int_dev_input(self._h, C.byref(st))
        return p, st.value

    def dev_output(self) -> Tuple[int, int]:
        st = C.c_size_t(0)
        p = self.lib.sdrd_int_dev_output(self._h, C.byref(st))
        return p, st.value

    def process_dev(self, n_in: int, stream: int = 0) -> int:
        n_out = C.c_size_t(0)
        self.lib.check(self.lib.sdrd_int_process_dev(self._h, n_in, C.byref(n_out), _P(stream)))
        return n_out.value

class Sink:
    """UDPSinkFEC framing + encode for n_streams streams (sdrd_sink_*)."""

    def __init__(self, n_streams: int = 1, max_samples: int = 1 << 20, center_freq_khz: int = 435000,
                 sample_rate: int = 625000, n_fec: int = 16, tv_sec: Optional[int] = 1700000000, tv_usec: int = 0,
                 sample_bytes: int = 2, sample_bits: int = 16, lib: Optional[Library] = None, _handle=None):
        self.lib = lib or load()
        self.n_streams = n_streams
        self._own = _handle is None
        self._h = _P()
        if _handle is None:
            self.lib.check(self.lib.sdrd_sink_create(C.byref(self._h), n_streams, max_samples))
        else:
            self._h = _handle
        self.set_meta(center_freq_khz, sample_rate, sample_bytes, sample_bits)
        self.set_nb_fec(n_fec)
        if tv_sec is not None:
            self.set_time(tv_sec, tv_usec)

    def close(self):
        if getattr(self, "_h", None) and self._own:
            self.lib.sdrd_sink_destroy(self._h)
        self._h = None

    __del__ = close

    def reset(self):
        self.lib.check(self.lib.sdrd_sink_reset(self._h))

    def set_meta(self, center_freq_khz, sample_rate, sample_bytes=2, sample_bits=16):
        self.lib.check(self.lib.sdrd_sink_set_meta(self._h, center_freq_khz, sample_rate, sample_bytes, sample_bits))

    def set_nb_fec(self, n_fec: int):
        self.lib.check(self.lib.sdrd_sink_set_nb_fec(self._h, n_fec))

    def set_time(self, tv_sec: int, tv_usec: int = 0, fixed: bool = True, per_frame: bool = False):
        """fixed: the time of a call is (tv_sec, tv_usec) instead of the wall clock; per_frame: every frame begun in a
        call is stamped with the call's time + its sample offset / sample_rate (UDPSinkFEC.cpp:89-95)"""
        self.lib.check(self.lib.sdrd_sink_set_time(self._h, (1 if fixed else 0) | (2 if per_frame else 0), tv_sec, tv_usec))

    @property
    def blocks_per_frame(self) -> int:
        return self.lib.sdrd_sink_blocks_per_frame(self._h)

    def frames_for(self, n: int) -> int:
        return self.lib.sdrd_sink_frames_for(self._h, n)

    @property
    def handle(self):
        return self._h

    @property
    def launches(self) -> int:
        return self.lib.sdrd_sink_launches(self._h)

    def write_dev(self, dev_ptr: int, n: int, stride: int, stream: int = 0) -> int:
        nfr = C.c_size_t(0)
        self.lib.check(self.lib.sdrd_sink_write_dev(self._h, _P(dev_ptr), n, stride, C.byref(nfr), _P(stream)))
        return nfr.value

    def dev_datagrams(self) -> Tuple[int, int]:
        st = C.c_size_t(0)
        p = self.lib.sdrd_sink_dev_datagrams(self._h, C.byref(st))
        return p, st.value

    def write(self, iq: np.ndarray) -> np.ndarray:
        """UDPSinkFEC::write.  Returns the datagrams of the frames this call completed:
        (S, n_frames, 128 + n_fec, 512) uint8 (leading S dropped for a 2-D input)."""
        single = np.asarray(iq).ndim == 2
        a = _iq3(iq)
        s, n, _ = a.shape
        if s != self.n_streams:
            raise ValueError(f"expected {self.n_streams} streams, got {s}")
        cap = self.frames_for(n)
        bpf = self.blocks_per_frame
        out = np.zeros((s, max(cap, 1), bpf, UDPSIZE), dtype=np.uint8)
        nfr = C.c_size_t(0)
        self.lib.check(self.lib.sdrd_sink_write(self._h, a.ctypes.data, n, n, out.ctypes.data, out.shape[1], C.byref(nfr)))
        out = out[:, : nfr.value]
        return out[0] if single else out


class Rx:
    """Downsampler -> UDPSinkFEC without leaving the device (sdrd_rx_*)."""

    def __init__(self, log2_decim: int, fcpos: int = FC_CENTER, variant: int = HB_EO1, n_streams: int = 1,
                 max_in: int = 1 << 20, lib: Optional[Library] = None, **sink_kw):
        self.lib = lib or load()
        self._h = _P()
        self.n_streams = n_streams
        self.log2_decim = log2_decim
        self.lib.check(self.lib.sdrd_rx_create(C.byref(self._h), log2_decim, fcpos, variant, n_streams, max_in))
        self.sink = Sink(n_streams=n_streams, lib=self.lib, _handle=_P(self.lib.sdrd_rx_sink(self._h)), **sink_kw)
        self.dec_handle = _P(self.lib.sdrd_rx_dec(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self.lib.sdrd_rx_destroy(self._h)
            self._h = None

    __del__ = close

    def reset(self):
        self.lib.check(self.lib.sdrd_rx_reset(self._h))

    @property
    def launches(self) -> int:
        return self.lib.sdrd_rx_launches(self._h)

    def frames_for(self, n_in: int) -> int:
        return self.sink.frames_for(n_in >> self.log2_decim)

    def set_slice_bytes(self, min_call_bytes: int) -> None:
        """calls of at least this many input bytes go through in 8 overlapped slices (0: the default, 32 MiB)"""
        self.lib.check(self.lib.sdrd_rx_set_slice_bytes(self._h, min_call_bytes))

    def process(self, iq: np.ndarray, out: Optional[np.ndarray] = None, sample_bits: int = 16) -> np.ndarray:
        """sample_bits: the source's bits per component; the decimator's output size is left in
        self.sample_bits_out and written into the frames' meta data (sdrdaemonrx.cpp:618-643)."""
        single = np.asarray(iq).ndim == 2
        a = _iq3(iq)
        s, n, _ = a.shape
        cap = max(self.frames_for(n), 1)
        bpf = self.sink.blocks_per_frame
        if out is None:
            out = np.zeros((s, cap, bpf, UDPSIZE), dtype=np.uint8)
        nfr = C.c_size_t(0)
        ss = C.c_uint(sample_bits)
        self.lib.check(self.lib.sdrd_rx_process(self._h, a.ctypes.data, n, n, out.ctypes.data, out.shape[1], C.byref(nfr),
                                                C.byref(ss)))
        self.sample_bits_out = ss.value
        res = out[:, : nfr.value]
        return res[0] if single else res

    def submit(self, iq: np.ndarray, sample_bits: int = 16) -> int:
        """queued form: returns once the block is staged; the decimator's output sample size"""
        a = _iq3(iq)
        s, n, _ = a.shape
        ss = C.c_uint(sample_bits)
        self.lib.check(self.lib.sdrd_rx_submit(self._h, a.ctypes.data, n, n, C.byref(ss)))
        return ss.value

    def collect(self, frame_capacity: int = 64, wait: bool = False) -> np.ndarray:
        """frames completed so far -> (S, n_frames, blocks_per_frame, 512)"""
        out = np.zeros(self.n_streams * max(frame_capacity, 1) * 256 * UDPSIZE, dtype=np.uint8)
        nfr = C.c_size_t(0)
        bpf = C.c_int(0)
        self.lib.check(self.lib.sdrd_rx_collect(self._h, out.ctypes.data, frame_capacity, C.byref(nfr), C.byref(bpf), 1 if wait else 0))
        if nfr.value == 0:
            return np.zeros((self.n_streams, 0, self.sink.blocks_per_frame, UDPSIZE), np.uint8)
        used = self.n_streams * frame_capacity * bpf.value * UDPSIZE  # stream pitch = frame_capacity frames
        return out[:used].reshape(self.n_streams, frame_capacity, bpf.value, UDPSIZE)[:, : nfr.value].copy()

    @property
    def chains(self) -> int:
        return self.lib.sdrd_rx_chains(self._h)

    def set_min_chain(self, min_samples: int) -> None:
        self.lib.check(self.lib.sdrd_rx_set_min_chain(self._h, min_samples))

    def set_staging_threads(self, n_helpers: int) -> None:
        """Helper threads that share submit's copy into page-locked memory with the caller."""
        self.lib.check(self.lib.sdrd_rx_set_staging_threads(self._h, n_helpers))

    def dev_input(self) -> Tuple[int, int]:
        st = C.c_size_t(0)
        p = self.lib.sdrd_dec_dev_input(self.dec_handle, C.byref(st))
        return p, st.value

    def dev_datagrams(self) -> Tuple[int, int]:
        st = C.c_size_t(0)
        p = self.lib.sdrd_rx_dev_datagrams(self._h, C.byref(st))
        return p, st.value

    def process_dev(self, n_in: int, stream: int = 0, sample_bits: int = 16) -> int:
        nfr = C.c_size_t(0)
        ss = C.c_uint(sample_bits)
        self.lib.check(self.lib.sdrd_rx_process_dev(self._h, n_in, C.byref(nfr), C.byref(ss), _P(stream)))
        self.sample_bits_out = ss.value
        return nfr.value


class Source:
    """Batched receiver framing (sdrd_src_*): SDRdaemonFECBuffer::writeAndRead over bursts of datagrams."""

    def __init__(self, max_datagrams: int = 4096, lib: Optional[Library] = None):
        self.lib = lib or load()
        self._h = _P()
        self.max_datagrams = max_datagrams
        self.lib.check(self.lib.sdrd_src_create(C.byref(self._h), max_datagrams))

    def close(self):
        if getattr(self, "_h", None):
            self.lib.sdrd_src_destroy(self._h)
            self._h = None

    __del__ = close

    def reset(self):
        self.lib.check(self.lib.sdrd_src_reset(self._h))

    def feed(self, datagrams: np.ndarray):
        """datagrams (n, 512) uint8 in arrival order -> (payload (f,127,508), block0 (f,508), status, nb_blocks, nb_recovery)
        for the f frames closed by this burst."""
        dg = np.ascontiguousarray(datagrams, dtype=np.uint8).reshape(-1, 512)
        n = len(dg)
        cap = n + 1
        pay = np.zeros((cap, 127, 508), np.uint8)
        b0 = np.zeros((cap, 508), np.uint8)
        st = np.zeros(cap, np.int32)
        nbl = np.zeros(cap, np.int32)
        nrec = np.zeros(cap, np.int32)
        nf = C.c_size_t(0)
        self.lib.check(self.lib.sdrd_src_feed(self._h, dg.ctypes.data, n, pay.ctypes.data, b0.ctypes.data, cap, C.byref(nf),
                                              st.ctypes.data, nbl.ctypes.data, nrec.ctypes.data))
        f = nf.value
        return pay[:f].copy(), b0[:f].copy(), st[:f].copy(), nbl[:f].copy(), nrec[:f].copy()

    def stats(self) -> Tuple[int, int]:
        return self.lib.sdrd_src_cur_nb_blocks(self._h), self.lib.sdrd_src_cur_nb_recovery(self._h)

    def min_nb_blocks(self) -> int:
        return self.lib.sdrd_src_min_nb_blocks(self._h)

    def max_nb_recovery(self) -> int:
        return self.lib.sdrd_src_max_nb_recovery(self._h)

    @property
    def launches(self) -> int:
        return self.lib.sdrd_src_launches(self._h)
