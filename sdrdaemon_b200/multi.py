"""Multi-GPU sharding of the hot path: one process per GPU, independent IQ streams partitioned in
contiguous ranges (stream s -> rank s // (S / G)); there is no data-path collective -- the
only exchange is the gather of one 8-byte digest per stream (torch.distributed all_gather)."""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np


def stream_range(n_streams: int, world: int, rank: int) -> Tuple[int, int]:
    """(first stream, number of streams) owned by `rank`: contiguous, sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(n_streams, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def datagram_digest(dgrams: np.ndarray) -> np.ndarray:
    """One uint64 per stream over all of its datagram bytes (order-sensitive FNV-style fold of the
    64-bit words, cheap and good enough to compare shards with a single-process run)."""
    d = np.ascontiguousarray(dgrams).reshape(dgrams.shape[0], -1)
    w = d.view(np.uint64)
    k = (np.arange(w.shape[1], dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) | np.uint64(1)
    with np.errstate(over="ignore"):
        return (w * k).sum(axis=1, dtype=np.uint64)


def gather_digests(local: np.ndarray, n_streams: int, world: int, rank: int, device=None) -> Optional[np.ndarray]:
    """all_gather the per-stream digests of every rank's shard; returns the (n_streams,) array."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return local.copy()
    counts = [stream_range(n_streams, world, r)[1] for r in range(world)]
    width = max(counts)
    buf = torch.zeros(width, dtype=torch.int64, device=device)
    buf[: len(local)] = torch.from_numpy(local.view(np.int64)).to(buf.device)
    out = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    parts: List[np.ndarray] = [o.cpu().numpy().view(np.uint64)[:c] for o, c in zip(out, counts)]
    return np.concatenate(parts)


def rx_sharded(x_local: np.ndarray, log2_decim: int, n_fec: int, lib=None, **sink_kw) -> np.ndarray:
    """Run the Rx path (decimate + frame + encode) on this rank's streams; returns datagrams
    (S_local, n_frames, 128 + n_fec, 512)."""
    from . import capi

    s, n, _ = x_local.shape
    rx = capi.Rx(log2_decim, n_streams=s, max_in=n, n_fec=n_fec, lib=lib, **sink_kw)
    out = rx.process(x_local)
    rx.close()
    return out
