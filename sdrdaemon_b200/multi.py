"""Multi-GPU sharding of the hot path: one process per GPU, independent IQ streams partitioned in
contiguous ranges (stream s -> rank s // (S / G)); there is no data-path collective.  The
exchanges are the trivial ones either side of the path: the scatter of streams that only rank 0 holds
(drop-in mode: one host process feeds all GPUs), the gather of the datagram images back to it, and the
gather of one 8-byte digest per stream -- torch.distributed point-to-point / all_gather, NCCL over NVLink
on GPUs (one ncclGroup per scatter), gloo in the CPU tests."""
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


def _wire(t):
    """the same memory as bytes: NCCL has no 16-bit integer type, and the exchange only moves data"""
    import torch

    return t.contiguous().view(torch.uint8) if t.dtype not in (torch.uint8, torch.int8) else t.contiguous()


def scatter_streams(x_all, n_streams: int, world: int, rank: int, src: int = 0):
    """Rank `src` holds every stream, x_all (n_streams, ...) torch tensor (on the GPU for NCCL, on the CPU
    for gloo); every rank returns its contiguous range of streams (a tensor shaped (count, ...)).  Other
    ranks pass a tensor of the per-stream shape/dtype/device to receive into (x_all[:0] is enough)."""
    import torch
    import torch.distributed as dist

    first, count = stream_range(n_streams, world, rank)
    if world == 1:
        return x_all[first:first + count]
    if rank == src:
        ops = []
        for r in range(world):
            a, c = stream_range(n_streams, world, r)
            if r != src and c:
                ops.append(dist.P2POp(dist.isend, _wire(x_all[a:a + c]), r))
        for w in (dist.batch_isend_irecv(ops) if ops else []):
            w.wait()
        return x_all[first:first + count]
    mine = torch.empty((count,) + tuple(x_all.shape[1:]), dtype=x_all.dtype, device=x_all.device)
    if count:
        for w in dist.batch_isend_irecv([dist.P2POp(dist.irecv, _wire(mine), src)]):
            w.wait()
    return mine


def slice_range(count: int, n_slices: int, k: int) -> Tuple[int, int]:
    """(first, number) of slice k when `count` streams are cut into n_slices contiguous pieces"""
    base, extra = divmod(count, n_slices)
    first = k * base + min(k, extra)
    return first, base + (1 if k < extra else 0)


def scatter_streams_sliced(x_all, n_streams: int, world: int, rank: int, n_slices: int, src: int = 0):
    """The scatter of scatter_streams cut into n_slices pieces per rank so that a receiver can work on piece k while
    piece k + 1 is still on the wire: rank `src` sends piece 0 of every rank's range, then piece 1, ... (one group
    of point-to-point operations per piece, executed in order).  Returns a list of n_slices (work, tensor) pairs:
    `work.wait()` (None for data that is already local) makes the current stream wait for that piece; the tensor
    holds this rank's streams of the piece."""
    import torch
    import torch.distributed as dist

    first, count = stream_range(n_streams, world, rank)
    out = []
    if rank == src:
        for k in range(n_slices):
            ops = []
            for r in range(world):
                a, c = stream_range(n_streams, world, r)
                f, n = slice_range(c, n_slices, k)
                if r != src and n:
                    ops.append(dist.P2POp(dist.isend, _wire(x_all[a + f:a + f + n]), r))
            works = dist.batch_isend_irecv(ops) if ops else []
            f, n = slice_range(count, n_slices, k)
            out.append((works, x_all[first + f:first + f + n]))
        return out
    for k in range(n_slices):
        f, n = slice_range(count, n_slices, k)
        piece = torch.empty((n,) + tuple(x_all.shape[1:]), dtype=x_all.dtype, device=x_all.device)
        works = dist.batch_isend_irecv([dist.P2POp(dist.irecv, _wire(piece), src)]) if n else []
        out.append((works, piece))
    return out


def gather_datagrams(dg_local, n_streams: int, world: int, rank: int, dst: int = 0):
    """The reverse exchange: every rank's datagram images (count, frames, blocks, 512) to rank `dst`, which
    returns the (n_streams, frames, blocks, 512) tensor in stream order (the other ranks return None)."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return dg_local
    if rank != dst:
        if dg_local.shape[0]:
            for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, _wire(dg_local), dst)]):
                w.wait()
        return None
    out = torch.empty((n_streams,) + tuple(dg_local.shape[1:]), dtype=dg_local.dtype, device=dg_local.device)
    ops = []
    for r in range(world):
        a, c = stream_range(n_streams, world, r)
        if r == dst:
            out[a:a + c] = dg_local
        elif c:
            ops.append(dist.P2POp(dist.irecv, _wire(out[a:a + c]), r))
    for w in (dist.batch_isend_irecv(ops) if ops else []):
        w.wait()
    return out


def rx_sharded(x_local: np.ndarray, log2_decim: int, n_fec: int, lib=None, **sink_kw) -> np.ndarray:
    """Run the Rx path (decimate + frame + encode) on this rank's streams; returns datagrams
    (S_local, n_frames, 128 + n_fec, 512)."""
    from . import capi

    s, n, _ = x_local.shape
    rx = capi.Rx(log2_decim, n_streams=s, max_in=n, n_fec=n_fec, lib=lib, **sink_kw)
    out = rx.process(x_local)
    rx.close()
    return out
