"""N > 1: the stream sharding and digest gather, exercised with world_size 2 on the gloo backend (CPU).
Each rank runs its shard through the emulated library; the gathered digests must equal those of a
single-process run over all streams."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from sdrdaemon_b200 import multi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
S, M, F = 5, 2, 4
N_IN = (127 * 127 + 50) << M


def _inputs():
    rng = np.random.default_rng(77)
    return rng.integers(-32768, 32768, size=(S, N_IN, 2), dtype=np.int16)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sdrdaemon_b200 import capi

    lib = capi.load(os.path.join(ROOT, "tests", "emu", "libsdrd_emu.so"))
    import torch

    # drop-in mode: only rank 0 holds the streams; scatter, process the shard, gather the datagrams back
    x_all = torch.from_numpy(_inputs()) if rank == 0 else torch.empty((0, N_IN, 2), dtype=torch.int16)
    x = multi.scatter_streams(x_all, S, world, rank).numpy()
    first, count = multi.stream_range(S, world, rank)
    assert x.shape == (count, N_IN, 2) and np.array_equal(x, _inputs()[first:first + count])
    dg = multi.rx_sharded(x, M, F, lib=lib)
    # the sliced scatter (pieces processed while later ones are on the wire) delivers the same streams, piece by piece
    pieces = multi.scatter_streams_sliced(x_all, S, world, rank, 2)
    got = []
    for works, t in pieces:
        for w in works:
            w.wait()
        got.append(t.numpy())
    assert np.array_equal(np.concatenate(got), x)
    dig = multi.gather_digests(multi.datagram_digest(dg), S, world, rank)
    all_dg = multi.gather_datagrams(torch.from_numpy(dg), S, world, rank)
    if rank == 0:
        q.put((dig, all_dg.numpy()))
    else:
        assert all_dg is None
    dist.barrier()
    dist.destroy_process_group()


def test_stream_range_partitions():
    for n in (1, 5, 256, 2048):
        for w in (1, 2, 3, 8):
            got = []
            for r in range(w):
                a, c = multi.stream_range(n, w, r)
                got += list(range(a, a + c))
            assert got == list(range(n))
    with pytest.raises(ValueError):
        multi.stream_range(4, 2, 2)


def test_two_ranks_match_single_process(emu_lib):
    x = _inputs()
    want_dg = multi.rx_sharded(x, M, F, lib=emu_lib)
    want = multi.datagram_digest(want_dg)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, got_dg = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert np.array_equal(got, want)
    assert np.array_equal(got_dg, want_dg)  # scatter -> shards -> gather == one process over all streams
