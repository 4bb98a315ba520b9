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
    x = _inputs()
    first, count = multi.stream_range(S, world, rank)
    dg = multi.rx_sharded(x[first:first + count], M, F, lib=lib)
    dig = multi.gather_digests(multi.datagram_digest(dg), S, world, rank)
    if rank == 0:
        q.put(dig)
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
    want = multi.datagram_digest(multi.rx_sharded(x, M, F, lib=emu_lib))
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert np.array_equal(got, want)
