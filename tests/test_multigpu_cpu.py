"""CPU, world_size 2 over gloo: the host-side sharding logic of the multi-GPU path (segment
bounds, halo hand-off to the right neighbour, final gather).  The per-rank FIR stands in with the
CPU oracle -- this file tests the plumbing, the kernels are tested in test_gpu_parity.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as o
from newsched_b200 import multigpu as mg


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, T, D, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)
        x = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)
        taps = (rng.uniform(-1, 1, T) / T).astype(np.float32)
        lo, hi = mg.time_segments(n, world, D)[rank]
        seg = torch.from_numpy(x[lo:hi].copy())
        halo = mg.exchange_halo(seg, T - 1, rank, world)
        if rank == 0:
            assert halo is None
        else:
            assert np.array_equal(halo.numpy(), x[lo - (T - 1):lo])
        y = o.fir(seg.numpy(), taps, D, hist=None if halo is None else halo.numpy(), precise=False)
        full = mg.gather_concat(torch.from_numpy(y), rank, world)
        if rank == 0:
            ref = o.fir(x, taps, D, precise=False)
            q.put(bool(np.array_equal(full.numpy(), ref)))
        # channel sharding bookkeeping
        b, c = mg.channel_slice(64, rank, world)
        assert (b, c) == (rank * 32, 32)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("T,D,n", [(64, 1, 40000), (257, 4, 100003)])
def test_time_segment_sharding_two_ranks(T, D, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, T, D, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_segment_bounds():
    segs = mg.time_segments(1000, 3, 4)
    assert segs == [(0, 332), (332, 664), (664, 1000)]
    assert all(lo % 4 == 0 for lo, _ in segs)
    assert mg.time_segments(4096 * 10, 1, 4096) == [(0, 40960)]
    with pytest.raises(ValueError):
        mg.channel_slice(64, 0, 3)
