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


def test_segments_too_short_raise_on_every_rank_before_any_exchange():
    # global quantities only: every rank computes the same answer (the old code left
    # neighbours blocked in irecv while one rank raised)
    with pytest.raises(ValueError):
        mg.time_segments(3, 8, 4)                 # per == 0
    with pytest.raises(ValueError):
        mg.time_segments(1000, 4, 1, halo_len=4095)
    assert mg.time_segments(1000, 1, 1, halo_len=4095) == [(0, 1000)]   # one rank: no halo needed


class _OracleFir:
    """Stands in for newsched_b200.FirFilter on CPU tensors (the plumbing is what is tested here)."""

    def __init__(self, taps, D):
        self.taps, self.D, self.n_taps = taps, D, len(taps)

    def work_segment(self, x, halo=None, out=None):
        y = o.fir(x.numpy(), self.taps, self.D, hist=None if halo is None else halo.numpy(), precise=False)
        return torch.from_numpy(y)


def _worker_segfir(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(3)
        n, T = 30000, 129
        x = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)
        taps = (rng.uniform(-1, 1, T) / T).astype(np.float32)
        lo, hi = mg.time_segments(n, world, 1, halo_len=T - 1)[rank]
        sf = mg.SegmentedFir(_OracleFir(taps, 1), rank, world)
        y = sf.run(torch.from_numpy(x[lo:hi].copy()))
        full = mg.gather_concat(y, rank, world)
        ok = True
        if rank == 0:
            ok = bool(np.array_equal(full.numpy(), o.fir(x, taps, 1, precise=False)))
        # a segment shorter than the halo: BOTH ranks must raise (collective check), nobody hangs
        short = torch.zeros(T - 1 if rank == 0 else 10, dtype=torch.complex64)
        try:
            mg.SegmentedFir(_OracleFir(taps, 1), rank, world).run(short)
            ok = False
        except ValueError:
            pass
        q.put(ok)
    finally:
        dist.destroy_process_group()


def test_segmented_fir_and_collective_validation_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_segfir, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True and q.get(timeout=5) is True
