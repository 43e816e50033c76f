"""Multi-GPU partitioning of the hot path: one process per GPU over
``torch.distributed``; the blocks shard with NO data-path collective.

* by stream / channel: independent per-GPU streams, or a slice of the channelizer's output
  channels per GPU (``channel_slice``);
* by time segment (BASELINE config 5): contiguous segments aligned to the decimation, each rank
  needs the ``ntaps-1`` samples before its segment -- a one-off halo from the left neighbour
  (``exchange_halo``: point-to-point send/recv, NCCL over NVLink on GPUs, gloo in the CPU tests);
* the only collective is the final gather of the outputs (``gather_concat``), kept outside the
  timed data path unless asked for.

The functions work on CPU tensors with the gloo backend too, which is how the host-side logic is
tested without GPUs (tests/test_multigpu_cpu.py).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def _wire(t: torch.Tensor) -> torch.Tensor:
    """Collectives move raw floats: complex64 is viewed as [..., 2] float32 (same memory)."""
    return torch.view_as_real(t) if t.is_complex() else t


def time_segments(n_items: int, world: int, align: int = 1) -> List[Tuple[int, int]]:
    """Split [0, n_items) into `world` contiguous segments whose starts are multiples of `align`
    (decimation D, FFT size N, or channel count M); the last segment takes the remainder."""
    per = (n_items // world) // align * align
    segs = []
    for r in range(world):
        lo = r * per
        hi = n_items if r == world - 1 else (r + 1) * per
        segs.append((lo, hi))
    return segs


def channel_slice(n_channels: int, rank: int, world: int) -> Tuple[int, int]:
    """(channel_begin, channel_count) of this rank for channel sharding."""
    per = n_channels // world
    if per * world != n_channels:
        raise ValueError("n_channels must be divisible by the number of ranks")
    return rank * per, per


def exchange_halo(x_local: torch.Tensor, halo_len: int, rank: int, world: int,
                  group=None) -> Optional[torch.Tensor]:
    """Every rank sends the last `halo_len` items of its segment to rank+1 and receives the halo
    that precedes its own segment from rank-1.  Rank 0 returns None (zeros / stream start)."""
    if halo_len <= 0 or world == 1:
        return None
    if x_local.numel() < halo_len:
        raise ValueError("segment shorter than the halo")
    ops = []
    halo = None
    if rank + 1 < world:
        tail = x_local[-halo_len:].contiguous()
        ops.append(dist.P2POp(dist.isend, _wire(tail), rank + 1, group))
    if rank > 0:
        halo = torch.empty(halo_len, dtype=x_local.dtype, device=x_local.device)
        ops.append(dist.P2POp(dist.irecv, _wire(halo), rank - 1, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return halo


def gather_concat(y_local: torch.Tensor, rank: int, world: int, dst: int = 0, group=None,
                  sizes: Optional[List[int]] = None) -> Optional[torch.Tensor]:
    """Final gather of the per-rank outputs onto `dst` (concatenated along dim 0).  Equal sizes
    use one all-gather-style collective; ragged sizes fall back to point-to-point."""
    if world == 1:
        return y_local
    if sizes is None:
        n = torch.tensor([y_local.shape[0]], dtype=torch.int64, device=y_local.device)
        all_n = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(all_n, n, group=group)
        sizes = [int(t.item()) for t in all_n]
    if len(set(sizes)) == 1:
        out = None
        if rank == dst:
            out = torch.empty((sum(sizes),) + tuple(y_local.shape[1:]), dtype=y_local.dtype,
                              device=y_local.device)
            parts = [_wire(p) for p in out.split(sizes[0], dim=0)]
        else:
            parts = None
        dist.gather(_wire(y_local.contiguous()), parts, dst=dst, group=group)
        return out
    if rank == dst:
        out = torch.empty((sum(sizes),) + tuple(y_local.shape[1:]), dtype=y_local.dtype,
                          device=y_local.device)
        off = 0
        for r in range(world):
            view = out[off:off + sizes[r]]
            if r == dst:
                view.copy_(y_local)
            else:
                dist.recv(_wire(view), src=r, group=group)
            off += sizes[r]
        return out
    dist.send(_wire(y_local.contiguous()), dst=dst, group=group)
    return None


class SegmentedFir:
    """BASELINE config 5: one long stream filtered by `world` GPUs, each owning a time segment.

    `fir` is a newsched_b200.FirFilter.  run(x_local) exchanges the (ntaps-1)-sample halo with the
    left neighbour and launches the stateless segment form of the kernel; outputs stay sharded."""

    def __init__(self, fir, rank: int, world: int, group=None):
        self.fir, self.rank, self.world, self.group = fir, rank, world, group

    def run(self, x_local: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        halo = exchange_halo(x_local, self.fir.n_taps - 1, self.rank, self.world, self.group)
        return self.fir.work_segment(x_local, halo, out)
