"""Multi-GPU partitioning of the hot path: one process per GPU over
``torch.distributed``; the blocks shard with NO data-path collective.

* by stream / channel: independent per-GPU streams, or a slice of the channelizer's output
  channels per GPU (``channel_slice``);
* by time segment (BASELINE config 5): contiguous segments aligned to the decimation, each rank
  needs the ``ntaps-1`` samples before its segment.  On GPUs the left neighbour's segment buffer is
  mapped into this process once (``PeerHalo``: CUDA IPC, peer access over NVLink) and the kernels read
  the halo in place -- no exchange step, no copy, nothing in front of the kernel.  ``exchange_halo``
  (point-to-point send/recv: NCCL on GPUs, gloo in the CPU tests) is the portable form;
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


def time_segments(n_items: int, world: int, align: int = 1, halo_len: int = 0) -> List[Tuple[int, int]]:
    """Split [0, n_items) into `world` contiguous segments whose starts are multiples of `align`
    (decimation D, FFT size N, or channel count M); the last segment takes the remainder.
    Raises (identically on every rank: only global quantities are used) when a segment would be
    empty or shorter than the halo its right neighbour needs."""
    per = (n_items // world) // align * align
    if world > 1 and (per <= 0 or per < halo_len):
        raise ValueError(f"{n_items} items over {world} ranks (align {align}) leave segments of {per} items; "
                         f"need at least max(1, halo {halo_len})")
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
    that precedes its own segment from rank-1.  Rank 0 returns None (zeros / stream start).
    The segment lengths must have been validated collectively first (`time_segments(..., halo_len=)`
    or `check_segments`): a rank that raised here on its own would leave its neighbours blocked."""
    if halo_len <= 0 or world == 1:
        return None
    if x_local.numel() < halo_len:
        raise ValueError("segment shorter than the halo (validate with check_segments() before any exchange)")
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


def check_segments(n_local: int, halo_len: int, world: int, device=None, group=None) -> None:
    """Collective validation (one small all-reduce, every rank calls it): raises on ALL ranks when any
    rank's segment is shorter than the halo, so nobody is left waiting in a point-to-point call."""
    if world == 1:
        return
    t = torch.tensor([n_local], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    if int(t.item()) < halo_len:
        raise ValueError(f"shortest segment has {int(t.item())} items, the halo needs {halo_len}")


class PeerHalo:
    """The halo read in place: maps the left neighbour's segment buffer into this process (CUDA IPC; the
    mapping enables peer access, so kernels on this GPU load it over NVLink) and exposes the device
    address of its last `halo_len` items.  Setup is collective and happens once per buffer; afterwards a
    step is just the kernel launch.  The neighbour must keep the buffer alive and must have finished
    writing it before this rank's kernel runs (a barrier / event of the caller, as for any shared buffer)."""

    def __init__(self, x_local: torch.Tensor, halo_len: int, rank: int, world: int, group=None):
        import newsched_b200 as nb
        self.ptr: Optional[int] = None
        self._raw = None
        self._base = None
        self.key = (x_local.data_ptr(), x_local.numel())
        if world == 1 or halo_len <= 0:
            return
        infos = [None] * world
        dist.all_gather_object(infos, (nb.ipc_export(x_local), x_local.numel(), x_local.element_size()), group=group)
        short = min(n for _, n, _ in infos)
        if short < halo_len:      # same data on every rank: everybody raises
            raise ValueError(f"shortest segment has {short} items, the halo needs {halo_len}")
        if rank > 0:
            raw, n_left, item = infos[rank - 1]
            self._raw = raw
            self._base = nb.ipc_import(raw)
            self.ptr = self._base + (n_left - halo_len) * item

    def close(self):
        if self._base is not None:
            import newsched_b200 as nb
            nb.ipc_close(self._raw, self._base)
            self._base = None
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


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

    `fir` is a newsched_b200.FirFilter (or PfbChannelizer: anything with work_segment(x, halo, out)).
    run(x_local) launches the stateless segment form of the kernel with the `halo_len` items that
    precede the segment; outputs stay sharded.  peer=True (CUDA tensors): the halo is read in place from
    the left neighbour's buffer (PeerHalo; the first run() on a buffer is collective).  peer=False: the
    halo is exchanged point-to-point every run (NCCL / gloo)."""

    def __init__(self, fir, rank: int, world: int, group=None, peer: bool = False, halo_len: Optional[int] = None):
        self.fir, self.rank, self.world, self.group, self.peer = fir, rank, world, group, peer
        self.halo_len = fir.n_taps - 1 if halo_len is None else halo_len
        self._peer: Optional[PeerHalo] = None
        self._checked = None

    def run(self, x_local: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if self.peer and x_local.is_cuda:
            if self._peer is None or self._peer.key != (x_local.data_ptr(), x_local.numel()):
                if self._peer is not None:
                    self._peer.close()
                self._peer = PeerHalo(x_local, self.halo_len, self.rank, self.world, self.group)
                if self.world > 1:
                    dist.barrier(group=self.group)      # every neighbour's buffer is mapped and written
            return self.fir.work_segment(x_local, self._peer.ptr, out)
        if self._checked != x_local.numel():
            check_segments(x_local.numel(), self.halo_len, self.world,
                           x_local.device if x_local.is_cuda else None, self.group)
            self._checked = x_local.numel()
        halo = exchange_halo(x_local, self.halo_len, self.rank, self.world, self.group)
        return self.fir.work_segment(x_local, halo, out)
