"""GPU, >= 2 devices (skipped on a 1-GPU box): BASELINE configs 4 and 5 over NCCL.
config 5: one long stream split into time segments, (ntaps-1)-sample halo from the left
neighbour, NCCL gather of the outputs == the single-GPU result.
config 4: channelizer output channels sharded per GPU == the columns of the full output."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    import scipy.signal as sig
    import newsched_b200 as nb
    from newsched_b200 import multigpu as mg
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ok = True
    try:
        rng = np.random.default_rng(11)
        n = 1 << 21
        x = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)
        for T, D, algo in ((64, 1, 1), (4096, 1, 3), (1024, 4, 0)):
            taps = (rng.uniform(-1, 1, T) / T).astype(np.float32)
            lo, hi = mg.time_segments(n, world, D)[rank]
            seg = torch.from_numpy(x[lo:hi]).cuda()
            fir = nb.FirFilter(taps, D, algorithm=algo)
            y = mg.SegmentedFir(fir, rank, world).run(seg)
            full = mg.gather_concat(y, rank, world)
            if rank == 0:
                ref = nb.FirFilter(taps, D, algorithm=algo).work(torch.from_numpy(x).cuda())[0]
                err = (full - ref).abs().pow(2).mean().sqrt() / ref.abs().pow(2).mean().sqrt()
                ok &= bool(full.shape == ref.shape) and float(err) < 1e-5
                if algo == 1:
                    ok &= bool(torch.equal(full, ref))      # direct form: bit-exact across segments
        # config 4: channels sharded, every GPU reads the whole stream
        M, P = 64, 16
        pt = sig.firwin(M * P, 1.0 / M).astype(np.float32)
        xs = torch.from_numpy(x[: M * 4096]).cuda()
        b, c = mg.channel_slice(M, rank, world)
        part = nb.PfbChannelizer(pt, M, channel_begin=b, channel_count=c).work(xs)[0]
        cols = mg.gather_concat(part.t().contiguous(), rank, world)   # gather along channels
        if rank == 0:
            ref = nb.PfbChannelizer(pt, M).work(xs)[0]
            ok &= bool(torch.equal(cols.t(), ref))
        torch.cuda.synchronize()
        if rank == 0:
            q.put(ok)
    finally:
        dist.destroy_process_group()


def test_configs_4_and_5_on_two_gpus():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under `gpurun --gpus 2`)")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
