"""One small invocation of every kernel path (for compute-sanitizer memcheck / racecheck runs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import newsched_b200 as nb

rng = np.random.default_rng(0)
def cx(n): return torch.from_numpy((rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)).cuda()
def taps(T): return (rng.uniform(-1, 1, T) / T).astype(np.float32)

x = cx(3 * 2048 * 3 + 77)
xr = torch.view_as_real(x)[:, 0].contiguous()
# elementwise (aligned + misaligned views)
for off in (0, 1):
    v = x[off:off + 5001]
    nb.copy(v); nb.multiply_const(v, 0.5 - 0.25j); nb.complex_to_mag(v); nb.complex_to_mag(v, squared=True)
    nb.multiply(v, v); nb.add(v, v)
nb.multiply_const(xr[:4097], 3.0); nb.multiply(xr[:4097], xr[:4097])
nb.multiply_const(torch.arange(1000, dtype=torch.int16, device="cuda"), 3)
nb.multiply_const(torch.arange(1000, dtype=torch.int32, device="cuda"), 3)
# FIR: direct D=1 (TMA interior + edges), unaligned input (manual), decimating, naive, fff, fused
for T, D, c, algo in ((64, 1, True, 1), (37, 1, True, 1), (64, 1, False, 1), (100, 4, True, 1), (129, 5, False, 1),
                      (300, 40, True, 0), (256, 1, True, 3), (1024, 4, True, 3), (4096, 1, True, 3),
                      (256, 1, False, 3), (3000, 3, False, 3)):
    src = x if c else xr
    f = nb.FirFilter(taps(T), D, is_complex=c, algorithm=algo, multiply_const=(0.5 - 0.25j) if c else 2.0)
    f.work(src)
    f.work(src[1:])          # 8-byte / 4-byte aligned only
    f.work_segment(src[T:], src[1:T])
# FFT: every kernel family, both directions, every output mode
for N in (8, 16, 64, 128, 256, 1024, 2048, 4096, 8192):
    w = rng.uniform(0.1, 1, N).astype(np.float32)
    nv = max(1, 3 * 4096 // N) + 1
    v = cx(nv * N)
    for fwd in (True, False):
        for outm in (nb.OUT_COMPLEX, nb.OUT_MAG):
            nb.FFT(N, fwd, w, shift=True, output=outm, pre_multiply_const=0.5 - 0.25j).work(v)
    nb.FFT(N).work(v[1:1 + (nv - 1) * N])   # not 16-byte aligned: non-TMA paths
# channelizer
import scipy.signal as sig
for M, P in ((64, 16), (64, 5), (16, 8), (256, 4)):
    pt = sig.firwin(M * P, 1.0 / M).astype(np.float32)
    ch = nb.PfbChannelizer(pt, M)
    ch.work(x[: M * 200]); ch.work(x[1: 1 + M * 70])
    nb.PfbChannelizer(pt, M, channel_begin=M // 4, channel_count=M // 2).work(x[: M * 100])
# chain, host streaming, ring
c = nb.Chain([nb.FirFilter(taps(128), 4), ("multiply_const_cc", 0.5), nb.FFT(256, True)], in_item_bytes=8, chunk_items=4096)
hx = x[:8192].cpu().pin_memory(); hy = torch.empty(2048, dtype=torch.complex64).pin_memory()
c.run_host(hx, hy)
ring = nb.DeviceRing(1 << 20)
nb._check(nb.lib().b200_copy(ring.base + ring.size - 4096, x.data_ptr(), 8192, nb._stream()))
torch.cuda.synchronize()
print("sanity_small ok", nb.launch_count(), "launches")
