"""Small-size runs of the round-1 late kernels (polyphase / two-phase overlap-save, channelizer DFT,
resampler) for compute-sanitizer memcheck / racecheck."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import newsched_b200 as nb
import oracle as o
rng = np.random.default_rng(3)
def cplx(n): return (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)
n = 4096 * 4 * 5 + 321
x = cplx(n + 1)
for start in (0, 1):
    dx = torch.from_numpy(x).cuda()[start:]
    xs = x[start:]
    for T, D in ((1024, 4), (300, 2), (777, 3), (2048, 1), (4095, 1)):
        taps = (rng.uniform(-1, 1, T) / T).astype(np.float32)
        y, _ = nb.FirFilter(taps, D, algorithm=3).work(dx)
        e = o.rel_rms(y.cpu().numpy(), o.fir(xs, taps, D))
        print("fir ols", T, D, start, e); assert e < 1e-5
import scipy.signal as sig
pt = sig.firwin(1024, 1 / 64).astype(np.float32)
xp = cplx(64 * 700)
yp, _ = nb.PfbChannelizer(pt, 64).work(torch.from_numpy(xp).cuda())
e = o.rel_rms(yp.cpu().numpy().reshape(-1), o.pfb_channelizer(xp, pt, 64).reshape(-1)); print("pfb64", e); assert e < 1e-5
yp, _ = nb.PfbChannelizer(pt, 64, 16, 24).work(torch.from_numpy(xp).cuda())
for T, L, D in ((96, 4, 1), (211, 3, 7), (64, 8, 3)):
    taps = rng.uniform(-1, 1, T).astype(np.float32)
    xr = cplx(50000)
    y, _ = nb.RationalResampler(taps, L, D).work(torch.from_numpy(xr).cuda())
    e = o.rel_rms(y.cpu().numpy(), o.resample(xr, taps, L, D)); print("resample", T, L, D, e); assert e < 1e-5
torch.cuda.synchronize()
print("sanity_new ok")
