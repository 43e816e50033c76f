"""A/B timing of the FFT-shaped kernels for one build of the library (B200_LIB=... selects it)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import newsched_b200 as nb
import scipy.signal as sig
g = torch.Generator(device="cuda").manual_seed(1)
n = 1 << 27
x = torch.view_as_complex(torch.rand(n, 2, device="cuda", generator=g) * 2 - 1)
rng = np.random.default_rng(1)
t = np.arange(4096) / 4095
w = (0.35875 - 0.48829 * np.cos(2 * np.pi * t) + 0.14128 * np.cos(4 * np.pi * t) - 0.01168 * np.cos(6 * np.pi * t)).astype(np.float32)


def taps(T):
    return (rng.uniform(-1, 1, T) / T).astype(np.float32)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


res = {}
op = nb.FFT(4096, True, w, output=nb.OUT_MAG)
o = torch.empty(n, dtype=torch.float32, device="cuda")
res["fftmag"] = n / timeit(lambda: op.work(x, o)) / 1e6
op2 = nb.FFT(4096, True, w)
o2 = torch.empty_like(x)
res["fft"] = n / timeit(lambda: op2.work(x, o2)) / 1e6
op3 = nb.FFT(1024, True, w[:1024].copy())
res["fft1024"] = n / timeit(lambda: op3.work(x, o2)) / 1e6
p = nb.PfbChannelizer(sig.firwin(1024, 1 / 64).astype(np.float32), 64)
op_ = o2.view(-1, 64)
res["pfb64"] = n / timeit(lambda: p.work_segment(x, None, op_)) / 1e6
m = 1 << 26
for name, T, D in (("ols128", 128, 1), ("ols512", 512, 1), ("olsd1024d4", 1024, 4), ("ols2_4096", 4096, 1)):
    f = nb.FirFilter(taps(T), D, algorithm=3)
    res[name] = m / timeit(lambda: f.work_segment(x[:m], None, o2[: m // D]), 10) / 1e6
print(os.environ.get("B200_LIB", "default").split("/")[-1], " ".join(f"{k}={v:.1f}" for k, v in res.items()))
