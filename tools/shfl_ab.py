"""A/B of the shuffle-based pass-2 -> pass-3 exchange in fft4096_tma_kernel (library built with -DB200_FFT_SHFL=1
selected through B200_LIB): parity against the oracle first, then the timing of tools/pk_ab.py's FFT lines."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import newsched_b200 as nb
import oracle as o
rng = np.random.default_rng(3)
N, nv = 4096, 300
x = (rng.uniform(-1, 1, N * nv) + 1j * rng.uniform(-1, 1, N * nv)).astype(np.complex64)
w = o.window_blackmanharris(N)
dx = torch.from_numpy(x).cuda()
ref = o.fft(x, N, True, w).astype(np.complex128)
y = nb.FFT(N, True, w).work(dx).cpu().numpy()
m = nb.FFT(N, True, w, output=nb.OUT_MAG).work(dx).cpu().numpy()
print("lib", os.environ.get("B200_LIB", "default").split("/")[-2:], "rel rms complex", o.rel_rms(y, ref), "mag", o.rel_rms(m, np.abs(ref)))
g = torch.Generator(device="cuda").manual_seed(1)
n = 1 << 27
xx = torch.view_as_complex(torch.rand(n, 2, device="cuda", generator=g) * 2 - 1)
for name, op, out in (("fft4096+mag", nb.FFT(N, True, w, output=nb.OUT_MAG), torch.empty(n, dtype=torch.float32, device="cuda")),
                      ("fft4096 complex", nb.FFT(N, True, w), torch.empty_like(xx))):
    for _ in range(3): op.work(xx, out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): op.work(xx, out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"  {name}: {ms:.4f} ms  {n/ms/1e6:.1f} GS/s")
