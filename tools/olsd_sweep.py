"""Decimating complex FIR by overlap-save: polyphase form (fir_olsd_kernel) vs the full-rate form
(B200_OLS_POLY=0), CUDA events, 64 Mi-sample input."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import newsched_b200 as nb
n = 1 << 26
g = torch.Generator(device="cuda").manual_seed(1)
xc = torch.view_as_complex(torch.rand(n, 2, device="cuda", generator=g) * 2 - 1)
rng = np.random.default_rng(1)
for T, D in [(1024, 4), (1024, 2), (1024, 8), (256, 2), (512, 4), (4096, 4), (2048, 8), (768, 3), (1024, 5), (1024, 7)]:
    taps = (rng.uniform(-1, 1, T) / T).astype(np.float32)
    res = []
    for poly in ("0", "1"):
        os.environ["B200_OLS_POLY"] = poly
        f = nb.FirFilter(taps, D, algorithm=3)
        out = torch.empty(n // D, dtype=xc.dtype, device="cuda")
        for _ in range(3): f.work_segment(xc, None, out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(20): f.work_segment(xc, None, out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        res.append((ms, n / (ms * 1e-3) / 1e9, out.clone()))
    err = (res[0][2] - res[1][2]).abs().max().item()
    gs = res[1][1]
    print(f"ccf T={T:5d} D={D}: full-rate {res[0][1]:7.1f} GS/s  polyphase {gs:7.1f} GS/s ({res[1][0]:.4f} ms)  "
          f"x{gs / res[0][1]:.2f}  {gs * 8 * (1 + 1 / D):6.0f} GB/s ({gs * 8 * (1 + 1 / D) / 65.565:4.1f}% hbm)  max|diff| {err:.2e}")
