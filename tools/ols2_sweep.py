"""Full-rate long-tap complex FIR: two-phase overlap-save (fir_ols2_kernel) vs the one-phase /
partitioned form (B200_OLS_TWO=0), CUDA events, 64 Mi-sample input."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import newsched_b200 as nb
n = 1 << 26
g = torch.Generator(device="cuda").manual_seed(1)
xc = torch.view_as_complex(torch.rand(n, 2, device="cuda", generator=g) * 2 - 1)
rng = np.random.default_rng(1)
for T in (256, 512, 1024, 1536, 2048, 3072, 4096, 6000):
    taps = (rng.uniform(-1, 1, T) / T).astype(np.float32)
    res = []
    for two in ("0", "1"):
        os.environ["B200_OLS_TWO"] = two   # 1 = two-phase from 1 tap on
        f = nb.FirFilter(taps, 1, algorithm=3)
        out = torch.empty(n, dtype=xc.dtype, device="cuda")
        for _ in range(3): f.work_segment(xc, None, out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(10): f.work_segment(xc, None, out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        res.append((ms, n / (ms * 1e-3) / 1e9, out.clone()))
    err = (res[0][2] - res[1][2]).abs().max().item()
    print(f"ccf T={T:5d} D=1: one-phase {res[0][1]:7.1f} GS/s  two-phase {res[1][1]:7.1f} GS/s ({res[1][0]:.4f} ms)  "
          f"x{res[1][1] / res[0][1]:.2f}  max|diff| {err:.2e}")
