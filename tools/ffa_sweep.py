"""Full-rate FIR: direct form (algorithm 1) vs 2-parallel fast FIR (algorithm 5) vs overlap-save (3),
CUDA events, 16 Mi-sample input."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import newsched_b200 as nb
n = 1 << 24
g = torch.Generator(device="cuda").manual_seed(1)
xc = torch.view_as_complex(torch.rand(n, 2, device="cuda", generator=g) * 2 - 1)
xf = torch.rand(2 * n, device="cuda", generator=g) * 2 - 1
rng = np.random.default_rng(1)
fp32, _ = nb.measure_fp32_tflops(8192)
print("fp32 peak", fp32)
for cplx, T in [(True, 16), (True, 32), (True, 48), (True, 64), (True, 96), (True, 128), (True, 192),
                (False, 32), (False, 64), (False, 128), (False, 256), (False, 512)]:
    x = xc if cplx else xf
    taps = (rng.uniform(-1, 1, T) / T).astype(np.float32)
    line = f"{'ccf' if cplx else 'fff'} T={T:4d}:"
    for algo in (1, 5, 3):
        try:
            f = nb.FirFilter(taps, 1, is_complex=cplx, algorithm=algo)
        except Exception:
            line += f"  algo{algo}    n/a      "; continue
        out = torch.empty_like(x)
        for _ in range(3): f.work_segment(x, None, out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(20): f.work_segment(x, None, out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        gs = x.numel() / ms / 1e6
        tf = gs * (4 if cplx else 2) * T / 1e3
        line += f"  algo{algo} {gs:6.1f} GS/s ({tf / fp32 * 100:5.1f}% fp32)"
    print(line)
