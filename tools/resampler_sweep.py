"""interp_fir_filter / rational_resampler throughput (CUDA events, 16 Mi-sample complex input)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import newsched_b200 as nb
n = 1 << 24
g = torch.Generator(device="cuda").manual_seed(1)
CPLX = os.environ.get("REAL", "0") != "1"
xc = torch.view_as_complex(torch.rand(n, 2, device="cuda", generator=g) * 2 - 1) if CPLX else torch.rand(n, device="cuda", generator=g) * 2 - 1
rng = np.random.default_rng(1)
fp32, _ = nb.measure_fp32_tflops(8192)
for T, L, D in [(32, 2, 1), (64, 2, 1), (128, 2, 1), (256, 2, 1), (48, 3, 1), (96, 3, 1), (192, 3, 1), (128, 4, 1), (256, 8, 1), (96, 3, 2), (320, 5, 4), (1024, 16, 1), (128, 2, 3)] if 'RATIOS' not in os.environ and 'L4' not in os.environ else [(T, 4, 1) for T in (32, 64, 128, 192, 256, 512)] if 'L4' in os.environ else [(T * L, L, D) for (L, D) in ((3, 2), (2, 3), (4, 3), (3, 4), (5, 4), (4, 5), (5, 3), (3, 5), (5, 2), (2, 5)) for T in (16, 32, 64)]:
    taps = (rng.uniform(-1, 1, T) / T).astype(np.float32)
    r = nb.RationalResampler(taps, L, D, is_complex=CPLX)
    out = torch.empty((n // D) * L, dtype=xc.dtype, device="cuda")
    for _ in range(3): r.work_segment(xc, None, out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): r.work_segment(xc, None, out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gs_in, gs_out = n / ms / 1e6, out.numel() / ms / 1e6
    tq = -(-T // L)
    tf = gs_out * (4 if CPLX else 2) * tq / 1e3
    gb = (n + out.numel()) * 8 / ms / 1e6
    print(f"{'ccf' if CPLX else 'fff'} T={T:5d} L={L:2d} D={D}: {ms:7.4f} ms  in {gs_in:6.1f} GS/s  out {gs_out:6.1f} GS/s  "
          f"{tf:5.1f} TF ({tf / fp32 * 100:4.1f}% fp32)  {gb:6.0f} GB/s ({gb / 65.565:4.1f}% hbm)")
