"""FIR throughput sweep over taps / decimation / type (CUDA events, 16 Mi-sample input)."""
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
cases = [(True, 8, 1), (True, 16, 1), (True, 32, 1), (True, 64, 1), (True, 96, 1), (True, 128, 1), (True, 64, 2), (True, 128, 2),
         (True, 256, 4), (True, 128, 4), (True, 64, 4), (True, 512, 8), (False, 64, 1), (False, 128, 1), (False, 256, 1), (False, 64, 2), (False, 1024, 4)]
for cplx, T, D in cases:
    for algo in ((1, 3) if cplx and T >= 32 else (1,)):
        x = xc if cplx else xf
        taps = (rng.uniform(-1, 1, T) / T).astype(np.float32)
        try:
            f = nb.FirFilter(taps, D, is_complex=cplx, algorithm=algo)
        except Exception as e:
            print(cplx, T, D, algo, "unsupported"); continue
        out = torch.empty(x.numel() // D, dtype=x.dtype, device="cuda")
        for _ in range(3): f.work_segment(x, None, out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(10): f.work_segment(x, None, out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        ns = x.numel()
        gs = ns / (ms * 1e-3) / 1e9
        flop = (4 if cplx else 2) * T / D
        bytes_ = (8 if cplx else 4) * (1 + 1 / D)
        print(f"{'ccf' if cplx else 'fff'} T={T:5d} D={D} algo={f.algorithm} {ms:8.4f} ms {gs:8.1f} GS/s  direct-equiv {gs*flop/1e3:7.1f} TF ({gs*flop/1e3/fp32*100:5.1f}% fp32)  {gs*bytes_:7.0f} GB/s ({gs*bytes_/6556.5*100:5.1f}% hbm)")
