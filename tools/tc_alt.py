"""Alternating big / tiny streaming calls (the pattern a ring with aligned windows produces), per-pair cost."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import newsched_b200 as nb
g = torch.Generator(device="cuda").manual_seed(1)
n = 1 << 24
x = torch.view_as_complex(torch.rand(n, 2, device="cuda", generator=g) * 2 - 1)
out = torch.empty(n, dtype=torch.complex64, device="cuda")
rng = np.random.default_rng(1)
taps = (rng.uniform(-1, 1, 64) / 64).astype(np.float32)
for algo in (2, 1):
    f = nb.FirFilter(taps, 1, algorithm=algo)
    for big, small in ((1 << 23, 3000), (1 << 23, 15), (1 << 23, 0), ((1 << 23) - 16, 16), (1 << 20, 1 << 20)):
        for _ in range(3):
            f.work(x[:big], out); small and f.work(x[big:big + small], out)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            f.work(x[:big], out)
            if small: f.work(x[big:big + small], out)
        e1.record(); torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / 50 * 1e6
        print(f"algo={f.algorithm} big={big} small={small}: {e0.elapsed_time(e1)/50*1e3:8.1f} us/pair on device, {wall:8.1f} us wall", flush=True)
