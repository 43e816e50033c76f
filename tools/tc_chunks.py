"""Per-call cost of the FIR block for work()-sized chunks: streaming calls (kernel + history kernel) back to back,
tensor-core form (2) against the SIMT direct form (1) at 64 taps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import newsched_b200 as nb
g = torch.Generator(device="cuda").manual_seed(1)
n = 1 << 26
x = torch.view_as_complex(torch.rand(n, 2, device="cuda", generator=g) * 2 - 1)
out = torch.empty(n, dtype=torch.complex64, device="cuda")
rng = np.random.default_rng(1)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 64
taps = (rng.uniform(-1, 1, T) / T).astype(np.float32)
for algo in (2, 1, 3):
    f = nb.FirFilter(taps, 1, algorithm=algo)
    for chunk in (1 << 16, 1 << 18, 1 << 20, 1 << 22, 1 << 24, 1 << 26):
        reps = 20
        for _ in range(3): f.work(x[:chunk], out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(reps): f.work(x[:chunk], out)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3
        print(f"T={T} algo={f.algorithm} chunk={chunk:>9d}: {us:8.1f} us/call  {chunk/us/1e3:7.1f} GS/s", flush=True)
