"""Direct (1) vs one-phase overlap-save (3) around the crossover, complex stream, D = 1."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import newsched_b200 as nb
g = torch.Generator(device="cuda").manual_seed(1)
rng = np.random.default_rng(1)
for logn in (22, 24, 27):
    n = 1 << logn
    x = torch.view_as_complex(torch.rand(n, 2, device="cuda", generator=g) * 2 - 1)
    out = torch.empty_like(x)
    line = f"2^{logn}:"
    for T in (56, 64, 72, 80, 88, 96):
        r = []
        for algo in (1, 3):
            f = nb.FirFilter((rng.uniform(-1, 1, T) / T).astype(np.float32), 1, algorithm=algo)
            for _ in range(3): f.work_segment(x, None, out)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(20): f.work_segment(x, None, out)
            e1.record(); torch.cuda.synchronize()
            r.append(n / (e0.elapsed_time(e1) / 20) / 1e6)
        line += f"  T{T}: {r[0]:.0f}/{r[1]:.0f}"
    print(line)
    del x, out
