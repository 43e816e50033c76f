"""A/B timing of the direct FIR kernel for one build of the library (B200_LIB=... selects it)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import newsched_b200 as nb
g = torch.Generator(device="cuda").manual_seed(1)
rng = np.random.default_rng(1)
res = {}
for logn in (24, 27):
    n = 1 << logn
    x = torch.view_as_complex(torch.rand(n, 2, device="cuda", generator=g) * 2 - 1)
    out = torch.empty_like(x)
    for T in (16, 32, 48, 64):
        f = nb.FirFilter((rng.uniform(-1, 1, T) / T).astype(np.float32), 1, algorithm=1)
        for _ in range(3): f.work_segment(x, None, out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(20): f.work_segment(x, None, out)
        e1.record(); torch.cuda.synchronize()
        res[f"T{T}_2^{logn}"] = n / (e0.elapsed_time(e1) / 20) / 1e6
    del x, out
print(os.environ.get("B200_LIB", "default").split("/")[-1], " ".join(f"{k}={v:.1f}" for k, v in res.items()))
