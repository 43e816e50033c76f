"""Folded decimation by 3/5/6/7 (direct kernel, algorithm 1) for one build of the library: input rate, GS/s."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import newsched_b200 as nb
g = torch.Generator(device="cuda").manual_seed(1)
rng = np.random.default_rng(1)
n = 1 << 26
x = torch.view_as_complex(torch.rand(n, 2, device="cuda", generator=g) * 2 - 1)
def rate(f, D):
    out = torch.empty(n // D, dtype=x.dtype, device="cuda")
    for _ in range(3): f.work_segment(x, None, out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): f.work_segment(x, None, out)
    e1.record(); torch.cuda.synchronize()
    return n / (e0.elapsed_time(e1) / 10) / 1e6
for D in (3, 5, 6, 7):
    line = f"D={D}:"
    for T in (16, 32, 64, 128, 192, 256):
        taps = (rng.uniform(-1, 1, T) / T).astype(np.float32)
        line += f" T{T}={rate(nb.FirFilter(taps, D, algorithm=1), D):.0f}"
    print(os.environ.get("B200_LIB", "default").split("/")[-1], line, flush=True)
