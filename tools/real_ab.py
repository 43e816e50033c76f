"""Full-rate real FIR (fff): float-pair form vs the scalar kernel (B200_FIR_REAL_SCALAR=1) vs overlap-save."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import newsched_b200 as nb
n = 1 << 27
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.rand(n, device="cuda", generator=g) * 2 - 1
out = torch.empty_like(x)
rng = np.random.default_rng(1)
def rate(f):
    for _ in range(3): f.work_segment(x, None, out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): f.work_segment(x, None, out)
    e1.record(); torch.cuda.synchronize()
    return n / (e0.elapsed_time(e1) / 10) / 1e6
for T in (16, 32, 64, 96, 128, 192, 256, 384):
    taps = (rng.uniform(-1, 1, T) / T).astype(np.float32)
    os.environ.pop("B200_FIR_REAL_SCALAR", None)
    a = rate(nb.FirFilter(taps, 1, is_complex=False, algorithm=1))
    os.environ["B200_FIR_REAL_SCALAR"] = "1"
    b = rate(nb.FirFilter(taps, 1, is_complex=False, algorithm=1))
    os.environ.pop("B200_FIR_REAL_SCALAR", None)
    c = rate(nb.FirFilter(taps, 1, is_complex=False, algorithm=3))
    print(f"fff T={T:4d}: pairs {a:6.1f}  scalar {b:6.1f}  overlap-save {c:6.1f}  G real samples/s", flush=True)
