"""Decimating complex FIR: direct kernel (algorithm 1; B200_FIR_PLANES=1 selects the old phase-plane
staging) vs polyphase overlap-save (3), input rate in GS/s, 64 Mi-sample input."""
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
CPLX = os.environ.get("REAL", "0") != "1"
if not CPLX:
    x = torch.rand(2 * n, device="cuda", generator=g) * 2 - 1
    n = 2 * n
DS = tuple(int(d) for d in os.environ["DS"].split(",")) if "DS" in os.environ else (2, 4, 8, 16) + (() if CPLX else (32,))
for D in DS:
    line = f"{'ccf' if CPLX else 'fff'} D={D:2d}:"
    for T in (32, 64, 128, 192, 256, 384, 512, 768):
        taps = (rng.uniform(-1, 1, T) / T).astype(np.float32)
        os.environ.pop("B200_FIR_PLANES", None)
        a = rate(nb.FirFilter(taps, D, is_complex=CPLX, algorithm=1), D)
        os.environ["B200_FIR_PLANES"] = "1"
        try:
            b = rate(nb.FirFilter(taps, D, is_complex=CPLX, algorithm=1), D)
        except Exception:
            b = float("nan")
        os.environ.pop("B200_FIR_PLANES", None)
        try:
            c = rate(nb.FirFilter(taps, D, is_complex=CPLX, algorithm=3), D)
        except Exception:
            c = float("nan")
        line += f"  T{T}: {a:.0f}/{b:.0f}/{c:.0f}"
    print(line, flush=True)
