import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, scipy.signal as sig
import newsched_b200 as nb
n = 1 << 26
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.view_as_complex(torch.rand(n, 2, device="cuda", generator=g) * 2 - 1)
for M, P in ((64, 16), (64, 8), (64, 32), (64, 10), (16, 8), (16, 16), (32, 16), (32, 8), (128, 8), (128, 16), (256, 4), (256, 8), (8, 16), (8, 8), (4, 8), (4, 16)):
    pt = sig.firwin(M * P, 1.0 / M).astype(np.float32)
    ch = nb.PfbChannelizer(pt, M)
    out = torch.empty((n // M, M), dtype=torch.complex64, device="cuda")
    for _ in range(2): ch.work_segment(x, None, out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): ch.work_segment(x, None, out)
    e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 5
    print(f"pfb M={M:4d} P={P:3d} {ms:8.3f} ms {n/ms/1e6:8.1f} GS/s  {n*16/ms/1e6/6556.5*100:5.1f}% hbm")
