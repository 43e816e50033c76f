"""Bottleneck attribution of the pipelined tensor-core FIR: times the kernel with roles switched off
(B200_TC_DBG bit 1 = no conversion, 2 = no epilogue, 4 = no MMAs; results are garbage then)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import newsched_b200 as nb
n = 1 << 26
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.view_as_complex(torch.rand(n, 2, device="cuda", generator=g) * 2 - 1)
out = torch.empty(n, dtype=torch.complex64, device="cuda")
rng = np.random.default_rng(1)
for T in [int(t) for t in (sys.argv[1:] or ["64", "256"])]:
    taps = (rng.uniform(-1, 1, T) / T).astype(np.float32)
    f = nb.FirFilter(taps, 1, algorithm=2)
    for dbg in (0, 1, 2, 4, 8, 9, 11, 13, 6, 7, 14, 15):
        os.environ["B200_TC_DBG"] = str(dbg)
        for _ in range(2): f.work_segment(x, None, out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(5): f.work_segment(x, None, out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"T={T} dbg={dbg} (load={'off' if dbg&8 else 'on'} conv={'off' if dbg&1 else 'on'} epi={'off' if dbg&2 else 'on'} mma={'off' if dbg&4 else 'on'}) {ms:.3f} ms {n/ms/1e6:.1f} GS/s  us per 4096-tile per SM {ms*1e3/(n/4096/148):.2f}", flush=True)
    os.environ["B200_TC_DBG"] = "0"
