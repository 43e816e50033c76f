"""64-channel channelizer: the DFT across branches as a tcgen05 GEMM (algorithm 2) against the SIMT DFT
(algorithm 1) -- parity vs the fp64 oracle and A/B timing on a 1 GiB stream.  Every case runs in its own
process, so a trapped kernel poisons only that case.
usage: python tools/pfb_tc_ab.py [parity|time|all]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(kind, P, algo):
    import numpy as np
    import scipy.signal as sig
    import torch
    import newsched_b200 as nb
    M = 64
    taps = sig.firwin(M * P, 1.0 / M).astype(np.float32)
    if kind == "parity":
        import oracle as o
        rng = np.random.default_rng(P)
        n = M * (64 * 301 + 17)
        x = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)
        y, _ = nb.PfbChannelizer(taps, M, algorithm=algo).work(torch.from_numpy(x).cuda())
        torch.cuda.synchronize()
        ref = o.pfb_channelizer(x, taps, M)
        yh = y.cpu().numpy().astype(np.complex128)
        per_ch = np.sqrt((np.abs(yh - ref) ** 2).sum(axis=0) / (np.abs(ref) ** 2).sum(axis=0))
        print(json.dumps({"kind": kind, "P": P, "algorithm": algo, "rel_rms": o.rel_rms(y.cpu().numpy(), ref),
                          "worst_channel_rel_rms": float(per_ch.max()), "worst_channel": int(per_ch.argmax())}))
    else:
        n = 1 << 27
        g = torch.Generator(device="cuda").manual_seed(1)
        x = torch.view_as_complex(torch.rand(n, 2, device="cuda", generator=g) * 2 - 1)
        out = torch.empty((n // M, M), dtype=torch.complex64, device="cuda")
        f = nb.PfbChannelizer(taps, M, algorithm=algo)
        for _ in range(3):
            f.work_segment(x, None, out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            f.work_segment(x, None, out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(json.dumps({"kind": kind, "P": P, "algorithm": algo, "ms": ms, "GS_s": n / (ms * 1e-3) / 1e9,
                          "frac_of_hbm": 16.0 * n / (ms * 1e-3) / 6556.5e9}))


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what == "child":
        child(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]))
        return
    kinds = ["parity", "time"] if what == "all" else [what]
    for kind in kinds:
        for P in (16, 8, 4, 12):
            for algo in (1, 2):
                r = subprocess.run([sys.executable, __file__, "child", kind, str(P), str(algo)], capture_output=True,
                                   text=True, timeout=300)
                line = [l for l in r.stdout.splitlines() if l.startswith("{")]
                print(line[-1] if line else json.dumps({"kind": kind, "P": P, "algorithm": algo, "rc": r.returncode,
                                                        "stderr": r.stderr[-400:]}), flush=True)


if __name__ == "__main__":
    main()
