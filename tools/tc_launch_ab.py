"""Per-launch overhead of the tap-stationary tensor-core FIR at config 1's size (16 Mi samples, 64 taps): A/B of the
programmatic dependent launch (B200_TC_PDL) and of the staged head tile (B200_TC_TS_HEAD), one process per variant.
usage: python tools/tc_launch_ab.py"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import numpy as np
    import torch
    import newsched_b200 as nb
    rng = np.random.default_rng(7)
    res = {"pdl": os.environ.get("B200_TC_PDL", "1"), "head": os.environ.get("B200_TC_TS_HEAD", "1")}
    for T in (64, 128, 256):
        taps = (rng.uniform(-1, 1, T) / T).astype(np.float32)
        for logn in (24, 26):
            n = 1 << logn
            g = torch.Generator(device="cuda").manual_seed(1)
            x = torch.view_as_complex(torch.rand(n, 2, device="cuda", generator=g) * 2 - 1)
            out = torch.empty(n, dtype=torch.complex64, device="cuda")
            f = nb.FirFilter(taps, 1, algorithm=2)
            for mode in ("segment", "work"):
                fn = (lambda: f.work_segment(x, None, out)) if mode == "segment" else (lambda: f.work(x, out))
                for _ in range(5):
                    fn()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(40):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) / 40 * 1e3
                res[f"T{T}_2^{logn}_{mode}"] = {"us": round(us, 2), "GS_s": round(n / us / 1e3, 1),
                                                "frac_hbm": round(16 * n / (us * 1e-6) / 6556.5e9, 3)}
    print(json.dumps(res))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        for pdl in ("0", "1"):
            for head in ("0", "1"):
                env = dict(os.environ, B200_TC_PDL=pdl, B200_TC_TS_HEAD=head)
                r = subprocess.run([sys.executable, __file__, "child"], capture_output=True, text=True, env=env, timeout=300)
                line = [l for l in r.stdout.splitlines() if l.startswith("{")]
                print(line[-1] if line else json.dumps({"pdl": pdl, "head": head, "rc": r.returncode, "stderr": r.stderr[-500:]}),
                      flush=True)
