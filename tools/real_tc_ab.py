"""Float-stream (fff) FIR on the tensor cores (the tap-stationary kernel with two 4096-sample runs per tile) against
the SIMT forms: parity vs the fp64 oracle (one-shot, chunked, fused constant, unaligned stream) and A/B timing.
usage: python tools/real_tc_ab.py [parity|time|all]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TAPS = (16, 32, 33, 48, 64, 96, 128, 192, 256, 384, 448)


def child(kind, T):
    import numpy as np
    import torch
    import newsched_b200 as nb
    rng = np.random.default_rng(T)
    taps = (rng.uniform(-1, 1, T) / T).astype(np.float32)
    if kind == "parity":
        import oracle as o
        n = 8192 * 301 + 1003
        x = rng.uniform(-1, 1, n).astype(np.float32)
        dx = torch.from_numpy(x).cuda()
        ref = o.fir(x, taps, 1)
        f = nb.FirFilter(taps, 1, is_complex=False, algorithm=2)
        y, nc = f.work(dx)
        e1 = o.rel_rms(y.cpu().numpy(), ref)
        f2 = nb.FirFilter(taps, 1, is_complex=False, algorithm=2)
        cut = 8192 * 150 + 333
        ya, _ = f2.work(dx[:cut])
        yb, _ = f2.work(dx[cut:])
        e2 = o.rel_rms(torch.cat([ya, yb]).cpu().numpy(), ref)
        f3 = nb.FirFilter(taps, 1, is_complex=False, algorithm=2, multiply_const=0.5)
        y3, _ = f3.work(dx)
        e3 = o.rel_rms(y3.cpu().numpy(), ref * 0.5)
        xo = torch.empty(n + 1, dtype=torch.float32, device="cuda")
        xo[1:] = dx
        y4, _ = nb.FirFilter(taps, 1, is_complex=False, algorithm=2).work(xo[1:])      # 4-byte aligned stream: element-wise tiles
        e4 = o.rel_rms(y4.cpu().numpy(), ref)
        print(json.dumps({"kind": kind, "T": T, "algorithm": f.algorithm, "rel_rms": e1, "chunked": e2, "fused": e3,
                          "unaligned": e4, "ok": bool(max(e1, e2, e3, e4) < 1e-5 and nc == n)}))
    else:
        n = 1 << 27
        x = torch.rand(n, device="cuda") * 2 - 1
        out = torch.empty(n, dtype=torch.float32, device="cuda")
        res = {"kind": kind, "T": T}
        for algo in (2, 1, 3):
            try:
                f = nb.FirFilter(taps, 1, is_complex=False, algorithm=algo)
            except Exception as e:
                res[f"algo{algo}"] = "unsupported"
                continue
            for _ in range(3):
                f.work_segment(x, None, out)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(10):
                f.work_segment(x, None, out)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            res[f"algo{algo}_Greal_s"] = round(n / ms / 1e6, 1)
            res[f"algo{algo}_hbm"] = round(8 * n / (ms * 1e-3) / 6556.5e9, 3)
        print(json.dumps(res))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what == "child":
        child(sys.argv[2], int(sys.argv[3]))
    else:
        for kind in (["parity", "time"] if what == "all" else [what]):
            for T in TAPS:
                r = subprocess.run([sys.executable, __file__, "child", kind, str(T)], capture_output=True, text=True, timeout=300)
                line = [l for l in r.stdout.splitlines() if l.startswith("{")]
                print(line[-1] if line else json.dumps({"kind": kind, "T": T, "rc": r.returncode, "stderr": r.stderr[-400:]}), flush=True)
