"""fft8192_kernel A/B: one process per library build (B200_LIB), complex and |.| output, parity vs numpy on a sample.
usage: python tools/f8k_ab.py   (the variant library is built with
       make -C newsched_b200/csrc BUILD=build_f8k OUT=../../tools/variants/f8k/libb200dsp.so EXTRA=-DB200_F8K_EARLY=1)"""
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
    N = 8192
    n = 1 << 27
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.view_as_complex(torch.rand(n, 2, device="cuda", generator=g) * 2 - 1)
    t = np.arange(N) / (N - 1)
    w = (0.35875 - 0.48829 * np.cos(2 * np.pi * t) + 0.14128 * np.cos(4 * np.pi * t) - 0.01168 * np.cos(6 * np.pi * t)).astype(np.float32)
    res = {"lib": os.environ.get("B200_LIB", "default")}
    for name, out_kind, odt in (("complex", nb.OUT_COMPLEX, torch.complex64), ("mag", nb.OUT_MAG, torch.float32)):
        op = nb.FFT(N, True, w, output=out_kind)
        o = torch.empty(n, dtype=odt, device="cuda")
        for _ in range(3):
            op.work(x, o)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            op.work(x, o)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        ref = np.fft.fft(x[:N].cpu().numpy().astype(np.complex128) * w)
        got = o[:N].cpu().numpy()
        if name == "mag":
            ref = np.abs(ref)
        err = float(np.sqrt((np.abs(got - ref) ** 2).sum() / (np.abs(ref) ** 2).sum()))
        bytes_per = 16 if name == "complex" else 12
        res[name] = {"GS_s": round(n / ms / 1e6, 1), "frac_hbm": round(bytes_per * n / (ms * 1e-3) / 6556.5e9, 3), "rel_rms": err}
    print(json.dumps(res))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        for lib in (None, os.path.join(ROOT, "tools", "variants", "f8k", "libb200dsp.so")):
            env = dict(os.environ)
            if lib:
                env["B200_LIB"] = lib
            r = subprocess.run([sys.executable, __file__, "child"], capture_output=True, text=True, env=env, timeout=300)
            line = [l for l in r.stdout.splitlines() if l.startswith("{")]
            print(line[-1] if line else r.stderr[-500:], flush=True)
