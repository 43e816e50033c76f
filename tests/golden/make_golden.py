"""Generates tests/golden/*.npz -- small known-answer fixtures for the hot path.

Two kinds of vectors:

* the reference's own deterministic test inputs (the only known-answer data mormj/newsched
  holds for this path): ramp (2i, 2i+1) of schedulers/mt/test/qa_scheduler_mt.cpp:86-88 with
  multiply_const_cc(k=1) expected == input (:128-132), and ramp (i, -i), veclen 1024, of
  schedulers/mt/test/cuda/qa_scheduler_mt_cuda_copy.cpp:24-28 with copy expected == input;
* for the blocks the snapshot does not contain (fir/fft/complex_to_mag/pfb),
  seeded inputs with float64 numpy/scipy expectations (np.convolve, np.fft, scipy windows),
  i.e. a statement of SURVEY.md 8(c) that is independent of both oracle.c and the CUDA code.

The reference itself (C++/meson/VOLK/flatbuffers) cannot be built or imported in this image,
so no vector here is an output of reference code.  Run:  python tests/golden/make_golden.py
"""
import os

import numpy as np
import scipy.signal as sig
import scipy.signal.windows as win

HERE = os.path.dirname(os.path.abspath(__file__))


def cplx(rng, n):
    return (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)


def np_fir(x, taps, D, hist=None):
    T = len(taps)
    xd = x.astype(np.complex128 if np.iscomplexobj(x) else np.float64)
    pre = np.zeros(T - 1, xd.dtype) if hist is None else hist.astype(xd.dtype)
    y = sig.lfilter(taps.astype(np.float64), [1.0], np.concatenate([pre, xd]))[T - 1:]
    return y[::D][: len(x) // D]


def main():
    rng = np.random.default_rng(0x5EED)
    out = {}
    # --- reference ramps
    i = np.arange(4096, dtype=np.float32)
    out["ramp_qa_scheduler_mt"] = (2 * i + 1j * (2 * i + 1)).astype(np.complex64)
    j = np.arange(1024 * 4, dtype=np.float32)
    out["ramp_qa_cuda_copy"] = (j - 1j * j).astype(np.complex64)
    # --- multiply_const, k != 1 (float64 expectation)
    x = cplx(rng, 2048)
    out["mulc_x"] = x
    out["mulc_k"] = np.array([0.5 - 0.25j], np.complex64)
    out["mulc_y64"] = x.astype(np.complex128) * np.complex128(out["mulc_k"][0])
    out["mag_y64"] = np.abs(x.astype(np.complex128))
    # --- FIR: ccf 64 taps D=1, 37 taps D=3, fff 129 taps D=4, with and without history
    for name, T, D, real in (("fir_a", 64, 1, False), ("fir_b", 37, 3, False), ("fir_c", 129, 4, True)):
        taps = (rng.uniform(-1, 1, T) / T).astype(np.float32)
        xs = cplx(rng, 3000)
        if real:
            xs = xs.real.copy()
        hist = xs[: T - 1].copy()
        out[name + "_taps"] = taps
        out[name + "_x"] = xs
        out[name + "_D"] = np.array([D])
        out[name + "_y64"] = np_fir(xs, taps, D)
        out[name + "_y64_hist"] = np_fir(xs[T - 1:], taps, D, hist)
    # impulse -> taps
    taps = sig.firwin(48, 0.2).astype(np.float32)
    imp = np.zeros(128, np.complex64)
    imp[0] = 1.0
    out["fir_imp_taps"] = taps
    out["fir_imp_y64"] = np_fir(imp, taps, 1)
    # --- FFT 4096 + Blackman-Harris, fwd/rev x shift; and N=64, 1024 no window
    w = win.blackmanharris(4096, sym=True)
    out["bh4096"] = w
    xs = cplx(rng, 2 * 4096)
    out["fft_x"] = xs
    w32 = w.astype(np.float32)
    v = xs.reshape(2, 4096)
    vw = (v.real * w32 + 1j * (v.imag * w32)).astype(np.complex64).astype(np.complex128)
    out["fft_fwd"] = np.fft.fft(vw, axis=1).reshape(-1)
    out["fft_fwd_shift"] = np.fft.fftshift(np.fft.fft(vw, axis=1), axes=1).reshape(-1)
    out["fft_rev"] = (np.fft.ifft(vw, axis=1) * 4096).reshape(-1)
    vs = np.fft.ifftshift(v, axes=1)
    vsw = (vs.real * w32 + 1j * (vs.imag * w32)).astype(np.complex64).astype(np.complex128)
    out["fft_rev_shift"] = (np.fft.ifft(vsw, axis=1) * 4096).reshape(-1)
    for N in (64, 1024):
        xs = cplx(rng, 3 * N)
        out[f"fft{N}_x"] = xs
        out[f"fft{N}_fwd"] = np.fft.fft(xs.reshape(3, N).astype(np.complex128), axis=1).reshape(-1)
    # --- PFB channelizer M=64 P=16 (+ tone test data) and M=8 P=5
    for name, M, P in (("pfb_a", 64, 16), ("pfb_b", 8, 5)):
        taps = sig.firwin(M * P, 1.0 / M).astype(np.float32)
        xs = cplx(rng, M * 40)
        xx = np.concatenate([np.zeros((P - 1) * M, np.complex128), xs.astype(np.complex128)])
        t = np.arange(40)
        ii = np.arange(M)
        u = np.zeros((40, M), np.complex128)
        for r in range(P):
            idx = (P - 1) * M + (t[:, None] - r) * M + (M - 1 - ii)[None, :]
            u += taps[ii + r * M].astype(np.float64)[None, :] * xx[idx]
        out[name + "_taps"] = taps
        out[name + "_x"] = xs
        out[name + "_M"] = np.array([M])
        out[name + "_y64"] = np.fft.ifft(u, axis=1) * M
    np.savez_compressed(os.path.join(HERE, "hotpath_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "hotpath_golden.npz"),
          os.path.getsize(os.path.join(HERE, "hotpath_golden.npz")), "bytes")


if __name__ == "__main__":
    main()
