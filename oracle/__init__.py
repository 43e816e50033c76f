"""CPU oracle for the newsched data-parallel block hot path.

TEST INFRASTRUCTURE ONLY.  Importable from ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- never from ``newsched_b200``.

Two layers:

* ``oracle.c`` (built to ``oracle/_build/liboracle.so`` by ``oracle/Makefile``): the C
  restatement, bound here through ctypes (functions ``copy``, ``multiply_const_*``,
  ``complex_to_mag``, ``fir``, ``fft``, ``pfb_channelizer`` ...).
* ``np_*`` functions: an independent numpy/float64 statement of the same definitions
  (SURVEY.md 8c), used only to pin the C code in ``tests/test_oracle.py``.

Parity status: ``copy`` and ``multiply_const`` (k = 1) are pinned by the reference's own
known-answer tests (ramps of ``schedulers/mt/test/qa_scheduler_mt.cpp:86-88`` and
``schedulers/mt/test/cuda/qa_scheduler_mt_cuda_copy.cpp:24-28``); everything else is
"parity unpinned" because the mounted reference snapshot has no FIR/FFT/complex_to_mag/
channelizer block at all.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle.c (gcc).  Building the checker is not using it."""
    src = os.path.join(_HERE, "oracle.c")
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _declare(_lib)
    return _lib


_f32p = C.POINTER(C.c_float)


def _declare(L):
    L.orc_num_threads.restype = C.c_int
    L.orc_set_num_threads.argtypes = [C.c_int]
    L.orc_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
    L.orc_multiply_const_ff.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_int64]
    L.orc_multiply_const_cc.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int64]
    L.orc_multiply_const_cc_mt.argtypes = L.orc_multiply_const_cc.argtypes
    L.orc_multiply_const_ss.argtypes = [C.c_void_p, C.c_void_p, C.c_int16, C.c_int64]
    L.orc_multiply_const_ii.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64]
    L.orc_multiply_ff.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    L.orc_multiply_cc.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    L.orc_add_f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    L.orc_complex_to_mag.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    L.orc_complex_to_mag_squared.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    for name in ("orc_resample_ccf_f64", "orc_resample_fff_f64"):
        f = getattr(L, name)
        f.restype = C.c_int64
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    for name in ("orc_fir_ccf_f64", "orc_fir_fff_f64", "orc_fir_ccf_f32", "orc_fir_fff_f32",
                 "orc_fir_ccf_f64_mt", "orc_fir_ccf_f32_mt", "orc_fir_fff_f32_mt"):
        f = getattr(L, name)
        f.restype = C.c_int64
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.orc_window_blackmanharris.argtypes = [C.c_void_p, C.c_int]
    L.orc_fft_f64.restype = C.c_int
    L.orc_fft_f64.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int]
    L.orc_fft_plan_create.restype = C.c_void_p
    L.orc_fft_plan_create.argtypes = [C.c_int]
    L.orc_fft_plan_destroy.argtypes = [C.c_void_p]
    L.orc_fft_f32.restype = C.c_int
    L.orc_fft_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p,
                              C.c_int, C.c_int, C.c_int]
    L.orc_pfb_channelizer_f64.restype = C.c_int64
    L.orc_pfb_channelizer_f64.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int,
                                          C.c_int, C.c_void_p]
    L.orc_pfb_channelizer_f32.restype = C.c_int64
    L.orc_pfb_channelizer_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int,
                                          C.c_int, C.c_void_p, C.c_int]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dtype):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(int(n))


# ------------------------------------------------------------------ C-backed functions
def copy(x: np.ndarray) -> np.ndarray:
    """blocks::copy::work (copy.hpp:33-44)."""
    x = np.ascontiguousarray(x)
    out = np.empty_like(x)
    lib().orc_copy(_p(out), _p(x), x.size, x.itemsize)
    return out


def multiply_const(x: np.ndarray, k) -> np.ndarray:
    """blocks::multiply_const<T>::work (multiply_const.cpp:19-81); dtype selects ff/cc/ss/ii."""
    x = np.ascontiguousarray(x)
    out = np.empty_like(x)
    if x.dtype == np.float32:
        lib().orc_multiply_const_ff(_p(out), _p(x), float(k), x.size)
    elif x.dtype == np.complex64:
        k = complex(k)
        lib().orc_multiply_const_cc(_p(out), _p(x), k.real, k.imag, x.size)
    elif x.dtype == np.int16:
        lib().orc_multiply_const_ss(_p(out), _p(x), int(np.int16(k)), x.size)
    elif x.dtype == np.int32:
        lib().orc_multiply_const_ii(_p(out), _p(x), int(np.int32(k)), x.size)
    else:
        raise TypeError(x.dtype)
    return out


def multiply(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """two-input multiply (float32 / complex64)."""
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b, dtype=a.dtype)
    out = np.empty_like(a)
    (lib().orc_multiply_cc if a.dtype == np.complex64 else lib().orc_multiply_ff)(_p(out), _p(a), _p(b), a.size)
    return out


def add(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """two-input add (float32 / complex64)."""
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b, dtype=a.dtype)
    out = np.empty_like(a)
    lib().orc_add_f(_p(out), _p(a), _p(b), a.size * (2 if a.dtype == np.complex64 else 1))
    return out


def complex_to_mag(x: np.ndarray, squared: bool = False) -> np.ndarray:
    x = _c(x, np.complex64)
    out = np.empty(x.shape, np.float32)
    (lib().orc_complex_to_mag_squared if squared else lib().orc_complex_to_mag)(_p(out), _p(x), x.size)
    return out


def fir(x: np.ndarray, taps, decim: int = 1, hist=None, precise: bool = True, mt: bool = False) -> np.ndarray:
    """fir_filter_ccf (complex64 x) / fir_filter_fff (float32 x); SURVEY.md 8(c).

    precise=True: fp64 accumulate ("truth"); False: fp32 8-lane accumulate (CPU block).
    hist: the ntaps-1 samples preceding x[0] (oldest first) or None for zeros.
    """
    taps = _c(taps, np.float32)
    T = taps.size
    cplx = np.iscomplexobj(x)
    x = _c(x, np.complex64 if cplx else np.float32)
    if hist is not None:
        hist = _c(hist, x.dtype)
        assert hist.size == T - 1
    n_out = x.size // decim
    out = np.empty(n_out, x.dtype)
    name = "orc_fir_%s_%s" % ("ccf" if cplx else "fff", "f64" if precise else "f32")
    if mt:
        name += "_mt"
    r = getattr(lib(), name)(_p(out), _p(x), x.size, _p(taps), T, int(decim), _p(hist))
    assert r == n_out
    return out


def resample(x: np.ndarray, taps, interp: int = 1, decim: int = 1, hist=None) -> np.ndarray:
    """interp_fir_filter / rational_resampler (ccf / fff), fp64 accumulate: y[m] = sum_k h[k] xu[m D - k]
    with xu the interp-fold zero-stuffed x; (len(x) // decim) * interp outputs.
    hist: the ceil(T/interp)-1 samples preceding x[0] (oldest first) or None for zeros."""
    taps = _c(taps, np.float32)
    cplx = np.iscomplexobj(x)
    x = _c(x, np.complex64 if cplx else np.float32)
    if hist is not None:
        hist = _c(hist, x.dtype)
        assert hist.size == (taps.size + interp - 1) // interp - 1
    n_out = (x.size // decim) * interp
    out = np.empty(n_out, x.dtype)
    fn = lib().orc_resample_ccf_f64 if cplx else lib().orc_resample_fff_f64
    r = fn(_p(out), _p(x), x.size, _p(taps), taps.size, int(interp), int(decim), _p(hist))
    assert r == n_out
    return out


def window_blackmanharris(N: int) -> np.ndarray:
    w = np.empty(N, np.float32)
    lib().orc_window_blackmanharris(_p(w), N)
    return w


def fft(x: np.ndarray, N: int, forward: bool = True, window=None, shift: bool = False,
        precise: bool = True, mag: bool = False, mt: bool = False) -> np.ndarray:
    """fft_vcc semantics over x.size//N vectors.  precise=False runs the fp32
    Stockham path (optionally fused |.|, mag=True -> float32 output)."""
    x = _c(x, np.complex64).reshape(-1)
    n_vec = x.size // N
    window = _c(window, np.float32)
    if precise:
        out = np.empty(n_vec * N, np.complex64)
        r = lib().orc_fft_f64(_p(out), _p(x), n_vec, N, int(forward), _p(window), int(shift))
        assert r == 0
        return complex_to_mag(out) if mag else out
    plan = lib().orc_fft_plan_create(N)
    assert plan
    try:
        out = np.empty(n_vec * N, np.float32 if mag else np.complex64)
        r = lib().orc_fft_f32(plan, _p(out), _p(x), n_vec, int(forward), _p(window), int(shift),
                              int(mag), int(mt))
        assert r == 0
    finally:
        lib().orc_fft_plan_destroy(plan)
    return out


def pfb_channelizer(x: np.ndarray, taps, M: int, hist=None, precise: bool = True, mt: bool = False) -> np.ndarray:
    """Critically sampled M-channel analysis bank; returns [n_t, M] complex64."""
    taps = _c(taps, np.float32)
    assert taps.size % M == 0
    P = taps.size // M
    x = _c(x, np.complex64)
    if hist is not None:
        hist = _c(hist, np.complex64)
        assert hist.size == (P - 1) * M
    n_t = x.size // M
    out = np.empty((n_t, M), np.complex64)
    if precise:
        r = lib().orc_pfb_channelizer_f64(_p(out), _p(x), x.size, _p(taps), M, P, _p(hist))
    else:
        r = lib().orc_pfb_channelizer_f32(_p(out), _p(x), x.size, _p(taps), M, P, _p(hist), int(mt))
    assert r == n_t
    return out


# ------------------------------------------------ independent numpy statements (float64)
def np_fir(x, taps, decim=1, hist=None):
    taps = np.asarray(taps, np.float64)
    T = taps.size
    xd = np.asarray(x, np.complex128 if np.iscomplexobj(x) else np.float64)
    pre = np.zeros(T - 1, xd.dtype) if hist is None else np.asarray(hist, xd.dtype)
    full = np.convolve(np.concatenate([pre, xd]), taps)[T - 1:T - 1 + xd.size]
    return full[::decim][: xd.size // decim]


def np_blackmanharris(N):
    n = np.arange(N, dtype=np.float64) / max(N - 1, 1)
    return (0.35875 - 0.48829 * np.cos(2 * np.pi * n) + 0.14128 * np.cos(4 * np.pi * n)
            - 0.01168 * np.cos(6 * np.pi * n))


def np_fft(x, N, forward=True, window=None, shift=False):
    v = np.asarray(x, np.complex64).reshape(-1, N)
    if not forward and shift:
        v = np.fft.ifftshift(v, axes=1)
    if window is not None:
        w = np.asarray(window, np.float32)
        v = (v.real * w + 1j * (v.imag * w)).astype(np.complex64)  # fp32 product like the block
    v = v.astype(np.complex128)
    if forward:
        y = np.fft.fft(v, axis=1)
        if shift:
            y = np.fft.fftshift(y, axes=1)
    else:
        y = np.fft.ifft(v, axis=1) * N
    return y.reshape(-1)


def np_pfb_channelizer(x, taps, M, hist=None):
    taps = np.asarray(taps, np.float64)
    P = taps.size // M
    xd = np.asarray(x, np.complex128)
    pre = np.zeros((P - 1) * M, np.complex128) if hist is None else np.asarray(hist, np.complex128)
    xx = np.concatenate([pre, xd])
    off = (P - 1) * M
    n_t = xd.size // M
    i = np.arange(M)
    u = np.zeros((n_t, M), np.complex128)
    t = np.arange(n_t)
    for r in range(P):
        idx = off + (t[:, None] - r) * M + (M - 1 - i)[None, :]
        u += taps[i + r * M][None, :] * xx[idx]
    return np.fft.ifft(u, axis=1) * M


def rel_rms(a, b) -> float:
    """relative RMS error of a against truth b."""
    a = np.asarray(a).astype(np.complex128 if np.iscomplexobj(a) or np.iscomplexobj(b) else np.float64).ravel()
    b = np.asarray(b).astype(a.dtype).ravel()
    den = np.sqrt(np.mean(np.abs(b) ** 2))
    num = np.sqrt(np.mean(np.abs(a - b) ** 2))
    return float(num / den) if den > 0 else float(num)
