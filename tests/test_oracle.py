"""CPU: pins oracle.c against the golden vectors and an independent numpy statement."""
import numpy as np
import pytest

import oracle as o
from conftest import cplx


def test_copy_reference_ramp(golden):
    # schedulers/mt/test/cuda/qa_scheduler_mt_cuda_copy.cpp:24-50: EXPECT_EQ(snk->data(), input)
    x = golden["ramp_qa_cuda_copy"]
    assert np.array_equal(o.copy(x).view(np.uint8), x.view(np.uint8))
    for dt in (np.uint8, np.int16, np.float64):
        a = np.arange(1001).astype(dt)
        assert np.array_equal(o.copy(a), a)


def test_multiply_const_reference_k1(golden):
    # qa_scheduler_mt.cpp:86-88,128-132: output == k * input exactly for k = 1
    x = golden["ramp_qa_scheduler_mt"]
    assert np.array_equal(o.multiply_const(x, 1.0 + 0j), x)
    assert np.array_equal(o.multiply_const(x.real.copy(), 1.0), x.real)


def test_multiply_const_values(golden):
    x, k = golden["mulc_x"], golden["mulc_k"][0]
    assert o.rel_rms(o.multiply_const(x, k), golden["mulc_y64"]) < 1e-6
    xi = np.arange(-300, 300, dtype=np.int16) * 100
    assert np.array_equal(o.multiply_const(xi, 7), (xi.astype(np.int32) * 7).astype(np.int16))
    xl = (np.arange(-300, 300, dtype=np.int64) * 12345677).astype(np.int32)
    assert np.array_equal(o.multiply_const(xl, 1000003),
                          (xl.astype(np.int64) * 1000003).astype(np.int32))


def test_two_input_blocks(golden):
    x = golden["mulc_x"]
    y = x[::-1].copy()
    assert o.rel_rms(o.multiply(x, y), x.astype(np.complex128) * y.astype(np.complex128)) < 1e-6
    assert np.array_equal(o.add(x, y), x + y)
    assert np.array_equal(o.multiply(x.real.copy(), y.imag.copy()), x.real * y.imag)
    # multiply by a constant vector == multiply_const
    k = np.full(x.size, 0.5 - 0.25j, np.complex64)
    assert np.array_equal(o.multiply(x, k), o.multiply_const(x, 0.5 - 0.25j))


def test_complex_to_mag(golden):
    x = golden["mulc_x"]
    assert o.rel_rms(o.complex_to_mag(x), golden["mag_y64"]) < 1e-6
    assert np.allclose(o.complex_to_mag(x, squared=True), np.abs(x.astype(np.complex128)) ** 2, rtol=1e-6)


@pytest.mark.parametrize("name", ["fir_a", "fir_b", "fir_c"])
def test_fir_golden(golden, name):
    taps, x, D = golden[name + "_taps"], golden[name + "_x"], int(golden[name + "_D"][0])
    T = taps.size
    for precise, tol in ((True, 2e-7), (False, 1e-5)):
        assert o.rel_rms(o.fir(x, taps, D, precise=precise), golden[name + "_y64"]) < tol
        y = o.fir(x[T - 1:], taps, D, hist=x[: T - 1], precise=precise)
        assert o.rel_rms(y, golden[name + "_y64_hist"]) < tol


def test_fir_impulse_and_phase(golden):
    taps = golden["fir_imp_taps"]
    imp = np.zeros(128, np.complex64)
    imp[0] = 1
    y = o.fir(imp, taps, 1)
    assert np.array_equal(y[:48].real, taps)  # impulse response == taps, bit exact
    assert not y[48:].any()
    # decimation phase 0: output m is aligned to input m*D
    y4 = o.fir(imp, taps, 4)
    assert np.array_equal(y4.real[:12], taps[::4])
    # empty and shorter-than-D inputs
    assert o.fir(np.zeros(0, np.complex64), taps, 4).size == 0
    assert o.fir(np.zeros(3, np.complex64), taps, 4).size == 0


def test_fir_chunking_matches_oneshot():
    rng = np.random.default_rng(3)
    x = cplx(rng, 5000)
    taps = (rng.uniform(-1, 1, 31) / 31).astype(np.float32)
    ref = o.fir(x, taps, 1, precise=False)
    for chunk in (1, 30, 31, 97, 4096):
        hist = np.zeros(30, np.complex64)
        outs = []
        for s in range(0, x.size, chunk):
            blk = x[s:s + chunk]
            outs.append(o.fir(blk, taps, 1, hist=hist, precise=False))
            hist = np.concatenate([hist, blk])[-30:]
        assert np.array_equal(np.concatenate(outs), ref)


def test_window(golden):
    assert np.abs(o.window_blackmanharris(4096) - golden["bh4096"]).max() < 6e-8
    assert np.abs(o.np_blackmanharris(4096) - golden["bh4096"]).max() < 1e-15


@pytest.mark.parametrize("key,fwd,shift", [("fft_fwd", True, False), ("fft_fwd_shift", True, True),
                                           ("fft_rev", False, False), ("fft_rev_shift", False, True)])
def test_fft_golden(golden, key, fwd, shift):
    x, w = golden["fft_x"], golden["bh4096"].astype(np.float32)
    assert o.rel_rms(o.fft(x, 4096, fwd, w, shift), golden[key]) < 1e-7
    assert o.rel_rms(o.fft(x, 4096, fwd, w, shift, precise=False), golden[key]) < 1e-5
    assert o.rel_rms(o.np_fft(x, 4096, fwd, w, shift), golden[key]) < 1e-12


@pytest.mark.parametrize("N", [64, 1024])
def test_fft_small(golden, N):
    x = golden[f"fft{N}_x"]
    assert o.rel_rms(o.fft(x, N), golden[f"fft{N}_fwd"]) < 1e-7
    assert o.rel_rms(o.fft(x, N, precise=False), golden[f"fft{N}_fwd"]) < 1e-5
    # round trip: reverse(forward(x)) == N x
    y = o.fft(o.fft(x, N), N, forward=False)
    assert o.rel_rms(y / N, x) < 1e-6


def test_fft_mag_fused_equals_unfused(golden):
    x, w = golden["fft_x"], golden["bh4096"].astype(np.float32)
    a = o.fft(x, 4096, True, w, False, precise=False, mag=True)
    assert o.rel_rms(a, np.abs(golden["fft_fwd"])) < 1e-5


@pytest.mark.parametrize("name", ["pfb_a", "pfb_b"])
def test_pfb_golden(golden, name):
    taps, x, M = golden[name + "_taps"], golden[name + "_x"], int(golden[name + "_M"][0])
    assert o.rel_rms(o.pfb_channelizer(x, taps, M), golden[name + "_y64"]) < 2e-7
    assert o.rel_rms(o.pfb_channelizer(x, taps, M, precise=False), golden[name + "_y64"]) < 1e-5
    assert o.rel_rms(o.np_pfb_channelizer(x, taps, M), golden[name + "_y64"]) < 1e-12


def test_pfb_tone_lands_in_its_channel():
    import scipy.signal as sig
    M, P, c0 = 64, 16, 5
    taps = sig.firwin(M * P, 1.0 / M).astype(np.float32)
    n = np.arange(M * 200)
    x = np.exp(2j * np.pi * c0 / M * n).astype(np.complex64)
    y = o.pfb_channelizer(x, taps, M)
    p = (np.abs(y[P:]) ** 2).mean(axis=0)
    assert int(np.argmax(p)) == c0
    assert p[c0] > 1e3 * np.delete(p, c0).max()


def test_pfb_history_chunking(golden):
    taps, x = golden["pfb_b_taps"], golden["pfb_b_x"]
    M, P = 8, 5
    ref = o.pfb_channelizer(x, taps, M)
    hist = np.zeros((P - 1) * M, np.complex64)
    outs = []
    for s in range(0, x.size, 5 * M):
        blk = x[s:s + 5 * M]
        outs.append(o.pfb_channelizer(blk, taps, M, hist=hist))
        hist = np.concatenate([hist, blk])[-(P - 1) * M:]
    assert np.array_equal(np.concatenate(outs), ref)


def test_mt_variants_agree():
    rng = np.random.default_rng(5)
    x = cplx(rng, 1 << 14)
    taps = (rng.uniform(-1, 1, 64) / 64).astype(np.float32)
    assert np.array_equal(o.fir(x, taps, 1, precise=False, mt=True), o.fir(x, taps, 1, precise=False))
    w = o.window_blackmanharris(4096)
    assert np.array_equal(o.fft(x, 4096, True, w, precise=False, mt=True),
                          o.fft(x, 4096, True, w, precise=False))


@pytest.mark.parametrize("T,L,D,cplxin", [(48, 4, 1, True), (49, 3, 2, True), (160, 8, 5, False),
                                          (7, 1, 3, True), (33, 5, 5, False), (3, 7, 2, True)])
def test_resampler_matches_upfirdn(T, L, D, cplxin):
    """interp_fir_filter / rational_resampler definition against scipy.signal.upfirdn (fp64), and
    streaming with history equals one-shot."""
    import scipy.signal as sig
    rng = np.random.default_rng(T * 31 + L * 7 + D)
    n = 3000 + 17
    x = cplx(rng, n) if cplxin else rng.uniform(-1, 1, n).astype(np.float32)
    taps = rng.uniform(-1, 1, T).astype(np.float32)
    y = o.resample(x, taps, L, D)
    ref = sig.upfirdn(taps.astype(np.float64), x.astype(np.complex128 if cplxin else np.float64), up=L, down=D)
    assert y.size == (n // D) * L
    assert o.rel_rms(y, ref[: y.size]) < 2e-7
    # an impulse interpolated by L reproduces the taps exactly (phase 0 first)
    imp = np.zeros(64, x.dtype)
    imp[0] = 1
    yi = o.resample(imp, taps, L, 1)
    assert np.array_equal(yi[:T].real if cplxin else yi[:T], taps) and not yi[T:].any()
    # streaming: cut at a multiple of D, carry ceil(T/L)-1 samples of history
    cut = (n // 2) // D * D
    nh = (T + L - 1) // L - 1
    y1 = o.resample(x[:cut], taps, L, D)
    y2 = o.resample(x[cut:], taps, L, D, hist=x[cut - nh:cut] if nh else None)
    assert np.array_equal(np.concatenate([y1, y2]), y)
