// fft_core.cuh -- register-level radix-4 / radix-16 DFT building blocks shared by the FFT kernels
// (fft.cu) and the overlap-save FIR (fir_ols.cu).
#pragma once
#include "common.cuh"

#ifndef B200_PK_ROT
#define B200_PK_ROT 1 // +-j rotations as one packed add: bit-identical results, +0.5-1 % (measured)
#endif
#ifndef B200_PK_W16
#define B200_PK_W16 0 // packed constant twiddles: within +-2 % either way per kernel (measured), off
#endif

namespace b200 {

constexpr float C8 = 0.92387953251128674f;  // cos(pi/8)
constexpr float S8 = 0.38268343236508977f;  // sin(pi/8)
constexpr float R2 = 0.70710678118654752f;  // sqrt(1/2)

// complex add / subtract as ONE packed instruction each (add.rn.f32x2 / fma.rn.f32x2 -> SASS FADD2 /
// FFMA2): the butterflies are ~75 % of the FFT's arithmetic, and the kernel is issue-bound, so
// halving their issue slots is the cheapest speed-up available.  fma(b, -1, a) == a - b exactly.
__device__ __forceinline__ float2 operator+(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 operator-(float2 a, float2 b)
{
    return __ffma2_rn(b, make_float2(-1.f, -1.f), a);
}

// in: x0..x3 in (a,b,c,d); out: X0..X3 in (a,b,c,d)
template <bool FWD>
__device__ __forceinline__ void dft4(float2& a, float2& b, float2& c, float2& d)
{
    float2 t0 = a + c, t1 = a - c, t2 = b + d, t3 = b - d;
    a = t0 + t2;
    c = t0 - t2;
#if B200_PK_ROT
    // t1 -/+ j t3: one packed add each, the rotated operand is a swap + negate modifier (FADD2 ... LO_HI.NP)
    const float2 mj = make_float2(t3.y, -t3.x), pj = make_float2(-t3.y, t3.x);
    b = __fadd2_rn(t1, FWD ? mj : pj);
    d = __fadd2_rn(t1, FWD ? pj : mj);
#else
    if (FWD) {
        b = make_float2(t1.x + t3.y, t1.y - t3.x);
        d = make_float2(t1.x - t3.y, t1.y + t3.x);
    } else {
        b = make_float2(t1.x - t3.y, t1.y + t3.x);
        d = make_float2(t1.x + t3.y, t1.y - t3.x);
    }
#endif
}

// multiply by W16^m (forward: e^{-j 2 pi m/16}; reverse: conjugate), m compile-time
template <bool FWD, int M>
__device__ __forceinline__ float2 mul_w16(float2 z)
{
    constexpr float cr[10] = { 1.f, C8, R2, S8, 0.f, -S8, -R2, -C8, -1.f, -C8 };
    constexpr float si[10] = { 0.f, S8, R2, C8, 1.f, C8, R2, S8, 0.f, -S8 };
    constexpr float wr = cr[M];
    constexpr float wi = FWD ? -si[M] : si[M];
    if (M == 0)
        return z;
    if (M == 4)
        return FWD ? make_float2(z.y, -z.x) : make_float2(-z.y, z.x);
#if B200_PK_W16
    // z * (wr + j wi) in two packed instructions (see cmul in common.cuh)
    const float2 r = __fmul2_rn(z, make_float2(wr, wr));
    return __ffma2_rn(make_float2(-z.y, z.x), make_float2(wi, wi), r);
#else
    return make_float2(fmaf(-z.y, wi, z.x * wr), fmaf(z.x, wi, z.y * wr));
#endif
}

// 16-point DFT in registers.  Input natural order v[n]; output X[k] lands in v[pos16(k)].
__host__ __device__ constexpr int pos16(int k) { return 4 * (k & 3) + (k >> 2); }

template <bool FWD>
__device__ __forceinline__ void dft16(float2 (&v)[16])
{
#pragma unroll
    for (int b = 0; b < 4; b++)
        dft4<FWD>(v[b], v[4 + b], v[8 + b], v[12 + b]);
    // v[4c+b] = y_b[c]; twiddle by W16^{b c}
    v[5] = mul_w16<FWD, 1>(v[5]);
    v[6] = mul_w16<FWD, 2>(v[6]);
    v[7] = mul_w16<FWD, 3>(v[7]);
    v[9] = mul_w16<FWD, 2>(v[9]);
    v[10] = mul_w16<FWD, 4>(v[10]);
    v[11] = mul_w16<FWD, 6>(v[11]);
    v[13] = mul_w16<FWD, 3>(v[13]);
    v[14] = mul_w16<FWD, 6>(v[14]);
    v[15] = mul_w16<FWD, 9>(v[15]);
#pragma unroll
    for (int c = 0; c < 4; c++)
        dft4<FWD>(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}

__device__ __forceinline__ float sqrt_approx(float x)
{
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

constexpr int F4K_STRIDE = 257;


} // namespace b200
