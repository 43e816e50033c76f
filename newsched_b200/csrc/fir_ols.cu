// fir_ols.cu -- long-tap complex FIR (ccf) by overlap-save fast convolution, FFT size 4096.
//
// For T taps the direct form costs 4T flop per sample and is FP32-pipe bound (fir.cu); from
// roughly 100 taps per output on, one forward FFT + spectrum multiply + inverse FFT per block of
// V = 4096 - (T-1) valid samples is cheaper (~2 x 62 flop per transformed sample).  One CTA does
// the whole block in shared memory / registers, reusing the 3 x radix-16 machinery of fft.cu:
//   TMA bulk load of the 4096-sample segment (overlapping the previous block's compute)
//   -> forward passes 1-3 -> X[k] * H[k] in registers -> inverse passes 1-3 -> store the V
//   valid samples (every D-th one when decimating).
// Pass 3 of the forward transform leaves X[tid + 256 k2] in the registers of thread tid, which
// is exactly the input arrangement of pass 1, so the inverse transform starts from registers
// with no extra exchange.  H = FFT(taps) / 4096 (and the fused multiply_const k) is computed
// once at create time in double precision.  Samples stay fp32; the result differs from the
// direct form only by FFT rounding (~3e-7 relative RMS, tests bound it by 1e-5).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include <cuda.h>

#include "fft_core.cuh"
#include "fir_ols.cuh"

// Pass-2 twiddles W256^{n0 k1} depend on the thread only (n0 = tid & 15).  The one-phase kernel keeps
// them in registers for the lifetime of the CTA instead of re-reading the shared table for every
// transform (measured +5..6.5 %: T=128 181 -> 193 GS/s).  The polyphase / two-phase kernels are at
// the 128-register limit already: the same change spills and costs 10 % / 1.5 % there, so they read
// the table (T2R = false).
namespace b200 {

constexpr int OLS_N = 4096;
// + the spectrum table H (32 KiB), resident in shared memory for the lifetime of the CTA (B200_OLS_HSMEM): the
// multiply step reads it with 16 conflict-free LDS.64 per thread instead of 16 L2-latency loads per block
#ifndef B200_OLS_HSMEM
#define B200_OLS_HSMEM 1
#endif
constexpr size_t OLS_SMEM = OLS_N * 8 + 16 * F4K_STRIDE * 8 + 256 * 8 + 16 + (B200_OLS_HSMEM ? OLS_N * 8 : 0);

struct ols_geom {
    int Ov;  // samples of overlap discarded at the head of every block (>= T-1, even)
    int V;   // new samples per block (multiple of 2*D)
    int D;
    int Tm1;
    int tma_ok;
    int shift;      // this partition filters x delayed by `shift` samples (partitioned convolution)
    int accumulate; // add to y instead of overwriting it (partitions after the first)
    long long n_in, n_out, n_blocks;
};

__device__ __forceinline__ float2 ols_fetch(const float2* __restrict__ x, const float2* __restrict__ hist,
                                            int Tm1, long long g, long long n_in)
{
    if (g >= 0)
        return g < n_in ? __ldg(x + g) : make_float2(0.f, 0.f);
    if (hist && g >= -(long long)Tm1)
        return __ldg(hist + (Tm1 + g));
    return make_float2(0.f, 0.f);
}

__device__ __forceinline__ float ols_fetch_real(const float* __restrict__ x, const float* __restrict__ hist,
                                                int Tm1, long long g, long long n_in)
{
    if (g >= 0)
        return g < n_in ? __ldg(x + g) : 0.f;
    if (hist && g >= -(long long)Tm1)
        return __ldg(hist + (Tm1 + g));
    return 0.f;
}

// three radix-16 passes over registers v[] (natural order in, X[tid + 256 j] in v[pos16(j)] out)
template <bool FWD, bool T2R = false>
__device__ __forceinline__ void fft4096_passes(float2 (&v)[16], float2* sA, const float2* sT2,
                                               const float2 (&t1)[16], int tid, const float2* t2r = nullptr)
{
    dft16<FWD>(v);
#pragma unroll
    for (int k0 = 0; k0 < 16; k0++)
        sA[k0 * F4K_STRIDE + tid] = FWD ? cmul(v[pos16(k0)], t1[k0]) : cmul_conj(v[pos16(k0)], t1[k0]);
    __syncthreads();
    {
        const int k0 = tid >> 4, n0 = tid & 15;
        float2* row = sA + k0 * F4K_STRIDE + n0;
#pragma unroll
        for (int i = 0; i < 16; i++)
            v[i] = row[i * 16];
        dft16<FWD>(v);
        row[0] = v[pos16(0)];
#pragma unroll
        for (int k1 = 1; k1 < 16; k1++)
            {
                const float2 w2 = T2R ? t2r[k1] : sT2[k1 * 16 + n0];
                row[k1 * 16] = FWD ? cmul(v[pos16(k1)], w2) : cmul_conj(v[pos16(k1)], w2);
            }
    }
    __syncthreads();
    {
        const int k0 = tid & 15, k1 = tid >> 4;
        const float2* row = sA + k0 * F4K_STRIDE + k1 * 16;
#pragma unroll
        for (int i = 0; i < 16; i++)
            v[i] = row[i];
        dft16<FWD>(v);
    }
}

// REAL = true: float stream (fff).  Real taps commute with re/im, so TWO consecutive real blocks ride
// through one complex transform as (block A) + j (block B): twice the sample rate of the complex
// path for the same arithmetic.
template <bool REAL>
__global__ void __launch_bounds__(256, 2)
    fir_ols4096_kernel(const float2* __restrict__ x, const float2* __restrict__ hist, float2* __restrict__ y,
                       const float2* __restrict__ Htab, const float2* __restrict__ tw1,
                       const float2* __restrict__ tw2, ols_geom g)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2* sIn = reinterpret_cast<float2*>(smem_raw);
    float2* sA = sIn + OLS_N;
    float2* sT2 = sA + 16 * F4K_STRIDE;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sT2 + 256);
#if B200_OLS_HSMEM
    float2* sH = reinterpret_cast<float2*>(bar + 2);
#endif
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    float2 t1[16];
#pragma unroll
    for (int i = 0; i < 16; i++)
        t1[i] = __ldg(tw1 + i * 256 + tid);
    sT2[tid] = __ldg(tw2 + tid);
#if B200_OLS_HSMEM
#pragma unroll
    for (int i = 0; i < 16; i++)
        sH[i * 256 + tid] = __ldg(Htab + i * 256 + tid); // each thread reads back exactly what it wrote
#endif
    __syncthreads();
    float2 t2r[16]; // this thread's pass-2 twiddles, register resident (see top of file)
#pragma unroll
    for (int k1 = 1; k1 < 16; k1++)
        t2r[k1] = sT2[k1 * 16 + (tid & 15)];

    // a "block" is one complex transform: one segment of the complex stream, or two consecutive
    // segments (A, B = A + V) of the real stream
    constexpr int SEGS = REAL ? 2 : 1;
    auto seg_start = [&](long long b) { return b * SEGS * g.V - g.Ov - g.shift; };
    auto tma_block = [&](long long b) {
        const long long s = seg_start(b);
        return g.tma_ok && s >= 0 && s + (SEGS - 1) * g.V + OLS_N <= g.n_in;
    };
    auto issue = [&](long long b) { // one elected thread
        mbar_arrive_expect_tx(bar, OLS_N * 8);
        if (REAL) {
            const float* xr = reinterpret_cast<const float*>(x);
            float* sf = reinterpret_cast<float*>(sIn);
            bulk_copy_g2s(sf, xr + seg_start(b), OLS_N * 4, bar);
            bulk_copy_g2s(sf + OLS_N, xr + seg_start(b) + g.V, OLS_N * 4, bar);
        } else {
            bulk_copy_g2s(sIn, x + seg_start(b), OLS_N * 8, bar);
        }
    };
    // programmatic dependent launch (B200_LAUNCH_PDL): the prologue above ran while the previous kernel in the
    // stream drained; nothing below may start before that kernel has completed and flushed
    pdl_wait();
    pdl_launch_dependents();
    long long blk = blockIdx.x;
    if (tid == 0 && blk < g.n_blocks && tma_block(blk))
        issue(blk);
    uint32_t phase = 0;
    for (; blk < g.n_blocks; blk += gridDim.x) {
        const long long s = seg_start(blk);
        float2 v[16];
        const bool via_tma = tma_block(blk);
        if (via_tma) {
            mbar_wait(bar, phase);
            phase ^= 1;
            if (REAL) {
                const float* sf = reinterpret_cast<const float*>(sIn);
#pragma unroll
                for (int i = 0; i < 16; i++)
                    v[i] = make_float2(sf[i * 256 + tid], sf[OLS_N + i * 256 + tid]);
            } else {
#pragma unroll
                for (int i = 0; i < 16; i++)
                    v[i] = sIn[i * 256 + tid];
            }
        } else if (REAL) {
            const float* xr = reinterpret_cast<const float*>(x);
            const float* hr = reinterpret_cast<const float*>(hist);
#pragma unroll
            for (int i = 0; i < 16; i++)
                v[i] = make_float2(ols_fetch_real(xr, hr, g.Tm1, s + i * 256 + tid, g.n_in),
                                   ols_fetch_real(xr, hr, g.Tm1, s + g.V + i * 256 + tid, g.n_in));
        } else {
#pragma unroll
            for (int i = 0; i < 16; i++)
                v[i] = ols_fetch(x, hist, g.Tm1, s + i * 256 + tid, g.n_in);
        }
        // ---- forward transform; the first barrier inside also retires every read of sIn
        dft16<true>(v);
#pragma unroll
        for (int k0 = 0; k0 < 16; k0++)
            sA[k0 * F4K_STRIDE + tid] = cmul(v[pos16(k0)], t1[k0]);
        __syncthreads();
        {
            const long long nxt = blk + gridDim.x;
            if (tid == 0 && nxt < g.n_blocks && tma_block(nxt))
                issue(nxt);
        }
        {
            const int k0 = tid >> 4, n0 = tid & 15;
            float2* row = sA + k0 * F4K_STRIDE + n0;
#pragma unroll
            for (int i = 0; i < 16; i++)
                v[i] = row[i * 16];
            dft16<true>(v);
            row[0] = v[pos16(0)];
#pragma unroll
            for (int k1 = 1; k1 < 16; k1++)
                row[k1 * 16] = cmul(v[pos16(k1)], t2r[k1]);
        }
        __syncthreads();
        float2 u[16];
        {
            const int k0 = tid & 15, k1 = tid >> 4;
            const float2* row = sA + k0 * F4K_STRIDE + k1 * 16;
#pragma unroll
            for (int i = 0; i < 16; i++)
                v[i] = row[i];
            dft16<true>(v);
            // ---- spectrum multiply: thread holds X[tid + 256 k2] in v[pos16(k2)]
#pragma unroll
            for (int k2 = 0; k2 < 16; k2++)
#if B200_OLS_HSMEM
                u[k2] = cmul(v[pos16(k2)], sH[k2 * 256 + tid]);
#else
                u[k2] = cmul(v[pos16(k2)], __ldg(Htab + k2 * 256 + tid));
#endif
        }
        __syncthreads(); // pass-3 reads of sA done before the inverse transform overwrites it
        // ---- inverse transform straight from registers (u[k2] plays x[n2*256 + tid])
        fft4096_passes<false, true>(u, sA, sT2, t1, tid, t2r);
        // ---- u[pos16(j)] = y_circ[tid + 256 j]; keep n >= Ov, every D-th input-rate sample
        const long long out_base = blk * SEGS * g.V - g.Ov; // input-rate index of circular sample 0
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const int n = tid + 256 * j;
            if (n >= g.Ov && n < g.Ov + g.V) {
#pragma unroll
                for (int sg = 0; sg < SEGS; sg++) {
                    const long long gi = out_base + n + (long long)sg * g.V;
                    long long m = gi;
                    bool ok = true;
                    if (g.D != 1) {
                        ok = (gi % g.D == 0);
                        m = gi / g.D;
                    }
                    if (ok && m < g.n_out) {
                        if (REAL) {
                            float* yr = reinterpret_cast<float*>(y);
                            float r = sg == 0 ? u[pos16(j)].x : u[pos16(j)].y;
                            if (g.accumulate)
                                r += yr[m];
                            __stcs(yr + m, r);
                        } else {
                            float2 r = u[pos16(j)];
                            if (g.accumulate) {
                                const float2 prev = y[m];
                                r.x += prev.x;
                                r.y += prev.y;
                            }
                            __stcs(y + m, r);
                        }
                    }
                }
            }
        }
        __syncthreads(); // pass-3 reads done before the next block's pass-1 writes
    }
}


// ---------------------------------------------------------------------------------------------
// Decimating complex filters, D <= 16 (TMA-staged for even D): polyphase overlap-save.
//   y[m] = sum_p (h_p * x_p)[m],  h_p[q] = h[qD + p],  x_p[j] = x[jD - p]
// so one block = D forward transforms (one per phase stream, each at the OUTPUT rate), the sum
// sum_p X_p G_p accumulated in registers, and ONE inverse transform: (D + 1) transforms per
// (4096 - ceil(T/D)) * D input samples instead of 2 per 4096 - (T - 1) with 1/D of the results
// kept.  T = 1024, D = 4: 1.33 transformed samples per input sample instead of 2.67.
//
// Staging: the input is viewed as a 2-D tensor of rows of D samples (8-byte elements, row stride
// D*8 bytes).  One TMA box {2, 256} pulls the two phases that share a 16-byte granule for 256
// consecutive rows; 16 boxes fill a 64 KiB pair plane sP[n][2].  Every 32-byte sector is
// requested exactly twice per block (once per pair), all requests 16-byte aligned.  A phase whose
// sample falls into the next row is read one output period late and its spectrum table G_p is
// advanced by one sample to compensate (tables for both cases are built at create time).
constexpr int OLSD_MAXD = 16;
constexpr size_t OLSD_SMEM = OLS_N * 16 + 16 * F4K_STRIDE * 8 + 256 * 8 + 16;

struct olsd_geom {
    int D, Ov, V, Tm1;
    int tma_ok;
    int rmin;  // row of circular sample 0 relative to (b V - Ov)
    int gbase; // sample index of (row r, slot s) = r D + s + gbase
    int off;   // samples of the 16-byte granule that precede x[0] (0 or 1): fixes the slot -> phase map
    long long n_in, n_out, n_blocks;
};

__device__ __forceinline__ void tma_load_box2d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1,
                                               uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
                 "[%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(smem_dst)),
                 "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}

__global__ void __launch_bounds__(256, 2)
    fir_olsd_kernel(const float2* __restrict__ x, const float2* __restrict__ hist, float2* __restrict__ y,
                    const float2* __restrict__ Gtab, const float2* __restrict__ tw1,
                    const float2* __restrict__ tw2, const __grid_constant__ CUtensorMap tmap, olsd_geom g)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2* sP = reinterpret_cast<float2*>(smem_raw); // [4096][2]
    float2* sA = sP + 2 * OLS_N;
    float2* sT2 = sA + 16 * F4K_STRIDE;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sT2 + 256);
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    float2 t1[16];
#pragma unroll
    for (int i = 0; i < 16; i++)
        t1[i] = __ldg(tw1 + i * 256 + tid);
    sT2[tid] = __ldg(tw2 + tid);
    __syncthreads();

    const int n_pairs = g.D >> 1; // odd D: no pair planes (row stride not a multiple of 16 bytes), loads go direct
    auto row0 = [&](long long b) { return b * g.V - g.Ov + g.rmin; };
    auto tma_block = [&](long long b) {
        const long long r = row0(b);
        return g.tma_ok && r * g.D + g.gbase >= 0 && (r + OLS_N - 1) * g.D + g.D - 1 + g.gbase < g.n_in;
    };
    auto issue = [&](long long b, int pair) { // one elected thread: 16 boxes of {2 samples, 256 rows}
        mbar_arrive_expect_tx(bar, OLS_N * 16);
        const long long r = row0(b);
#pragma unroll 1
        for (int c = 0; c < 16; c++)
            tma_load_box2d(sP + c * 512, &tmap, 2 * pair, (int)(r + 256 * c), bar);
    };
    // programmatic dependent launch (B200_LAUNCH_PDL): the prologue above ran while the previous kernel in the
    // stream drained; nothing below may start before that kernel has completed and flushed
    pdl_wait();
    pdl_launch_dependents();
    long long blk = blockIdx.x;
    if (tid == 0 && blk < g.n_blocks && tma_block(blk))
        issue(blk, 0);
    uint32_t phase = 0;
    for (; blk < g.n_blocks; blk += gridDim.x) {
        const bool via_tma = tma_block(blk);
        const long long r0 = row0(blk);
        float2 acc[16];
#pragma unroll
        for (int i = 0; i < 16; i++)
            acc[i] = make_float2(0.f, 0.f);
#pragma unroll 1
        for (int slot = 0; slot < g.D; slot++) {
            const int e = slot & 1;
            float2 v[16];
            if (via_tma) {
                if (e == 0) {
                    mbar_wait(bar, phase);
                    phase ^= 1;
                }
#pragma unroll
                for (int i = 0; i < 16; i++)
                    v[i] = sP[(i * 256 + tid) * 2 + e];
            } else {
#pragma unroll
                for (int i = 0; i < 16; i++)
                    v[i] = ols_fetch(x, hist, g.Tm1, (r0 + i * 256 + tid) * g.D + slot + g.gbase, g.n_in);
            }
            dft16<true>(v);
#pragma unroll
            for (int k0 = 0; k0 < 16; k0++)
                sA[k0 * F4K_STRIDE + tid] = cmul(v[pos16(k0)], t1[k0]);
            __syncthreads(); // also retires this slot's reads of sP
            if (e == 1 && tid == 0) {
                // refill the pair plane: next pair of this block, or pair 0 of this CTA's next block
                if ((slot >> 1) + 1 < n_pairs) {
                    if (via_tma)
                        issue(blk, (slot >> 1) + 1);
                } else {
                    const long long nxt = blk + gridDim.x;
                    if (nxt < g.n_blocks && tma_block(nxt))
                        issue(nxt, 0);
                }
            }
            {
                const int k0 = tid >> 4, n0 = tid & 15;
                float2* row = sA + k0 * F4K_STRIDE + n0;
#pragma unroll
                for (int i = 0; i < 16; i++)
                    v[i] = row[i * 16];
                dft16<true>(v);
                row[0] = v[pos16(0)];
#pragma unroll
                for (int k1 = 1; k1 < 16; k1++)
                    row[k1 * 16] = cmul(v[pos16(k1)], sT2[k1 * 16 + n0]);
            }
            __syncthreads();
            {
                const int k0 = tid & 15, k1 = tid >> 4;
                const float2* row = sA + k0 * F4K_STRIDE + k1 * 16;
#pragma unroll
                for (int i = 0; i < 16; i++)
                    v[i] = row[i];
                dft16<true>(v);
                // slot -> (phase, late): x_ph[j] sits at row j + floor((off - ph) / D), slot (off - ph) mod D
                const int ph = slot <= g.off ? g.off - slot : g.off - slot + g.D;
                const int late = (slot <= g.off ? 0 : -1) - g.rmin;
                const float2* G = Gtab + (size_t)(ph * 2 + late) * OLS_N + tid;
#pragma unroll
                for (int k2 = 0; k2 < 16; k2++) {
                    const float2 w = __ldg(G + k2 * 256), a = v[pos16(k2)];
                    acc[k2] = cmac(a, w, acc[k2]);
                }
            }
            __syncthreads(); // pass-3 reads of sA done before the next transform's pass-1 writes
        }
        // ---- one inverse transform at the output rate, straight from the accumulators
        fft4096_passes<false>(acc, sA, sT2, t1, tid);
        const long long out_base = blk * g.V - g.Ov;
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const int n = tid + 256 * j;
            const long long m = out_base + n;
            if (n >= g.Ov && n < g.Ov + g.V && m < g.n_out)
                __stcs(y + m, acc[pos16(j)]);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Full-rate complex filters with many taps: two-phase (2 x 2 polyphase) overlap-save.
//   y[2m + a] = sum_{p,q} h[2q + p] x[2(m - q) + a - p]
// i.e. with the even / odd streams x_0, x_1 and tap phases h_0, h_1 (all at HALF rate)
//   Y_0 = H_0 X_0 + z^-1 H_1 X_1,   Y_1 = H_1 X_0 + H_0 X_1.
// One block = 2 forward + 2 inverse 4096-point transforms for 2 (4095 - ceil(T/2)) samples: the
// overlap costs T/2 instead of T - 1 of every 4096 points (T = 4096: 4 transformed samples per
// output instead of 8 with two 2048-tap partitions; T = 2048: 2.7 instead of 4).
// The block's input is one CONTIGUOUS run of 8192 samples, which already is the pair plane
// sP[n][2] = (x_0[n], x_1[n]): a single 64 KiB bulk copy, read back with one LDS.128 per point.
// If x[0] sits 8 bytes past a 16-byte boundary the roles of the streams swap (x'_c = granule-
// aligned phases) and one term needs a one-sample ADVANCE instead of the delay; the host picks the
// four coefficient tables C[a][c] per call from {H_0, H_1, z^-1 H_1, z^+1 H_0}.
struct ols2_geom {
    int Ov, V, Tm1, tma_ok, off;
    int c00, c01, c10, c11; // table index of C[a][c]
    long long n_in, n_out, n_blocks;
};

__global__ void __launch_bounds__(256, 2)
    fir_ols2_kernel(const float2* __restrict__ x, const float2* __restrict__ hist, float2* __restrict__ y,
                    const float2* __restrict__ Htab, const float2* __restrict__ tw1,
                    const float2* __restrict__ tw2, ols2_geom g)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2* sP = reinterpret_cast<float2*>(smem_raw); // [4096][2]
    float2* sA = sP + 2 * OLS_N;
    float2* sT2 = sA + 16 * F4K_STRIDE;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sT2 + 256);
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    float2 t1[16];
#pragma unroll
    for (int i = 0; i < 16; i++)
        t1[i] = __ldg(tw1 + i * 256 + tid);
    sT2[tid] = __ldg(tw2 + tid);
    __syncthreads();

    // block b: half-rate times j0 + n, j0 = b V - Ov; aligned-stream element 2 (j0 + n) + c is
    // sample x[2 (j0 + n) + c - off]
    auto first = [&](long long b) { return 2 * (b * g.V - g.Ov) - g.off; };
    auto tma_block = [&](long long b) {
        const long long s = first(b);
        return g.tma_ok && s >= 0 && s + 2 * OLS_N <= g.n_in;
    };
    auto issue = [&](long long b) {
        mbar_arrive_expect_tx(bar, OLS_N * 16);
        bulk_copy_g2s(sP, x + first(b), OLS_N * 8, bar);
        bulk_copy_g2s(sP + OLS_N, x + first(b) + OLS_N, OLS_N * 8, bar);
    };
    // programmatic dependent launch (B200_LAUNCH_PDL): the prologue above ran while the previous kernel in the
    // stream drained; nothing below may start before that kernel has completed and flushed
    pdl_wait();
    pdl_launch_dependents();
    long long blk = blockIdx.x;
    if (tid == 0 && blk < g.n_blocks && tma_block(blk))
        issue(blk);
    uint32_t phase = 0;
    const bool out16 = ((uintptr_t)y & 15) == 0;
    for (; blk < g.n_blocks; blk += gridDim.x) {
        const long long s0 = first(blk);
        float2 v0[16], v1[16];
        if (tma_block(blk)) {
            mbar_wait(bar, phase);
            phase ^= 1;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const float4 q = reinterpret_cast<const float4*>(sP)[i * 256 + tid];
                v0[i] = make_float2(q.x, q.y);
                v1[i] = make_float2(q.z, q.w);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; i++) {
                v0[i] = ols_fetch(x, hist, g.Tm1, s0 + 2 * (i * 256 + tid), g.n_in);
                v1[i] = ols_fetch(x, hist, g.Tm1, s0 + 2 * (i * 256 + tid) + 1, g.n_in);
            }
        }
        // ---- X_0: the first barrier inside retires every read of sP, so the next block's copy
        //      can start right behind it and overlaps all four transforms
        dft16<true>(v0);
#pragma unroll
        for (int k0 = 0; k0 < 16; k0++)
            sA[k0 * F4K_STRIDE + tid] = cmul(v0[pos16(k0)], t1[k0]);
        __syncthreads();
        {
            const long long nxt = blk + gridDim.x;
            if (tid == 0 && nxt < g.n_blocks && tma_block(nxt))
                issue(nxt);
        }
        {
            const int k0 = tid >> 4, n0 = tid & 15;
            float2* row = sA + k0 * F4K_STRIDE + n0;
#pragma unroll
            for (int i = 0; i < 16; i++)
                v0[i] = row[i * 16];
            dft16<true>(v0);
            row[0] = v0[pos16(0)];
#pragma unroll
            for (int k1 = 1; k1 < 16; k1++)
                row[k1 * 16] = cmul(v0[pos16(k1)], sT2[k1 * 16 + n0]);
        }
        __syncthreads();
        {
            const int k0 = tid & 15, k1 = tid >> 4;
            const float2* row = sA + k0 * F4K_STRIDE + k1 * 16;
#pragma unroll
            for (int i = 0; i < 16; i++)
                v0[i] = row[i];
            dft16<true>(v0); // X_0[tid + 256 k2] in v0[pos16(k2)]
        }
        __syncthreads();
        // ---- X_1
        fft4096_passes<true>(v1, sA, sT2, t1, tid);
        // ---- Y_0, Y_1 in place (natural register order for the inverse pass 1)
        {
            float2 a0[16];
#pragma unroll
            for (int k2 = 0; k2 < 16; k2++) {
                const float2 xa = v0[pos16(k2)], xb = v1[pos16(k2)];
                const int k = k2 * 256 + tid;
                const float2 y0 = cmul(xa, __ldg(Htab + g.c00 * OLS_N + k)) + cmul(xb, __ldg(Htab + g.c01 * OLS_N + k));
                const float2 y1 = cmul(xa, __ldg(Htab + g.c10 * OLS_N + k)) + cmul(xb, __ldg(Htab + g.c11 * OLS_N + k));
                a0[k2] = y0;
                v1[pos16(k2)] = y1; // slot pos16(k2) of v1 is dead from here on
            }
            // permute Y_1 from pos16 order into natural order, Y_0 already natural in a0
#pragma unroll
            for (int k2 = 0; k2 < 16; k2++)
                v0[k2] = a0[k2];
        }
        __syncthreads(); // X_1's pass-3 reads of sA are done
        fft4096_passes<false>(v0, sA, sT2, t1, tid);
        __syncthreads();
        {
            float2 b1[16];
#pragma unroll
            for (int k2 = 0; k2 < 16; k2++)
                b1[k2] = v1[pos16(k2)];
            fft4096_passes<false>(b1, sA, sT2, t1, tid);
            // ---- v0[pos16(j)] = y[2 (j0 + n)], b1[pos16(j)] = y[2 (j0 + n) + 1], n = tid + 256 j
            const long long j0 = blk * g.V - g.Ov;
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const int n = tid + 256 * j;
                const long long m = 2 * (j0 + n);
                if (n >= g.Ov && n < g.Ov + g.V && m < g.n_out) {
                    const float2 e = v0[pos16(j)], o = b1[pos16(j)];
                    if (out16 && m + 1 < g.n_out)
                        __stcs(reinterpret_cast<float4*>(y + m), make_float4(e.x, e.y, o.x, o.y));
                    else {
                        __stcs(y + m, e);
                        if (m + 1 < g.n_out)
                            __stcs(y + m + 1, o);
                    }
                }
            }
        }
        __syncthreads();
    }
}

constexpr int OLS_PART = 2048; // taps per partition when the filter does not fit one block

struct ols_plan {
    int T = 0, D = 1;
    int real = 0;
    ols_geom g{};
    int n_parts = 1;
    float2* d_H = nullptr; // [n_parts][4096]
    float2* d_tw1 = nullptr;
    float2* d_tw2 = nullptr;
    int grid = 296;
    int poly = 0;          // polyphase form (fir_olsd_kernel): complex stream, even D <= 8
    float2* d_G = nullptr; // [D][2][4096]: phase spectra, on time / advanced by one output sample
    int pV = 0, pOv = 0;
    int two = 0;           // two-phase form (fir_ols2_kernel): complex stream, D == 1, many taps
    float2* d_H2 = nullptr; // [4][4096]: H_0, H_1, z^-1 H_1, z^+1 H_0
};

void ols_destroy(ols_plan* p)
{
    if (!p)
        return;
    cudaFree(p->d_H);
    cudaFree(p->d_tw1);
    cudaFree(p->d_tw2);
    cudaFree(p->d_G);
    cudaFree(p->d_H2);
    delete p;
}

// segment starts must be 16-byte aligned for the bulk copies: multiples of 2 complex / 4 real samples
static void ols_geometry(int T, int D, int real, int* n_parts, int* Ov, int* V)
{
    const int q = real ? 4 : 2;
    int Tp = T;
    *n_parts = 1;
    if (((T - 1) + q - 1) / q * q > OLS_N - 1024) { // does not leave >= 1024 valid samples: partition
        *n_parts = (T + OLS_PART - 1) / OLS_PART;
        Tp = OLS_PART;
    }
    *Ov = ((Tp - 1) + q - 1) / q * q;
    *V = (OLS_N - *Ov) / (q * D) * (q * D);
}

static bool olsd_supported(int T, int D, int real)
{
    if (real || D < 2 || D > OLSD_MAXD)
        return false;
    if (const char* e = getenv("B200_OLS_POLY"))
        if (atoi(e) == 0)
            return false;
    const int Tq = (T + D - 1) / D;
    return Tq >= 1 && OLS_N - Tq >= 1024;
}

// two-phase form: worth it once the overlap of the one-phase form eats a quarter of the block
static bool ols2_supported(int T, int D, int real)
{
    if (real || D != 1)
        return false;
    int min_taps = 1024;
    if (const char* e = getenv("B200_OLS_TWO"))
        min_taps = atoi(e) > 0 ? atoi(e) : (1 << 30);
    const int Tq = (T + 1) / 2;
    return T >= min_taps && OLS_N - 1 - Tq >= 1024;
}

int ols_polyphase(int T, int D, int real) { return olsd_supported(T, D, real) ? ((D & 1) ? 2 : 1) : 0; }

bool ols_supported(int T, int D, int real)
{
    if (olsd_supported(T, D, real) || ols2_supported(T, D, real))
        return true;
    int np, Ov, V;
    ols_geometry(T, D, real, &np, &Ov, &V);
    return T >= 2 && V >= 4 * D && np <= 16;
}

int ols_create(const float* taps, int T, int D, int real, int fuse, float kre, float kim, ols_plan** out)
{
    *out = nullptr;
    if (real)
        kim = 0.f; // fff: the fused constant is real
    if (!ols_supported(T, D, real))
        return set_err(B200_ERR_UNSUPPORTED, "fir overlap-save: %d taps / decimation %d do not fit the 4096-point block", T, D);
    ols_plan* p = new ols_plan();
    p->T = T;
    p->D = D;
    p->real = real;
    p->poly = olsd_supported(T, D, real) ? 1 : 0;
    if (p->poly) {
        const int Tq = (T + D - 1) / D;
        p->pOv = Tq - 1;
        p->pV = OLS_N - Tq;
        p->n_parts = 1;
        p->g.Ov = p->pOv;
        p->g.V = p->pV;
    } else if (ols2_supported(T, D, real)) {
        p->two = 1;
        p->pOv = (T + 1) / 2;
        p->pV = OLS_N - 1 - p->pOv;
        p->n_parts = 1;
        p->g.Ov = p->pOv;
        p->g.V = p->pV;
    } else
        ols_geometry(T, D, real, &p->n_parts, &p->g.Ov, &p->g.V);
    p->g.D = D;
    p->g.Tm1 = T - 1;
    // H[k] = sum_n h[n] e^{-j 2 pi k n / N} / N, times the fused multiply_const
    std::vector<double> cs(2 * OLS_N);
    for (int i = 0; i < OLS_N; i++) {
        cs[2 * i] = std::cos(2.0 * M_PI * i / OLS_N);
        cs[2 * i + 1] = -std::sin(2.0 * M_PI * i / OLS_N);
    }
    const double fr = fuse ? kre : 1.0, fi = fuse ? kim : 0.0;
    std::vector<float2> H((size_t)OLS_N * p->n_parts), t1(16 * 256), t2(256);
    const int Lp = p->n_parts == 1 ? T : OLS_PART;
    std::vector<float2> G;
    if (p->poly) {
        // G[ph][late][k] = sum_q h[q D + ph] e^{-j 2 pi k (q - late) / N} / N, times the fused constant
        G.resize((size_t)D * 2 * OLS_N);
        for (int ph = 0; ph < D; ph++)
            for (int late = 0; late < 2; late++)
                for (int k = 0; k < OLS_N; k++) {
                    double re = 0, im = 0;
                    for (int q = 0; q * D + ph < T; q++) {
                        int idx = (int)(((long long)k * (q - late)) & (OLS_N - 1));
                        re += taps[q * D + ph] * cs[2 * idx];
                        im += taps[q * D + ph] * cs[2 * idx + 1];
                    }
                    re /= OLS_N;
                    im /= OLS_N;
                    G[((size_t)ph * 2 + late) * OLS_N + k] =
                        make_float2((float)(re * fr - im * fi), (float)(re * fi + im * fr));
                }
    }
    std::vector<float2> H2;
    if (p->two) {
        H2.resize((size_t)4 * OLS_N);
        for (int tab = 0; tab < 4; tab++) {
            const int ph = (tab == 1 || tab == 2) ? 1 : 0;      // tap phase
            const int dl = tab == 2 ? 1 : tab == 3 ? -1 : 0;    // delay in half-rate samples
            for (int k = 0; k < OLS_N; k++) {
                double re = 0, im = 0;
                for (int q = 0; 2 * q + ph < T; q++) {
                    int idx = (int)(((long long)k * (q + dl)) & (OLS_N - 1));
                    re += taps[2 * q + ph] * cs[2 * idx];
                    im += taps[2 * q + ph] * cs[2 * idx + 1];
                }
                re /= OLS_N;
                im /= OLS_N;
                H2[(size_t)tab * OLS_N + k] = make_float2((float)(re * fr - im * fi), (float)(re * fi + im * fr));
            }
        }
    }
    for (int part = 0; part < ((p->poly || p->two) ? 0 : p->n_parts); part++) {
        const int t0 = part * Lp, tn = std::min(T - t0, Lp);
        for (int k = 0; k < OLS_N; k++) {
            double re = 0, im = 0;
            for (int n = 0; n < tn; n++) {
                int idx = (int)(((long long)k * n) & (OLS_N - 1));
                re += taps[t0 + n] * cs[2 * idx];
                im += taps[t0 + n] * cs[2 * idx + 1];
            }
            re /= OLS_N;
            im /= OLS_N;
            // layout [k2][tid] with k = tid + 256 k2  ==  plain index k
            H[(size_t)part * OLS_N + k] = make_float2((float)(re * fr - im * fi), (float)(re * fi + im * fr));
        }
    }
    for (int k0 = 0; k0 < 16; k0++)
        for (int L = 0; L < 256; L++) {
            double ang = -2.0 * M_PI * (double)((L * k0) % 4096) / 4096.0;
            t1[k0 * 256 + L] = make_float2((float)std::cos(ang), (float)std::sin(ang));
        }
    for (int k1 = 0; k1 < 16; k1++)
        for (int n0 = 0; n0 < 16; n0++) {
            double ang = -2.0 * M_PI * (double)(n0 * k1) / 256.0;
            t2[k1 * 16 + n0] = make_float2((float)std::cos(ang), (float)std::sin(ang));
        }
#define OLS_CUDA(call)                                                                       \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            ols_destroy(p);                                                                  \
            return set_err(B200_ERR_CUDA, "fir overlap-save: %s -> %s", #call, cudaGetErrorString(e__)); \
        }                                                                                    \
    } while (0)
    OLS_CUDA(cudaMalloc(&p->d_H, sizeof(float2) * H.size()));
    OLS_CUDA(cudaMemcpy(p->d_H, H.data(), sizeof(float2) * H.size(), cudaMemcpyHostToDevice));
    if (p->two) {
        OLS_CUDA(cudaMalloc(&p->d_H2, sizeof(float2) * H2.size()));
        OLS_CUDA(cudaMemcpy(p->d_H2, H2.data(), sizeof(float2) * H2.size(), cudaMemcpyHostToDevice));
        OLS_CUDA(cudaFuncSetAttribute(fir_ols2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)OLSD_SMEM));
    }
    if (p->poly) {
        OLS_CUDA(cudaMalloc(&p->d_G, sizeof(float2) * G.size()));
        OLS_CUDA(cudaMemcpy(p->d_G, G.data(), sizeof(float2) * G.size(), cudaMemcpyHostToDevice));
        OLS_CUDA(cudaFuncSetAttribute(fir_olsd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)OLSD_SMEM));
    }
    OLS_CUDA(cudaMalloc(&p->d_tw1, sizeof(float2) * t1.size()));
    OLS_CUDA(cudaMemcpy(p->d_tw1, t1.data(), sizeof(float2) * t1.size(), cudaMemcpyHostToDevice));
    OLS_CUDA(cudaMalloc(&p->d_tw2, sizeof(float2) * t2.size()));
    OLS_CUDA(cudaMemcpy(p->d_tw2, t2.data(), sizeof(float2) * t2.size(), cudaMemcpyHostToDevice));
    OLS_CUDA(cudaFuncSetAttribute(fir_ols4096_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)OLS_SMEM));
    OLS_CUDA(cudaFuncSetAttribute(fir_ols4096_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)OLS_SMEM));
#undef OLS_CUDA
    p->grid = 2 * sm_count();
    // polyphase form with D > 10: a block's input span is D * 32 KiB and every pair plane pass re-reads it
    // through L2; with 2 CTAs per SM the spans of all CTAs (D = 16: 151 MB) no longer fit the 126 MB L2
    // and each pass goes back to HBM (measured 96 GS/s).  One CTA per SM keeps the working set resident.
    if (p->poly && D > 10)
        p->grid = sm_count();
    *out = p;
    return B200_OK;
}

// polyphase form: geometry of the (row, slot) view for this call's pointer alignment, tensor map, launch
static int olsd_launch(ols_plan* p, const float* d_hist, const void* d_in, void* d_out, long long n_in,
                       long long n_out, cudaStream_t s)
{
    const int D = p->D;
    olsd_geom g{};
    g.D = D;
    g.Ov = p->pOv;
    g.V = p->pV;
    g.Tm1 = p->T - 1;
    g.n_in = n_in;
    g.n_out = n_out;
    g.n_blocks = (n_out + g.V - 1) / g.V;
    // rows of D samples start at the 16-byte granule at or below d_in; `off` samples precede x[0] in it
    const uintptr_t a = (uintptr_t)d_in;
    const int off = (a % 16 == 8) ? 1 : 0;
    g.gbase = -off;
    g.off = off;
    g.rmin = (D - 1 > off) ? -1 : 0; // some phase sits one row earlier unless D == 2 and off == 1
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    const long long nrows = (n_in + off) / D;
    g.tma_ok = 0;
    if (a % 8 == 0 && (D & 1) == 0 && nrows >= OLS_N && nrows < (1LL << 31)) {
        if (tmap_encode_fn enc = tmap_encode_tiled()) {
            cuuint64_t gdim[2] = { (cuuint64_t)D, (cuuint64_t)nrows };
            cuuint64_t gstride[1] = { (cuuint64_t)D * 8 };
            cuuint32_t box[2] = { 2, 256 };
            cuuint32_t estr[2] = { 1, 1 };
            CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, (void*)(a - 8 * off), gdim, gstride, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            g.tma_ok = (r == CUDA_SUCCESS) ? 1 : 0;
        }
    }
    if (const char* e = getenv("B200_OLS_TMA"))
        if (atoi(e) == 0)
            g.tma_ok = 0;
    const long long grid = g.n_blocks < p->grid ? g.n_blocks : p->grid;
    B200_LAUNCH_PDL(fir_olsd_kernel, (unsigned)grid, 256, OLSD_SMEM, s, (const float2*)d_in, (const float2*)d_hist,
                (float2*)d_out, p->d_G, p->d_tw1, p->d_tw2, tmap, g);
    return B200_OK;
}

static int ols2_launch(ols_plan* p, const float* d_hist, const void* d_in, void* d_out, long long n_in,
                       long long n_out, cudaStream_t s)
{
    ols2_geom g{};
    g.Ov = p->pOv;
    g.V = p->pV;
    g.Tm1 = p->T - 1;
    g.n_in = n_in;
    g.n_out = n_out;
    g.n_blocks = (n_out + 2LL * g.V - 1) / (2LL * g.V);
    const uintptr_t a = (uintptr_t)d_in;
    g.off = (a % 16 == 8) ? 1 : 0;
    g.tma_ok = (a % 8 == 0) ? 1 : 0;
    if (g.off == 0) { // Y_0 = H_0 X_0 + z^-1 H_1 X_1 ; Y_1 = H_1 X_0 + H_0 X_1
        g.c00 = 0, g.c01 = 2, g.c10 = 1, g.c11 = 0;
    } else {          // aligned streams x'_0 = x_1 delayed, x'_1 = x_0: Y_0 = H_1 X'_0 + H_0 X'_1 ; Y_1 = z H_0 X'_0 + H_1 X'_1
        g.c00 = 1, g.c01 = 0, g.c10 = 3, g.c11 = 1;
    }
    const long long grid = g.n_blocks < p->grid ? g.n_blocks : p->grid;
    B200_LAUNCH_PDL(fir_ols2_kernel, (unsigned)grid, 256, OLSD_SMEM, s, (const float2*)d_in, (const float2*)d_hist,
                (float2*)d_out, p->d_H2, p->d_tw1, p->d_tw2, g);
    return B200_OK;
}

int ols_launch(ols_plan* p, const float* d_hist, const void* d_in, void* d_out, long long n_in,
               long long n_out, cudaStream_t s)
{
    if (n_out <= 0)
        return B200_OK;
    if (p->poly)
        return olsd_launch(p, d_hist, d_in, d_out, n_in, n_out, s);
    if (p->two)
        return ols2_launch(p, d_hist, d_in, d_out, n_in, n_out, s);
    ols_geom g = p->g;
    g.n_in = n_in;
    g.n_out = n_out;
    const long long covered = n_out * p->D; // input-rate samples that carry an output
    const long long per_block = (long long)g.V * (p->real ? 2 : 1);
    g.n_blocks = (covered + per_block - 1) / per_block;
    g.tma_ok = ((uintptr_t)d_in % 16 == 0) ? 1 : 0; // segment starts are multiples of 16 bytes
    long long grid = g.n_blocks < p->grid ? g.n_blocks : p->grid;
    // uniformly partitioned convolution: y = sum_p (h_p * x delayed by p*2048), one pass each
    for (int part = 0; part < p->n_parts; part++) {
        g.shift = part * OLS_PART;
        g.accumulate = part > 0;
        if (p->real)
            B200_LAUNCH_PDL(fir_ols4096_kernel<true>, (unsigned)grid, 256, OLS_SMEM, s, (const float2*)d_in,
                        (const float2*)d_hist, (float2*)d_out, p->d_H + (size_t)part * OLS_N, p->d_tw1,
                        p->d_tw2, g);
        else
            B200_LAUNCH_PDL(fir_ols4096_kernel<false>, (unsigned)grid, 256, OLS_SMEM, s, (const float2*)d_in,
                        (const float2*)d_hist, (float2*)d_out, p->d_H + (size_t)part * OLS_N, p->d_tw1,
                        p->d_tw2, g);
    }
    return B200_OK;
}

} // namespace b200
