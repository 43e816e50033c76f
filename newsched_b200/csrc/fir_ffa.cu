// fir_ffa.cu -- full-rate FIR (ccf / fff, D = 1) by the 2-parallel fast FIR algorithm (algorithm 5).
//
// The direct form (fir.cu) is FP32-pipe bound from ~32 taps on and already runs its inner loop as
// packed FFMA2 at ~75 % of the measured FMA peak; the only way past that is to execute fewer FMAs.
// With the even / odd half-rate streams X_0[j] = x[2j], X_1[j] = x[2j+1] and tap phases
// H_0[q] = h[2q], H_1[q] = h[2q+1]:
//     A = H_0 * X_0,   B = H_1 * X_1,   C = (H_0 + H_1) * (X_0 + X_1)
//     y[2m]   = A[m] + B[m-1]
//     y[2m+1] = C[m] - A[m] - B[m]
// i.e. three T/2-tap filters per TWO outputs: 0.75 T products per output instead of T.
//
// One CTA produces 4096 outputs (2048 per parity).  The input tile arrives as the raw stream by TMA
// and is de-interleaved in place into the two phase planes (same 128B-swizzled row layout as
// fir.cu, so the per-thread 128 B-strided LDS.128 window reads are conflict-free).
// Each thread owns 16 consecutive half-rate outputs m and runs three register-blocked passes with
// the 64-float circular window of fir.cu:
//   pass A  (plane 0, taps H_0)            -> A[16], parked in a private shared row
//   pass B  (plane 1, taps H_1, window one sample earlier, 17 accumulators) -> B[m-1 .. m+15], so
//           both B[m-1] and B[m] of every m are thread-local: no neighbour exchange, no edge case
//           ye = A + B[m-1] is parked in the private row, -(A + B[m]) initialises pass C
//   pass C  (plane 0 + plane 1 summed while loading the window, taps H_0 + H_1) -> y[2m+1]
// and the 32 interleaved outputs of a thread are two 128-byte rows of the output tile, written back
// by one TMA tensor store.  49 T/2 packed FMAs per 32 outputs instead of 64 T/2.
// An output's instruction sequence depends only on the parity of its index within the call, so a
// stream cut at even offsets reproduces the one-shot result bit for bit; odd cuts and the direct
// form differ by fp32 rounding of the regrouped sums (tests: <= 1e-5 rel. RMS).
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "fir_ffa.cuh"

namespace b200 {

constexpr int FFA_NT = 128;  // threads per CTA
constexpr int FFA_RING = 64; // register window (floats)

struct ffa_geom {
    int Tm1, TQ;        // taps - 1; taps per phase, padded to a multiple of 2 steps
    int plane_rows;     // rows of 32 floats per phase plane
    int tma_in_ok, tma_out_ok;
    long long full_in_rows, full_out_rows; // rows of 32 floats completely inside the input / output
    long long n_in, n_out;
    int fuse;
    float kre, kim;
};

__device__ __forceinline__ int ffa_swz(int f)
{
    const int row = f >> 5, c = (f >> 2) & 7;
    return (row << 5) | ((c ^ (row & 7)) << 2) | (f & 3);
}

template <int VEC>
__device__ __forceinline__ void ffa_fetch(const float* __restrict__ x, const float* __restrict__ hist, int Tm1,
                                          long long g, long long n_in, float* v)
{
    const float* src = nullptr;
    if (g >= 0) {
        if (g < n_in)
            src = x + g * VEC;
    } else if (hist && g >= -(long long)Tm1)
        src = hist + ((long long)Tm1 + g) * VEC;
    if (VEC == 2) {
        float2 t = src ? __ldg(reinterpret_cast<const float2*>(src)) : make_float2(0.f, 0.f);
        v[0] = t.x;
        v[1] = t.y;
    } else {
        v[0] = src ? __ldg(src) : 0.f;
        v[1] = 0.f;
    }
}

// one row (32 floats) of a swizzled plane into one half of the ring; SUM: plane + plane2
template <int HALF, bool SUM>
__device__ __forceinline__ void ffa_load_half(float (&W)[FFA_RING], const float* __restrict__ plane,
                                              const float* __restrict__ plane2, int row)
{
    const float* rb = plane + (row << 5);
    const float* rb2 = plane2 + (row << 5);
    const int s = (row & 7) << 2;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        float4 t = *reinterpret_cast<const float4*>(rb + ((j << 2) ^ s));
        if (SUM) {
            const float4 u = *reinterpret_cast<const float4*>(rb2 + ((j << 2) ^ s));
            const float2 a = __fadd2_rn(make_float2(t.x, t.y), make_float2(u.x, u.y));
            const float2 b = __fadd2_rn(make_float2(t.z, t.w), make_float2(u.z, u.w));
            t = make_float4(a.x, a.y, b.x, b.y);
        }
        W[HALF * 32 + 4 * j + 0] = t.x;
        W[HALF * 32 + 4 * j + 1] = t.y;
        W[HALF * 32 + 4 * j + 2] = t.z;
        W[HALF * 32 + 4 * j + 3] = t.w;
    }
}

// One step of CH taps against the ring.  Tap q' of the step meets ring sample (q' + 1 - EARLY)
// relative to the accumulator's own index; NACC accumulator floats.
template <int VEC, int NACC, int EARLY, int OFF>
__device__ __forceinline__ void ffa_step(float (&acc)[NACC], const float (&W)[FFA_RING],
                                         const float* __restrict__ hs)
{
    constexpr int CH = 32 / VEC;
#pragma unroll
    for (int q4 = 0; q4 < CH; q4 += 4) {
        const float4 h4 = *reinterpret_cast<const float4*>(hs + q4);
        const float hv[4] = { h4.x, h4.y, h4.z, h4.w };
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (VEC == 2) {
                const float2 h2 = make_float2(hv[u], hv[u]);
#pragma unroll
                for (int l = 0; l < NACC; l += 2) {
                    const int i = (OFF + (q4 + u + 1 - EARLY) * VEC + l) % FFA_RING;
                    const float2 a = __ffma2_rn(make_float2(W[i], W[i + 1]), h2, make_float2(acc[l], acc[l + 1]));
                    acc[l] = a.x;
                    acc[l + 1] = a.y;
                }
            } else {
#pragma unroll
                for (int l = 0; l < NACC; l++)
                    acc[l] = fmaf(hv[u], W[(OFF + (q4 + u + 1 - EARLY) * VEC + l) % FFA_RING], acc[l]);
            }
        }
    }
}

template <int VEC, int NACC, int EARLY, bool SUM>
__device__ __forceinline__ void ffa_pass(float (&acc)[NACC], const float* __restrict__ plane,
                                         const float* __restrict__ plane2, const float* __restrict__ hp,
                                         int nsteps, int tid)
{
    constexpr int CH = 32 / VEC;
    float W[FFA_RING];
    ffa_load_half<0, SUM>(W, plane, plane2, tid);
    for (int b = 0; b < nsteps; b += 2) {
        ffa_load_half<1, SUM>(W, plane, plane2, tid + b + 1);
        ffa_step<VEC, NACC, EARLY, 0>(acc, W, hp + b * CH);
        ffa_load_half<0, SUM>(W, plane, plane2, tid + b + 2);
        ffa_step<VEC, NACC, EARLY, 32>(acc, W, hp + (b + 1) * CH);
    }
}

// one private 128-byte row <-> 32 registers (swizzled by the row's own index)
__device__ __forceinline__ void ffa_row_store(float* base, int row, const float* v)
{
    float* rb = base + (row << 5);
    const int s = (row & 7) << 2;
#pragma unroll
    for (int j = 0; j < 8; j++)
        *reinterpret_cast<float4*>(rb + ((j << 2) ^ s)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
__device__ __forceinline__ void ffa_row_load(const float* base, int row, float* v)
{
    const float* rb = base + (row << 5);
    const int s = (row & 7) << 2;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const float4 t = *reinterpret_cast<const float4*>(rb + ((j << 2) ^ s));
        v[4 * j] = t.x;
        v[4 * j + 1] = t.y;
        v[4 * j + 2] = t.z;
        v[4 * j + 3] = t.w;
    }
}

__device__ __forceinline__ void ffa_tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1,
                                                uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
                 "[%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(smem_dst)),
                 "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}

// smem: [mbarrier 16 B][taps 3*TQ floats][pad to 1024 B][plane 0][plane 1][128 private rows]
// Interior tiles: the 2*plane_rows rows of the tile arrive as the raw interleaved stream by TMA on
// top of the plane region; every thread pulls two of those rows into registers, and after a
// barrier writes the even samples as its row of plane 0 and the odd ones as its row of plane 1
// (an in-place de-interleave: 16 LDS.128 + 16 STS.128 per thread, no per-sample copies).
// The output tile (256 rows) is written over the planes once every thread is done with them.
template <int VEC>
__global__ void __launch_bounds__(FFA_NT, 4)
    fir_ffa_kernel(const float* __restrict__ x, const float* __restrict__ hist, float* __restrict__ y,
                   const float* __restrict__ taps3, const __grid_constant__ CUtensorMap tmap_in,
                   const __grid_constant__ CUtensorMap tmap_tail, const __grid_constant__ CUtensorMap tmap_out,
                   ffa_geom gm)
{
    constexpr int R = 32 / VEC;       // half-rate outputs per thread
    constexpr int MTh = FFA_NT * R;   // half-rate outputs per tile
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    float* hs = reinterpret_cast<float*>(smem_raw + 16);
    const int TQ = gm.TQ;
    float* planes;
    {
        uint32_t a = smem_u32(hs + 3 * TQ);
        uint32_t aligned = (a + 1023u) & ~1023u;
        planes = hs + 3 * TQ + (aligned - a) / 4;
    }
    const int plane_f = gm.plane_rows << 5;
    float* rows = planes + ((2 * plane_f + 255) & ~255); // 1024-byte aligned: TMA swizzles by address bits
    const int tid = threadIdx.x;
    const long long tile = blockIdx.x;
    const long long B0 = tile * MTh - TQ; // half-rate index of plane element 0
    const int PLs = (gm.plane_rows << 5) / VEC;
    const int nsteps = TQ / (32 / VEC); // even by construction; plane_rows = FFA_NT + nsteps

    const long long row0 = B0 * VEC / 16; // first row (of 32 floats) of the tile in the input stream
    const bool use_tma = gm.tma_in_ok && B0 >= 0 && row0 + 2 * gm.plane_rows <= gm.full_in_rows;
    if (use_tma && tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        mbar_arrive_expect_tx(bar, (uint32_t)(2 * gm.plane_rows) * 128u);
        // rows 0..255 of the tile on top of the plane region, the 2*nsteps tail rows into the
        // (still unused) private-row region
        ffa_tma_load_2d(planes, &tmap_in, 0, (int)row0, bar);
        ffa_tma_load_2d(rows, &tmap_tail, 0, (int)(row0 + 2 * FFA_NT), bar);
    }
    for (int i = tid; i < 3 * TQ; i += FFA_NT)
        hs[i] = __ldg(taps3 + i);
    if (use_tma) {
        __syncthreads(); // mbarrier initialised
        mbar_wait(bar, 0);
        // staging rows 2t, 2t+1 -> row t of both planes
        float S[64], p0[32], p1[32];
        ffa_row_load(planes, 2 * tid, S);
        ffa_row_load(planes, 2 * tid + 1, S + 32);
        __syncthreads(); // every staging row of the plane region is in registers
#pragma unroll
        for (int i = 0; i < 32; i++) {
            p0[i] = S[(2 * (i / VEC)) * VEC + i % VEC];
            p1[i] = S[(2 * (i / VEC) + 1) * VEC + i % VEC];
        }
        ffa_row_store(planes, tid, p0);
        ffa_row_store(planes + plane_f, tid, p1);
        if (tid < nsteps) { // tail rows 2(128+t), +1 -> row 128+t of both planes
            ffa_row_load(rows, 2 * tid, S);
            ffa_row_load(rows, 2 * tid + 1, S + 32);
#pragma unroll
            for (int i = 0; i < 32; i++) {
                p0[i] = S[(2 * (i / VEC)) * VEC + i % VEC];
                p1[i] = S[(2 * (i / VEC) + 1) * VEC + i % VEC];
            }
            ffa_row_store(planes, FFA_NT + tid, p0);
            ffa_row_store(planes + plane_f, FFA_NT + tid, p1);
        }
    } else {
        // edge tiles / unaligned input: element e of plane c  <-  x[2 (B0 + e) + c], sample by sample
        const long long g_lo = 2 * B0;
        const int total = 2 * PLs;
        for (int i = tid; i < total; i += FFA_NT) {
            const int e = i >> 1, c = i & 1;
            float* dst = planes + c * plane_f + ffa_swz(e * VEC);
            const long long g = g_lo + i;
            if (g >= 0 && g < gm.n_in) {
                if (VEC == 2)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(x + g * 2)
                                 : "memory");
                else
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(x + g)
                                 : "memory");
            } else {
                float v[2];
                ffa_fetch<VEC>(x, hist, gm.Tm1, g, gm.n_in, v);
                if (VEC == 2)
                    *reinterpret_cast<float2*>(dst) = make_float2(v[0], v[1]);
                else
                    dst[0] = v[0];
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();

    const float* P0 = planes;
    const float* P1 = planes + plane_f;
    float yo[32];
    {
        // ---- pass A -> private row
        float a[32];
#pragma unroll
        for (int l = 0; l < 32; l++)
            a[l] = 0.f;
        ffa_pass<VEC, 32, 0, false>(a, P0, P0, hs, nsteps, tid);
        ffa_row_store(rows, tid, a);
    }
    {
        // ---- pass B, one sample early, R + 1 outputs: bq[r'] = B[m0 + r' - 1]
        float bq[32 + VEC];
#pragma unroll
        for (int l = 0; l < 32 + VEC; l++)
            bq[l] = 0.f;
        ffa_pass<VEC, 32 + VEC, 1, false>(bq, P1, P1, hs + TQ, nsteps, tid);
        float a[32], ye[32];
        ffa_row_load(rows, tid, a);
#pragma unroll
        for (int l = 0; l < 32; l++) {
            ye[l] = a[l] + bq[l];          // y[2m]   = A[m] + B[m-1]
            yo[l] = -(a[l] + bq[l + VEC]); // y[2m+1] = C[m] - (A[m] + B[m]): C accumulates on top
        }
        ffa_row_store(rows, tid, ye);
    }
    // ---- pass C on the summed planes
    ffa_pass<VEC, 32, 0, true>(yo, P0, P1, hs + 2 * TQ, nsteps, tid);
    __syncthreads(); // every thread is done with the planes: the output tile goes on top of them

    {
        float ye[32];
        ffa_row_load(rows, tid, ye);
        // interleave: output row 2 tid holds m = 0 .. R/2-1 of this thread, row 2 tid + 1 the rest
        float o0[32], o1[32];
#pragma unroll
        for (int r = 0; r < R / 2; r++)
#pragma unroll
            for (int c = 0; c < VEC; c++) {
                o0[(2 * r) * VEC + c] = ye[r * VEC + c];
                o0[(2 * r + 1) * VEC + c] = yo[r * VEC + c];
                o1[(2 * r) * VEC + c] = ye[(r + R / 2) * VEC + c];
                o1[(2 * r + 1) * VEC + c] = yo[(r + R / 2) * VEC + c];
            }
        if (gm.fuse) {
            if (VEC == 2) {
#pragma unroll
                for (int l = 0; l < 32; l += 2) {
                    float2 v = cmul_nofma(make_float2(o0[l], o0[l + 1]), gm.kre, gm.kim);
                    o0[l] = v.x, o0[l + 1] = v.y;
                    v = cmul_nofma(make_float2(o1[l], o1[l + 1]), gm.kre, gm.kim);
                    o1[l] = v.x, o1[l + 1] = v.y;
                }
            } else {
#pragma unroll
                for (int l = 0; l < 32; l++) {
                    o0[l] = __fmul_rn(o0[l], gm.kre);
                    o1[l] = __fmul_rn(o1[l], gm.kre);
                }
            }
        }
        ffa_row_store(planes, 2 * tid, o0);
        ffa_row_store(planes, 2 * tid + 1, o1);
    }
    const long long orow0 = tile * (2 * FFA_NT); // 256 output rows per tile
    if (gm.tma_out_ok && orow0 + 2 * FFA_NT <= gm.full_out_rows) {
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                             &tmap_out),
                         "r"(0), "r"((int)orow0), "r"(smem_u32(planes))
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        return;
    }
    __syncthreads();
    const long long O0 = tile * (2LL * MTh);
#pragma unroll 4
    for (int i = tid; i < 2 * MTh; i += FFA_NT) {
        const long long m = O0 + i;
        if (m >= gm.n_out)
            break;
        const float* src = planes + ffa_swz(i * VEC);
        if (VEC == 2)
            __stcs(reinterpret_cast<float2*>(y) + m, *reinterpret_cast<const float2*>(src));
        else
            __stcs(y + m, src[0]);
    }
}

struct ffa_plan {
    int T = 0, vec = 2;
    ffa_geom g{};
    size_t smem = 0;
    float* d_taps3 = nullptr;
};

void ffa_destroy(ffa_plan* p)
{
    if (!p)
        return;
    cudaFree(p->d_taps3);
    delete p;
}

bool ffa_supported(int T, int D, int real)
{
    (void)real;
    if (D != 1 || T < 4)
        return false;
    const int step2 = 2 * (32 / (real ? 1 : 2)); // taps per phase come in pairs of steps
    const int TQ = ((T + 1) / 2 + step2 - 1) / step2 * step2;
    return TQ <= 1024 && TQ / (32 / (real ? 1 : 2)) <= 64; // tail rows (2 per step) fit the private-row region
}

int ffa_create(const float* taps, int T, int real, int fuse, float kre, float kim, ffa_plan** out)
{
    *out = nullptr;
    if (!ffa_supported(T, 1, real))
        return set_err(B200_ERR_UNSUPPORTED, "fir 2-parallel form: %d taps not supported", T);
    ffa_plan* p = new ffa_plan();
    p->T = T;
    p->vec = real ? 1 : 2;
    const int CH = 32 / p->vec, step2 = 2 * CH;
    ffa_geom& g = p->g;
    g.Tm1 = T - 1;
    g.TQ = ((T + 1) / 2 + step2 - 1) / step2 * step2;
    g.plane_rows = FFA_NT + g.TQ / CH; // one row per thread + one per tap step
    g.fuse = fuse;
    g.kre = kre;
    g.kim = real ? 0.f : kim;
    const int plane_f = g.plane_rows * 32;
    p->smem = 16 + sizeof(float) * ((size_t)3 * g.TQ + 2 * (size_t)plane_f + (size_t)FFA_NT * 32) + 2048;
    // hs[c][q'] = H_c[TQ - 1 - q'] (reversed: the window walks forward in time), c = 2: H_0 + H_1
    std::vector<float> t3((size_t)3 * g.TQ, 0.f);
    for (int qr = 0; qr < g.TQ; qr++) {
        const long long q = g.TQ - 1 - qr;
        const float h0 = 2 * q < T ? taps[2 * q] : 0.f;
        const float h1 = 2 * q + 1 < T ? taps[2 * q + 1] : 0.f;
        t3[qr] = h0;
        t3[(size_t)g.TQ + qr] = h1;
        t3[(size_t)2 * g.TQ + qr] = h0 + h1;
    }
    cudaError_t e = cudaMalloc(&p->d_taps3, t3.size() * sizeof(float));
    if (e == cudaSuccess)
        e = cudaMemcpy(p->d_taps3, t3.data(), t3.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(fir_ffa_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(fir_ffa_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) {
        ffa_destroy(p);
        return set_err(B200_ERR_CUDA, "fir 2-parallel form: %s", cudaGetErrorString(e));
    }
    *out = p;
    return B200_OK;
}

int ffa_launch(ffa_plan* p, const float* d_hist, const void* d_in, void* d_out, long long n_in, long long n_out,
               cudaStream_t s)
{
    if (n_out <= 0)
        return B200_OK;
    ffa_geom g = p->g;
    g.n_in = n_in;
    g.n_out = n_out;
    const long long per_tile = 2LL * FFA_NT * (32 / p->vec);
    const long long tiles = (n_out + per_tile - 1) / per_tile;
    if (tiles > 0x7fffffffLL)
        return set_err(B200_ERR_ARG, "fir: too many items for one call");
    CUtensorMap tmap_out, tmap_in, tmap_tail;
    memset(&tmap_out, 0, sizeof(tmap_out));
    memset(&tmap_in, 0, sizeof(tmap_in));
    memset(&tmap_tail, 0, sizeof(tmap_tail));
    g.full_out_rows = n_out * p->vec / 32;
    g.full_in_rows = n_in * p->vec / 32;
    g.tma_out_ok = g.tma_in_ok = 0;
    if (g.full_in_rows >= 2 * g.plane_rows && (uintptr_t)d_in % 16 == 0) {
        if (tmap_encode_fn enc = tmap_encode_tiled()) {
            cuuint64_t gdim[2] = { 32, (cuuint64_t)g.full_in_rows };
            cuuint64_t gstride[1] = { 128 };
            cuuint32_t box[2] = { 32, 2 * FFA_NT };
            cuuint32_t boxt[2] = { 32, (cuuint32_t)(2 * (g.plane_rows - FFA_NT)) };
            cuuint32_t estr[2] = { 1, 1 };
            CUresult r = enc(&tmap_in, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(d_in), gdim, gstride, box,
                             estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            CUresult r2 = enc(&tmap_tail, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(d_in), gdim, gstride,
                              boxt, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            g.tma_in_ok = (r == CUDA_SUCCESS && r2 == CUDA_SUCCESS) ? 1 : 0;
        }
    }
    if (const char* e = getenv("B200_FIR_TMA"))
        if (atoi(e) == 0)
            g.tma_in_ok = 0;
    if (g.full_out_rows >= 2 * FFA_NT && (uintptr_t)d_out % 16 == 0) {
        if (tmap_encode_fn enc = tmap_encode_tiled()) {
            cuuint64_t gdim[2] = { 32, (cuuint64_t)g.full_out_rows };
            cuuint64_t gstride[1] = { 128 };
            cuuint32_t box[2] = { 32, 2 * FFA_NT };
            cuuint32_t estr[2] = { 1, 1 };
            CUresult r = enc(&tmap_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d_out, gdim, gstride, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            g.tma_out_ok = (r == CUDA_SUCCESS) ? 1 : 0;
        }
    }
    if (p->vec == 2)
        B200_LAUNCH(fir_ffa_kernel<2>, (unsigned)tiles, FFA_NT, p->smem, s, (const float*)d_in, d_hist,
                    (float*)d_out, p->d_taps3, tmap_in, tmap_tail, tmap_out, g);
    else
        B200_LAUNCH(fir_ffa_kernel<1>, (unsigned)tiles, FFA_NT, p->smem, s, (const float*)d_in, d_hist,
                    (float*)d_out, p->d_taps3, tmap_in, tmap_tail, tmap_out, g);
    return B200_OK;
}

} // namespace b200
