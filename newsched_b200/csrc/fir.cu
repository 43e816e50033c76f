// fir.cu -- decimating FIR filter, ccf (complex64 stream, real taps) and fff (float stream).
//
//   y[m] = sum_{k<T} h[k] x[m*D - k],  m = 0 .. floor(n_in/D)-1      
//
// The reference snapshot has no FIR block; the work() contract it plugs
// into is gr::block::work (runtime/include/gnuradio/block.hpp:81-85) with
// n_consumed = n_produced*D set by the block (block_work_io.hpp:21,36).
//
// Direct-form SIMT kernel (algorithm 1), FP32-pipe bound for T >= 32:
//   * polyphase split: y = sum_p (h_p * x_p), h_p[q] = h[qD+p], x_p[j] = x[jD-p]; the
//     input tile is staged once into shared memory de-interleaved by phase, so every phase
//     is a unit-stride FIR and decimated-away products are never formed;
//   * each thread owns 32 fp32 accumulators (16 complex / 32 real consecutive outputs) and a
//     64-float circular register window over x_p; one step = 16 (32) taps = 512 (1024) FFMA
//     fed by 8 LDS.128 of samples + 4 (8) LDS.128 of taps (warp-uniform broadcast): the inner
//     loop is > 96 % FFMA issue slots;
//   * shared layout is padded 16 B per 128 B so that the per-thread 128 B-strided LDS.128
//     window reads are bank-conflict free;
//   * outputs go back through shared memory so global stores are fully coalesced.
// History (the T-1 samples before the next input item) lives in a ping-pong pair of small
// device buffers owned by the handle; nothing is re-read from the host and the caller may
// chunk the stream arbitrarily.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "fir_ffa.cuh"
#include "fir_interp.cuh"
#include "fir_ols.cuh"

namespace b200 {

// minimum resident CTAs the compiler must leave room for: 6 caps the kernel at 85 registers (with the
// odd tail step it would otherwise take 96 and drop to 5 CTAs/SM: -7 % on HBM-bound short filters)
#ifndef FIR_MINB
#define FIR_MINB 6
#endif
#ifndef B200_FIR_NT
#define B200_FIR_NT 128
#endif
constexpr int FIR_NT = B200_FIR_NT; // threads per CTA
constexpr int FIR_ACC = 32;  // fp32 accumulators per thread
constexpr int FIR_RING = 64; // register window (floats)

// Shared-memory tile layout = what TMA SWIZZLE_128B produces: rows of 32 floats (128 B), the
// 16-byte chunk c of row r stored at chunk position c ^ (r & 7).  Thread t's window starts at
// row t, so the 8 lanes of a quarter-warp hit 8 different chunk positions: the per-thread
// 128 B-strided LDS.128 reads are bank-conflict free without padding.
__host__ __device__ __forceinline__ int swz(int f)
{
    const int row = f >> 5, c = (f >> 2) & 7;
    return (row << 5) | ((c ^ (row & 7)) << 2) | (f & 3);
}

struct fir_epilogue {
    int fuse;
    float kre, kim;
};

struct fir_geom {
    int Tm1, D, TQ;
    int plane_rows;   // rows (of 32 floats) per phase plane
    int box_rows;     // rows per TMA box
    int n_boxes;      // TMA boxes per tile (D == 1 only)
    int tma_ok;       // input tensor map valid (input 16-byte aligned, enough full rows)
    int tma_out_ok;   // output tensor map valid
    long long full_rows;     // rows of the input that are completely inside [0, n_in)
    long long full_out_rows; // rows of the output completely inside [0, n_out)
    long long n_in, n_out;
    long long n_in_f, n_out_f; // RP mode: bounds in floats (n_in / n_out are in float PAIRS there)
};

// x value at global sample index g (may be negative -> history, or >= n_in -> 0)
template <int VEC>
__device__ __forceinline__ void fir_fetch(const float* __restrict__ x, const float* __restrict__ hist,
                                          int Tm1, long long g, long long n_in, float* v)
{
    const float* src = nullptr;
    if (g >= 0) {
        if (g < n_in)
            src = x + g * VEC;
    } else if (hist && g >= -(long long)Tm1) {
        src = hist + ((long long)Tm1 + g) * VEC;
    }
    if (VEC == 2) {
        float2 t = src ? __ldg(reinterpret_cast<const float2*>(src)) : make_float2(0.f, 0.f);
        v[0] = t.x;
        v[1] = t.y;
    } else {
        v[0] = src ? __ldg(src) : 0.f;
        v[1] = 0.f;
    }
}

// One step of CH taps against the register ring.  OFF = ring offset (floats) of the thread's row;
// tap q' of the step meets ring element (q'+1): the window is read from one sample early so that
// BOTH the input window and the output tile start on 128-byte rows (TMA load and TMA store).
// DD > 1 (decimation folded into the full-rate kernel): only every DD-th output position of the
// window owns an accumulator, the taps stay in natural order.
template <int VEC, int CH, int OFF, int DD = 1>
__device__ __forceinline__ void fir_step(float (&acc)[FIR_ACC], const float (&W)[FIR_RING],
                                         const float* __restrict__ hs)
{
#pragma unroll
    for (int q4 = 0; q4 < CH; q4 += 4) {
        float4 h4 = *reinterpret_cast<const float4*>(hs + q4);
        const float hv[4] = { h4.x, h4.y, h4.z, h4.w };
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (VEC == 2) {
                // complex sample x real tap = one packed FFMA2 (fma.rn.f32x2, tap broadcast):
                // half the issue slots of two FFMAs, so staging / epilogue instructions of the
                // other resident warps issue underneath the FMA pipe
                const float2 h2 = make_float2(hv[u], hv[u]);
#pragma unroll
                for (int l = 0; l < FIR_ACC; l += 2 * DD) {
                    const int i = (OFF + (q4 + u + 1) * VEC + l) % FIR_RING;
                    float2 a = __ffma2_rn(make_float2(W[i], W[i + 1]), h2, make_float2(acc[l], acc[l + 1]));
                    acc[l] = a.x;
                    acc[l + 1] = a.y;
                }
            } else {
#pragma unroll
                for (int l = 0; l < FIR_ACC; l += DD)
                    acc[l] = fmaf(hv[u], W[(OFF + (q4 + u + 1) * VEC + l) % FIR_RING], acc[l]);
            }
        }
    }
}

// 32 floats of row `row` of a swizzled plane into one half of the register ring
template <int HALF>
__device__ __forceinline__ void fir_load_half(float (&W)[FIR_RING], const float* __restrict__ plane, int row)
{
    const float* rb = plane + (row << 5);
    const int s = (row & 7) << 2;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        float4 t = *reinterpret_cast<const float4*>(rb + ((j << 2) ^ s));
        W[HALF * 32 + 4 * j + 0] = t.x;
        W[HALF * 32 + 4 * j + 1] = t.y;
        W[HALF * 32 + 4 * j + 2] = t.z;
        W[HALF * 32 + 4 * j + 3] = t.w;
    }
}

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1,
                                            uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
                 "[%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(smem_dst)),
                 "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}

// Tile k produces outputs [k*MT, (k+1)*MT).  Plane element 0 is x_p[k*MT - TQ] (one sample before
// the oldest sample the taps reach), so every thread's window and every thread's 128 B of outputs
// start on a 128-byte row: interior tiles are staged by ONE TMA tensor load (D == 1) and written
// back by ONE TMA tensor store.
// smem: [mbarrier 16 B][taps D*TQ floats][pad to 1024 B][D planes of plane_rows*32 floats]
// DECIM = false: D == 1 instantiation (TMA for interior tiles, register-staged loads for the few
// edge tiles).  DECIM = true: decimating filters -- phase de-interleave with cp.async.  (Keeping
// the cp.async path out of the D == 1 kernel is worth ~10 % on short filters: measured A/B.)
// ---- decimation by ANY small D folded into the full-rate kernel (DG) --------------------------------
// A thread owns D consecutive 16-sample rows of the input tile = exactly 16 outputs (one output row).
// Row c of the thread starts at input 16 (D t + c), so its outputs sit at the compile-time positions
// p0(c), p0(c) + D, ... with p0(c) = (-16 c) mod D, and they are the thread's outputs
// off(c) .. off(c) + n(c) - 1 with off(c) = ceil(16 c / D): D passes of the register-window loop, each
// over its own row window, accumulate disjoint slices of the SAME 16 accumulators, and the epilogue
// (row per thread, one TMA store) is the full-rate one.  Same products as the phase-plane kernel, no
// de-interleaving copies; the window loads per input sample equal those of the full-rate filter.
// threads per tile: D rows of 128 B per thread, so these kernels run 64-thread tiles (32 threads from D = 9;
// 24-60 KB, 3-8 CTAs per SM; with 128 threads D = 7 fits one CTA per SM: 64 taps 338 -> 492 GS/s, D = 5 471 -> 558, D = 3 476 -> 494)
#ifndef B200_FIR_DG_SMALL
#define B200_FIR_DG_SMALL 3
#endif
__host__ __device__ constexpr int fir_dg_nt(int dg) { return dg >= 9 ? FIR_NT / 4 : dg >= B200_FIR_DG_SMALL ? FIR_NT / 2 : FIR_NT; }
// Even D: the rows D t + k of the 8 lanes that share a shared-memory wavefront take only 8 / gcd(D, 8)
// different values of (row & 7), i.e. of the 128-byte swizzle: a 2-way (D = 6) bank conflict on every
// window load.  One unused row after every P = 8 / gcd(D, 8) threads' rows (P D rows, one TMA box each)
// makes lanes t and t + P differ by 1 (mod 8): fir_dg_group = rows per group, 0 = no padding.
#ifndef B200_FIR_DG_PAD
#define B200_FIR_DG_PAD 1
#endif
#ifndef B200_FIR_DG_PAD_REAL
#define B200_FIR_DG_PAD_REAL 9 // real streams: pad multiples of 4 from this D on (i.e. D = 12)
#endif
// (Real streams: only D = 12, where the conflict is 4-way -- 32 taps 671 -> 835 GS/s; the 2-way cases are
// FMA-bound in the scalar loop and the padding costs them a resident CTA per SM: D = 6 at 512 taps
// 194 -> 168, D = 10 at 32 taps 1017 -> 870.)
__host__ __device__ constexpr int fir_dg_group(int dg, int vec)
{
    return (!B200_FIR_DG_PAD || (vec != 2 && (dg < B200_FIR_DG_PAD_REAL || dg % 4)) || dg < 2 || dg % 2)
               ? 0
               : (dg % 8 == 0 ? 1 : dg % 4 == 0 ? 2 : 4) * dg;
}
template <int GR>
__device__ __forceinline__ int fir_prow(int r) { return GR ? r + r / GR : r; }

template <int DG, int C, int R = 16> // R = samples per row: 16 complex, 32 real
struct fir_dg {
    static constexpr int P0 = (DG - (R * C) % DG) % DG;
    static constexpr int OFF_OUT = (R * C + DG - 1) / DG;
    static constexpr int N = (R - P0 + DG - 1) / DG;
};

template <int VEC, int OFF, int DG, int C>
__device__ __forceinline__ void fir_step_dg(float (&acc)[FIR_ACC], const float (&W)[FIR_RING],
                                            const float* __restrict__ hs)
{
    constexpr int CH = FIR_ACC / VEC;
    using G = fir_dg<DG, C, CH>;
#pragma unroll
    for (int q4 = 0; q4 < CH; q4 += 4) {
        float4 h4 = *reinterpret_cast<const float4*>(hs + q4);
        const float hv[4] = { h4.x, h4.y, h4.z, h4.w };
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const float2 h2 = make_float2(hv[u], hv[u]);
#pragma unroll
            for (int j = 0; j < G::N; j++) {
                const int i = (OFF + (q4 + u + 1 + G::P0 + j * DG) * VEC) % FIR_RING;
                const int l = (G::OFF_OUT + j) * VEC;
                if (VEC == 2) {
                    float2 a = __ffma2_rn(make_float2(W[i], W[i + 1]), h2, make_float2(acc[l], acc[l + 1]));
                    acc[l] = a.x;
                    acc[l + 1] = a.y;
                } else
                    acc[l] = fmaf(hv[u], W[i], acc[l]);
            }
        }
    }
}

template <int HALF>
__device__ __forceinline__ void fir_load_half(float (&W)[FIR_RING], const float* __restrict__ plane, int row);

template <int VEC, int DG, int C = 0>
__device__ __forceinline__ void fir_passes_dg(float (&acc)[FIR_ACC], float (&W)[FIR_RING],
                                              const float* __restrict__ plane, const float* __restrict__ hp,
                                              int nsteps, int tid)
{
    if constexpr (C < DG) {
        constexpr int CH = FIR_ACC / VEC;
        constexpr int GR = fir_dg_group(DG, VEC);
        const int r0 = tid * DG + C;
        fir_load_half<0>(W, plane, fir_prow<GR>(r0));
        int b = 0;
        for (; b + 1 < nsteps; b += 2) {
            fir_load_half<1>(W, plane, fir_prow<GR>(r0 + b + 1));
            fir_step_dg<VEC, 0, DG, C>(acc, W, hp + b * CH);
            fir_load_half<0>(W, plane, fir_prow<GR>(r0 + b + 2));
            fir_step_dg<VEC, 32, DG, C>(acc, W, hp + (b + 1) * CH);
        }
        if (b < nsteps) {
            fir_load_half<1>(W, plane, fir_prow<GR>(r0 + b + 1));
            fir_step_dg<VEC, 0, DG, C>(acc, W, hp + b * CH);
        }
        fir_passes_dg<VEC, DG, C + 1>(acc, W, plane, hp, nsteps, tid);
    }
}

// DD > 1 (with DECIM = false): decimation by a divisor of the 16 (32) window positions of a thread.
// The tile is the SAME 2048-sample (4096 for fff) input tile as for D = 1, staged by the same single
// TMA tensor load with the taps in natural order; a thread simply keeps accumulators only for the
// positions 0, DD, 2 DD ... of its row, i.e. 16/DD outputs, and the output tile shrinks to 128/DD rows.
// No phase planes, no per-sample de-interleaving copies: a short decimating filter becomes HBM-bound
// like a short full-rate one (64 taps, decimation 4: 197 -> 400+ GS/s input rate).
// LL > 1 (with DECIM = false, DD = 1): interpolation by LL folded into the same kernel.  Output
// phase r of an interpolator, y[n LL + r] = sum_q h[q LL + r] x[n - q], is a full-rate filter over the
// SAME input tile with the taps of phase r: LL passes of the register-blocked loop over one
// TMA-staged tile, each scattering its 16 results per thread into the thread's own LL rows of a
// separate output tile (16 LL consecutive outputs), which then leaves by LL TMA tensor stores.
// RP ("real pairs", with VEC = 2, DECIM = false): a REAL stream f[] run through the packed
// complex x real loop.  With the float pairs P0[j] = (f[2j], f[2j+1]) (the stream itself) and
// P1[j] = (f[2j-1], f[2j]) (the stream one float later), (y[2m], y[2m+1]) = h_e * P0 + h_o * P1 with the
// even / odd taps: two passes of the FFMA2 loop, i.e. half the issue slots of the scalar fff loop.
// P0 is staged by the same TMA tensor load as a complex stream; P1 is derived from it in shared memory
// (each thread shifts its own row by one float).
template <int VEC, bool DECIM, int DD = 1, int LL = 1, bool RP = false, int DG = 1>
__global__ void __launch_bounds__(FIR_NT, FIR_MINB)
    fir_direct_kernel(const float* __restrict__ x, const float* __restrict__ hist,
                      float* __restrict__ y, const float* __restrict__ taps_pp,
                      const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_out,
                      fir_geom gm, fir_epilogue ep)
{
    constexpr int NT = fir_dg_nt(DG);  // threads (= output rows) per tile
    constexpr int R = FIR_ACC / VEC;  // window positions per thread (= outputs per thread for DD == 1)
    constexpr int CH = FIR_ACC / VEC; // taps per step
    constexpr int MT = NT * R * DG; // input-rate positions per tile
    constexpr int MTO = MT / DD / DG;   // outputs per tile
    static_assert(DG == 1 || (!DECIM && DD == 1 && LL == 1 && !RP), "DG is a mode of its own");
    static_assert(!DECIM || DD == 1, "DD applies to the TMA-staged full-rate kernel only");
    static_assert(R % DD == 0, "decimation must divide the positions per thread");
    static_assert(LL == 1 || (!DECIM && DD == 1), "LL applies to the TMA-staged full-rate kernel only");
    static_assert(!RP || (VEC == 2 && !DECIM && DD == 1 && LL == 1), "RP is the full-rate real-stream mode");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    float* hs = reinterpret_cast<float*>(smem_raw + 16);
    const int D = DECIM ? gm.D : 1, TQ = gm.TQ;
    const int NTAPROWS = RP ? 2 : LL > 1 ? LL : D; // tap rows: one per decimation / interpolation phase
    float* planes;
    {
        uint32_t a = smem_u32(hs + NTAPROWS * TQ);
        uint32_t aligned = (a + 1023u) & ~1023u;
        planes = hs + NTAPROWS * TQ + (aligned - a) / 4;
    }
    // phase planes are skewed by 32 B each so that the de-interleaving stores of one warp (same
    // element, different phase) land in different banks
    const int plane_f = (gm.plane_rows << 5) + (DECIM ? 8 : 0);
    const int tid = threadIdx.x;
    const long long tile = blockIdx.x;
    const long long B0 = tile * MT - TQ; // x_p index of plane element 0
    const long long O0 = tile * MTO;     // first output of this tile
    constexpr int GR = fir_dg_group(DG, VEC);              // DG, even D: one pad row per GR rows
    const int PLs = ((gm.box_rows * gm.n_boxes) << 5) / VEC; // samples per plane

    // interior tile of a D == 1 filter: one TMA tensor copy stages the whole window
    const long long row0 = B0 * VEC / 32;
    const bool use_tma = !DECIM && gm.tma_ok && B0 >= 0 &&
                         row0 + (long long)gm.box_rows * gm.n_boxes <= gm.full_rows;
    if (use_tma && tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        mbar_arrive_expect_tx(bar, (uint32_t)(gm.box_rows * gm.n_boxes) * 128u);
        for (int bx = 0; bx < gm.n_boxes; bx++)
            tma_load_2d(planes + (size_t)bx * (gm.box_rows + (GR ? 1 : 0)) * 32, &tmap, 0, (int)(row0 + (long long)bx * gm.box_rows),
                        bar);
    }
    for (int i = tid; i < NTAPROWS * TQ; i += NT)
        hs[i] = __ldg(taps_pp + i);
    if (DECIM) {
        // decimating filters: every sample inside the input goes global -> shared with cp.async
        // (LDGSTS: asynchronous, no register staging, all of a thread's copies in flight at once)
        // and is de-interleaved by phase on the way in; samples before the stream start come from
        // the history buffer
        const long long g_lo = B0 * D - (D - 1);
        const int total = PLs * D;
        for (int i = tid; i < total; i += NT) {
            const int e = i / D;
            const int p = D - 1 - (i - e * D);
            float* dst = planes + p * plane_f + swz(e * VEC);
            const long long g = g_lo + i;
            if (g >= 0 && g < gm.n_in) {
                if (VEC == 2)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)),
                                 "l"(x + g * 2)
                                 : "memory");
                else
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(x + g)
                                 : "memory");
            } else {
                float v[2];
                fir_fetch<VEC>(x, hist, gm.Tm1, g, gm.n_in, v);
                if (VEC == 2)
                    *reinterpret_cast<float2*>(dst) = make_float2(v[0], v[1]);
                else
                    dst[0] = v[0];
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else if (!use_tma) {
        // D == 1 edge tiles / unaligned input: coalesced loads, 8 independent loads in flight
        const long long g_lo = B0;
        const int total = PLs;
        for (int i0 = tid; i0 < total; i0 += NT * 8) {
            float v[8][2];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int i = i0 + u * NT;
                if (i < total) {
                    if (RP) { // pair g = floats 2g, 2g+1 of the real stream, each with its own bounds
                        float t0[2], t1_[2];
                        fir_fetch<1>(x, hist, gm.Tm1, 2 * (g_lo + i), gm.n_in_f, t0);
                        fir_fetch<1>(x, hist, gm.Tm1, 2 * (g_lo + i) + 1, gm.n_in_f, t1_);
                        v[u][0] = t0[0];
                        v[u][1] = t1_[0];
                    } else
                        fir_fetch<VEC>(x, hist, gm.Tm1, g_lo + i, gm.n_in, v[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int i = i0 + u * NT;
                if (i < total) {
                    const int fl = i * VEC;
                    float* dst = planes + swz(GR ? (fir_prow<GR>(fl >> 5) << 5) | (fl & 31) : fl);
                    if (VEC == 2)
                        *reinterpret_cast<float2*>(dst) = make_float2(v[u][0], v[u][1]);
                    else
                        dst[0] = v[u][0];
                }
            }
        }
    }
    __syncthreads();
    if (use_tma)
        mbar_wait(bar, 0);
    if (RP) {
        // plane 1 = plane 0 one float later: row r = [last float of row r-1, first 31 floats of row r]
        float first[2];
        fir_fetch<1>(x, hist, gm.Tm1, 2 * B0 - 1, gm.n_in_f, first); // the float in front of the tile
        float* P1 = planes + plane_f;
        for (int r = tid; r < gm.plane_rows; r += NT) {
            const float* rb = planes + (r << 5);
            const int sw = (r & 7) << 2;
            float a[32];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const float4 t = *reinterpret_cast<const float4*>(rb + ((j << 2) ^ sw));
                a[4 * j] = t.x, a[4 * j + 1] = t.y, a[4 * j + 2] = t.z, a[4 * j + 3] = t.w;
            }
            const float prev = r > 0 ? planes[swz(32 * r - 1)] : first[0];
            float* wb = P1 + (r << 5);
            *reinterpret_cast<float4*>(wb + (0 ^ sw)) = make_float4(prev, a[0], a[1], a[2]);
#pragma unroll
            for (int j = 1; j < 8; j++)
                *reinterpret_cast<float4*>(wb + ((j << 2) ^ sw)) =
                    make_float4(a[4 * j - 1], a[4 * j], a[4 * j + 1], a[4 * j + 2]);
        }
        __syncthreads();
    }

    // ---- register-blocked multiply-accumulate -----------------------------------------
    float acc[FIR_ACC];
#pragma unroll
    for (int l = 0; l < FIR_ACC; l++)
        acc[l] = 0.f;
    float W[FIR_RING];
    const int nsteps = TQ / CH; // any count >= 1: an odd tail step runs alone
    if (LL > 1) {
        // output tile: LL * NT rows behind the input plane (1024-byte aligned for the TMA stores)
        float* otile = planes + (((gm.plane_rows << 5) + 255) & ~255);
#pragma unroll 1
        for (int r = 0; r < LL; r++) {
#pragma unroll
            for (int l = 0; l < FIR_ACC; l++)
                acc[l] = 0.f;
            const float* hp = hs + r * TQ;
            fir_load_half<0>(W, planes, tid);
            int b = 0;
            for (; b + 1 < nsteps; b += 2) {
                fir_load_half<1>(W, planes, tid + b + 1);
                fir_step<VEC, CH, 0, 1>(acc, W, hp + b * CH);
                fir_load_half<0>(W, planes, tid + b + 2);
                fir_step<VEC, CH, 32, 1>(acc, W, hp + (b + 1) * CH);
            }
            if (b < nsteps) {
                fir_load_half<1>(W, planes, tid + b + 1);
                fir_step<VEC, CH, 0, 1>(acc, W, hp + b * CH);
            }
            // position p of the thread is output (tid R + p) LL + r of the tile
#pragma unroll
            for (int pz = 0; pz < R; pz++) {
                const int f = ((tid * R + pz) * LL + r) * VEC;
                if (VEC == 2)
                    *reinterpret_cast<float2*>(otile + swz(f)) = make_float2(acc[2 * pz], acc[2 * pz + 1]);
                else
                    otile[swz(f)] = acc[pz];
            }
        }
        const long long orow0 = tile * (NT * LL);
        if (gm.tma_out_ok && orow0 + NT * LL <= gm.full_out_rows) {
            fence_proxy_async();
            __syncthreads();
            if (tid == 0) {
#pragma unroll 1
                for (int r = 0; r < LL; r++)
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                                     &tmap_out),
                                 "r"(0), "r"((int)(orow0 + (long long)r * NT)),
                                 "r"(smem_u32(otile + (size_t)r * NT * 32))
                                 : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            return;
        }
        __syncthreads();
        const long long OL = tile * (long long)(MT * LL);
#pragma unroll 4
        for (int i = tid; i < MT * LL; i += NT) {
            const long long m = OL + i;
            if (m >= gm.n_out)
                break;
            const float* src = otile + swz(i * VEC);
            if (VEC == 2)
                __stcs(reinterpret_cast<float2*>(y) + m, *reinterpret_cast<const float2*>(src));
            else
                __stcs(y + m, src[0]);
        }
        return;
    }
    if constexpr (DG > 1) {
        fir_passes_dg<VEC, DG>(acc, W, planes, hs, nsteps, tid);
    } else
    for (int p = 0; p < (RP ? 2 : D); p++) {
        const float* plane = planes + p * plane_f;
        const float* hp = hs + p * TQ;
        fir_load_half<0>(W, plane, tid);
        int b = 0;
        for (; b + 1 < nsteps; b += 2) { // pairs of steps: the loop body the scheduler pipelines
            fir_load_half<1>(W, plane, tid + b + 1);
            fir_step<VEC, CH, 0, DD>(acc, W, hp + b * CH);
            fir_load_half<0>(W, plane, tid + b + 2);
            fir_step<VEC, CH, 32, DD>(acc, W, hp + (b + 1) * CH);
        }
        if (b < nsteps) { // odd tail step (ring half 0 holds row tid + b)
            fir_load_half<1>(W, plane, tid + b + 1);
            fir_step<VEC, CH, 0, DD>(acc, W, hp + b * CH);
        }
    }
    __syncthreads();

    // ---- outputs: registers -> shared (swizzled row per thread) -> global ----------------------
    if (ep.fuse && RP) {
#pragma unroll
        for (int l = 0; l < FIR_ACC; l++)
            acc[l] = __fmul_rn(acc[l], ep.kre);
    } else if (ep.fuse) {
        if (VEC == 2) {
#pragma unroll
            for (int l = 0; l < FIR_ACC; l += 2 * DD) {
                float2 v = cmul_nofma(make_float2(acc[l], acc[l + 1]), ep.kre, ep.kim);
                acc[l] = v.x;
                acc[l + 1] = v.y;
            }
        } else {
#pragma unroll
            for (int l = 0; l < FIR_ACC; l += DD)
                acc[l] = __fmul_rn(acc[l], ep.kre);
        }
    }
    if (DD == 1) {
        float* rb = planes + (tid << 5);
        const int s = (tid & 7) << 2;
#pragma unroll
        for (int j = 0; j < 8; j++)
            *reinterpret_cast<float4*>(rb + ((j << 2) ^ s)) =
                make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
    } else {
        // the thread's FIR_ACC/DD output floats, compacted, at their place in the (swizzled) output tile
        constexpr int NF = FIR_ACC / DD; // floats per thread
        float o[NF];
#pragma unroll
        for (int j = 0; j < NF; j++)
            o[j] = acc[(j / VEC) * DD * VEC + (j % VEC)];
        const int f0 = tid * NF;
        if (NF >= 4) {
#pragma unroll
            for (int j = 0; j < NF; j += 4)
                *reinterpret_cast<float4*>(planes + swz(f0 + j)) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < NF; j++)
                planes[swz(f0 + j)] = o[j];
        }
    }
    const long long orow0 = tile * (NT / DD); // output rows per tile
    if (NT / DD >= 8 && gm.tma_out_ok && orow0 + NT / DD <= gm.full_out_rows) {
        // whole tile inside the output: one TMA tensor store from the swizzled rows
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
#pragma unroll 4
    for (int i = tid; i < MTO; i += NT) {
        const long long m = O0 + i;
        if (m >= gm.n_out)
            break;
        const float* src = planes + swz(i * VEC);
        if (RP) { // the last pair of an odd-length real stream is half a pair
            __stcs(y + 2 * m, src[0]);
            if (2 * m + 1 < gm.n_out_f)
                __stcs(y + 2 * m + 1, src[1]);
        } else if (VEC == 2)
            __stcs(reinterpret_cast<float2*>(y) + m, *reinterpret_cast<const float2*>(src));
        else
            __stcs(y + m, src[0]);
    }
}

// Fallback for parameter combinations the tiled kernel cannot stage (very large D):
// one thread per output straight from global memory.
template <int VEC>
__global__ void __launch_bounds__(256)
    fir_naive_kernel(const float* __restrict__ x, const float* __restrict__ hist,
                     float* __restrict__ y, const float* __restrict__ taps, int T, int D,
                     long long n_in, long long n_out, fir_epilogue ep)
{
    long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_out)
        return;
    float a0 = 0.f, a1 = 0.f;
    for (int k = 0; k < T; k++) {
        float v[2];
        fir_fetch<VEC>(x, hist, T - 1, m * D - k, n_in, v);
        float h = __ldg(taps + k);
        a0 = fmaf(h, v[0], a0);
        if (VEC == 2)
            a1 = fmaf(h, v[1], a1);
    }
    if (VEC == 2) {
        float2 v = make_float2(a0, a1);
        if (ep.fuse)
            v = cmul_nofma(v, ep.kre, ep.kim);
        reinterpret_cast<float2*>(y)[m] = v;
    } else {
        y[m] = ep.fuse ? __fmul_rn(a0, ep.kre) : a0;
    }
}

// new_hist[j] = sample (n_consumed - Tm1 + j) of the stream (old history ++ x)
__global__ void fir_hist_kernel(const float* __restrict__ x, const float* __restrict__ hist_old,
                                float* __restrict__ hist_new, long long n_consumed, int Tm1, int VEC)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Tm1)
        return;
    long long s = n_consumed - Tm1 + j;
    for (int c = 0; c < VEC; c++)
        hist_new[j * VEC + c] = (s >= 0) ? x[s * VEC + c] : hist_old[(Tm1 + s) * VEC + c];
}

} // namespace b200

using namespace b200;

struct b200_fir {
    int T = 0, D = 1, vec = 2;
    int TQ = 0;      // taps per phase, padded
    int plane_rows = 0, box_rows = 0, n_boxes = 0; // smem plane geometry (rows of 128 B)
    size_t smem = 0;
    int use_tma = 1;
    int rp = 0;      // real stream through the packed complex x real loop (fir_direct_kernel<..., RP>)
    int interp = 0;  // > 1: interpolation folded into the full-rate kernel (created by fir_interp_create)
    int dg = 0;      // > 1: decimation by a non-divisor of 16 folded into the full-rate kernel (D rows per thread)
    int dd = 0;      // > 1: decimation folded into the TMA-staged full-rate kernel (geometry as for D = 1)
    ols_plan* ols = nullptr; // algorithm 3
    ffa_plan* ffa = nullptr; // algorithm 5
    int algorithm = 1;
    fir_epilogue ep{ 0, 1.f, 0.f };
    float* d_taps_pp = nullptr; // [D][TQ] reversed per phase
    float* d_taps = nullptr;    // [T] natural order (naive kernel)
    float* d_hist[2] = { nullptr, nullptr };
    int cur = 0;
    int device = 0;
};

// 2-D view of the input stream for TMA: rows of 32 floats (128 B), SWIZZLE_128B, box = box_rows
static int fir_make_tmap(CUtensorMap* tmap, const void* d_in, long long full_rows, int box_rows)
{
    tmap_encode_fn enc = tmap_encode_tiled();
    if (!enc)
        return set_err(B200_ERR_CUDA, "fir: cuTensorMapEncodeTiled unavailable");
    cuuint64_t gdim[2] = { 32, (cuuint64_t)full_rows };
    cuuint64_t gstride[1] = { 128 };
    cuuint32_t box[2] = { 32, (cuuint32_t)box_rows };
    cuuint32_t estr[2] = { 1, 1 };
    CUresult r = enc(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(d_in), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_err(B200_ERR_CUDA, "fir: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return B200_OK;
}

static int fir_launch(b200_fir* h, const float* d_hist, const void* d_in, void* d_out,
                      long long n_in, long long n_out, cudaStream_t s)
{
    if (n_out <= 0)
        return B200_OK;
    const float* x = (const float*)d_in;
    float* y = (float*)d_out;
    if (h->algorithm == 3)
        return ols_launch(h->ols, d_hist, d_in, d_out, n_in, n_out, s);
    if (h->algorithm == 5)
        return ffa_launch(h->ffa, d_hist, d_in, d_out, n_in, n_out, s);
    if (h->algorithm == 1 && h->rp) {
        // real stream as float pairs: tiles of 2048 pairs = 4096 real outputs
        const long long pairs_out = (n_out + 1) / 2, pairs_in = (n_in + 1) / 2;
        const int MTp = FIR_NT * (FIR_ACC / 2);
        const long long tiles = (pairs_out + MTp - 1) / MTp;
        if (tiles > 0x7fffffffLL)
            return set_err(B200_ERR_ARG, "fir: too many items for one call");
        fir_geom gm{};
        gm.Tm1 = h->T - 1;
        gm.D = 1;
        gm.TQ = h->TQ;
        gm.plane_rows = h->plane_rows;
        gm.box_rows = h->box_rows;
        gm.n_boxes = h->n_boxes;
        gm.n_in = pairs_in;
        gm.n_out = pairs_out;
        gm.n_in_f = n_in;
        gm.n_out_f = n_out;
        gm.full_rows = n_in / 32;
        gm.full_out_rows = n_out / 32;
        CUtensorMap tmap, tmap_out;
        memset(&tmap, 0, sizeof(tmap));
        memset(&tmap_out, 0, sizeof(tmap_out));
        if (h->use_tma && gm.full_rows >= h->plane_rows && (uintptr_t)d_in % 16 == 0 &&
            fir_make_tmap(&tmap, d_in, gm.full_rows, h->box_rows) == B200_OK)
            gm.tma_ok = 1;
        if (h->use_tma && gm.full_out_rows >= FIR_NT && (uintptr_t)d_out % 16 == 0 &&
            fir_make_tmap(&tmap_out, d_out, gm.full_out_rows, FIR_NT) == B200_OK)
            gm.tma_out_ok = 1;
        B200_LAUNCH((fir_direct_kernel<2, false, 1, 1, true>), (unsigned)tiles, FIR_NT, h->smem, s, x, d_hist, y,
                    h->d_taps_pp, tmap, tmap_out, gm, h->ep);
        return B200_OK;
    }
    if (h->algorithm == 1) {
        const int NTt = h->dg ? fir_dg_nt(h->dg) : FIR_NT;             // threads per tile
        const int MT = NTt * (FIR_ACC / h->vec) / (h->dd ? h->dd : 1); // outputs per tile
        long long tiles = (n_out + MT - 1) / MT;
        if (tiles > 0x7fffffffLL)
            return set_err(B200_ERR_ARG, "fir: too many items for one call");
        fir_geom gm{};
        gm.Tm1 = h->T - 1;
        gm.D = (h->dd || h->dg) ? 1 : h->D;
        gm.TQ = h->TQ;
        gm.plane_rows = h->plane_rows;
        gm.box_rows = h->box_rows;
        gm.n_boxes = h->n_boxes;
        gm.n_in = n_in;
        gm.n_out = n_out;
        gm.full_rows = n_in * h->vec / 32;
        gm.tma_ok = 0;
        CUtensorMap tmap;
        memset(&tmap, 0, sizeof(tmap));
        if (h->use_tma && (h->D == 1 || h->dd || h->dg) && gm.full_rows >= h->plane_rows && (uintptr_t)d_in % 16 == 0) {
            int rc = fir_make_tmap(&tmap, d_in, gm.full_rows, h->box_rows);
            if (rc != B200_OK)
                return rc;
            gm.tma_ok = 1;
        }
        CUtensorMap tmap_out;
        memset(&tmap_out, 0, sizeof(tmap_out));
        gm.full_out_rows = n_out * h->vec / 32;
        gm.tma_out_ok = 0;
        const int out_rows = NTt / (h->dd ? h->dd : 1); // rows of the output tile
        if (h->use_tma && out_rows >= 8 && gm.full_out_rows >= out_rows && (uintptr_t)d_out % 16 == 0) {
            int rc = fir_make_tmap(&tmap_out, d_out, gm.full_out_rows, out_rows);
            if (rc != B200_OK)
                return rc;
            gm.tma_out_ok = 1;
        }
        const bool decim = h->D > 1;
        if (h->dg) {
#define FIR_DG(V, DGV)                                                                                       \
    B200_LAUNCH((fir_direct_kernel<V, false, 1, 1, false, DGV>), (unsigned)tiles, NTt, h->smem, s, x, d_hist, \
                y, h->d_taps_pp, tmap, tmap_out, gm, h->ep)
#define FIR_DG_ALL(V)                                                                                        \
    switch (h->dg) {                                                                                         \
    case 3: FIR_DG(V, 3); break;                                                                             \
    case 5: FIR_DG(V, 5); break;                                                                             \
    case 6: FIR_DG(V, 6); break;                                                                             \
    case 7: FIR_DG(V, 7); break;                                                                             \
    case 9: FIR_DG(V, 9); break;                                                                             \
    case 10: FIR_DG(V, 10); break;                                                                           \
    case 11: FIR_DG(V, 11); break;                                                                           \
    case 12: FIR_DG(V, 12); break;                                                                           \
    case 13: FIR_DG(V, 13); break;                                                                           \
    case 14: FIR_DG(V, 14); break;                                                                           \
    default: FIR_DG(V, 15); break;                                                                           \
    }
            if (h->vec == 2) {
                FIR_DG_ALL(2);
            } else {
                FIR_DG_ALL(1);
            }
#undef FIR_DG_ALL
#undef FIR_DG
        } else if (h->dd) {
#define FIR_DD(V, DDV)                                                                                     \
    B200_LAUNCH((fir_direct_kernel<V, false, DDV>), (unsigned)tiles, FIR_NT, h->smem, s, x, d_hist, y,     \
                h->d_taps_pp, tmap, tmap_out, gm, h->ep)
            if (h->vec == 2) {
                switch (h->dd) {
                case 2: FIR_DD(2, 2); break;
                case 4: FIR_DD(2, 4); break;
                case 8: FIR_DD(2, 8); break;
                default: FIR_DD(2, 16); break;
                }
            } else {
                switch (h->dd) {
                case 2: FIR_DD(1, 2); break;
                case 4: FIR_DD(1, 4); break;
                case 8: FIR_DD(1, 8); break;
                case 16: FIR_DD(1, 16); break;
                default: FIR_DD(1, 32); break;
                }
            }
#undef FIR_DD
        } else if (h->vec == 2) {
            if (decim)
                B200_LAUNCH((fir_direct_kernel<2, true>), (unsigned)tiles, FIR_NT, h->smem, s, x, d_hist, y,
                            h->d_taps_pp, tmap, tmap_out, gm, h->ep);
            else
                B200_LAUNCH((fir_direct_kernel<2, false>), (unsigned)tiles, FIR_NT, h->smem, s, x, d_hist, y,
                            h->d_taps_pp, tmap, tmap_out, gm, h->ep);
        } else {
            if (decim)
                B200_LAUNCH((fir_direct_kernel<1, true>), (unsigned)tiles, FIR_NT, h->smem, s, x, d_hist, y,
                            h->d_taps_pp, tmap, tmap_out, gm, h->ep);
            else
                B200_LAUNCH((fir_direct_kernel<1, false>), (unsigned)tiles, FIR_NT, h->smem, s, x, d_hist, y,
                            h->d_taps_pp, tmap, tmap_out, gm, h->ep);
        }
    } else {
        long long blocks = (n_out + 255) / 256;
        if (h->vec == 2)
            B200_LAUNCH((fir_naive_kernel<2>), (unsigned)blocks, 256, 0, s, x, d_hist, y, h->d_taps,
                        h->T, h->D, n_in, n_out, h->ep);
        else
            B200_LAUNCH((fir_naive_kernel<1>), (unsigned)blocks, 256, 0, s, x, d_hist, y, h->d_taps,
                        h->T, h->D, n_in, n_out, h->ep);
    }
    return B200_OK;
}

namespace b200 {

bool fir_interp_supported(int T, int L, int is_complex)
{
    if (L < 2 || L > 4 || T < 1 || getenv("B200_INTERP_RB"))
        return false;
    // measured against the register-blocked resampler kernel (tools/resampler_sweep.py): +22..47 % for
    // L = 2, 3 at any length; at L = 4 the 64 KiB output tile leaves 2 CTAs per SM and short phases lose
    // (128 taps: 230 vs 259 GS/s out), so L = 4 folds only once the phases are FMA-bound
    if (L == 4 && T < 192)
        return false;
    const int vec = is_complex ? 2 : 1, CH = FIR_ACC / vec;
    const int TQ = ((T + L - 1) / L + CH - 1) / CH * CH;
    return TQ / CH <= 64;
}

int fir_interp_create(const float* taps, int T, int L, int is_complex, b200_fir** out)
{
    *out = nullptr;
    if (!fir_interp_supported(T, L, is_complex))
        return set_err(B200_ERR_UNSUPPORTED, "fir interpolation fold: L = %d, %d taps not supported", L, T);
    b200_fir* h = new b200_fir();
    h->T = T;
    h->D = 1;
    h->interp = L;
    h->vec = is_complex ? 2 : 1;
    const int CH = FIR_ACC / h->vec;
    h->TQ = ((T + L - 1) / L + CH - 1) / CH * CH;
    const int need = FIR_NT + h->TQ / CH;
    h->n_boxes = (need + 255) / 256;
    h->box_rows = (need + h->n_boxes - 1) / h->n_boxes;
    h->plane_rows = h->box_rows * h->n_boxes;
    h->smem = 16 + sizeof(float) * ((size_t)L * h->TQ + (size_t)h->plane_rows * 32 + 256) + 2048 +
              (size_t)L * FIR_NT * 128;
    std::vector<float> pp((size_t)L * h->TQ, 0.f);
    for (int ph = 0; ph < L; ph++)
        for (int qr = 0; qr < h->TQ; qr++) {
            const long long k = (long long)(h->TQ - 1 - qr) * L + ph;
            pp[(size_t)ph * h->TQ + qr] = k < T ? taps[k] : 0.f;
        }
    cudaError_t e = cudaMalloc(&h->d_taps_pp, pp.size() * sizeof(float));
    if (e == cudaSuccess)
        e = cudaMemcpy(h->d_taps_pp, pp.data(), pp.size() * sizeof(float), cudaMemcpyHostToDevice);
#define FIR_LL_ATTR(V, LLV)                                                                                  \
    if (e == cudaSuccess)                                                                                    \
    e = cudaFuncSetAttribute(fir_direct_kernel<V, false, 1, LLV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                             227 * 1024)
    FIR_LL_ATTR(2, 2);
    FIR_LL_ATTR(2, 3);
    FIR_LL_ATTR(2, 4);
    FIR_LL_ATTR(1, 2);
    FIR_LL_ATTR(1, 3);
    FIR_LL_ATTR(1, 4);
#undef FIR_LL_ATTR
    if (e != cudaSuccess) {
        b200_fir_destroy(h);
        return set_err(B200_ERR_CUDA, "fir interpolation fold: %s", cudaGetErrorString(e));
    }
    *out = h;
    return B200_OK;
}

int fir_interp_launch(b200_fir* h, const float* d_hist, const void* d_in, void* d_out, long long n_in,
                      cudaStream_t s)
{
    if (n_in <= 0)
        return B200_OK;
    const int L = h->interp;
    const int MT = FIR_NT * (FIR_ACC / h->vec);
    const long long tiles = (n_in + MT - 1) / MT;
    if (tiles > 0x7fffffffLL)
        return set_err(B200_ERR_ARG, "resampler: too many items for one call");
    fir_geom gm{};
    gm.Tm1 = (h->T + L - 1) / L - 1; // history is counted in input samples of a phase filter
    gm.D = 1;
    gm.TQ = h->TQ;
    gm.plane_rows = h->plane_rows;
    gm.box_rows = h->box_rows;
    gm.n_boxes = h->n_boxes;
    gm.n_in = n_in;
    gm.n_out = n_in * L;
    gm.full_rows = n_in * h->vec / 32;
    gm.full_out_rows = gm.n_out * h->vec / 32;
    CUtensorMap tmap, tmap_out;
    memset(&tmap, 0, sizeof(tmap));
    memset(&tmap_out, 0, sizeof(tmap_out));
    if (h->use_tma && gm.full_rows >= h->plane_rows && (uintptr_t)d_in % 16 == 0 &&
        fir_make_tmap(&tmap, d_in, gm.full_rows, h->box_rows) == B200_OK)
        gm.tma_ok = 1;
    if (h->use_tma && gm.full_out_rows >= FIR_NT && (uintptr_t)d_out % 16 == 0 &&
        fir_make_tmap(&tmap_out, d_out, gm.full_out_rows, FIR_NT) == B200_OK)
        gm.tma_out_ok = 1;
    const float* x = (const float*)d_in;
    float* y = (float*)d_out;
    fir_epilogue ep{ 0, 1.f, 0.f };
#define FIR_LL(V, LLV)                                                                                        \
    B200_LAUNCH((fir_direct_kernel<V, false, 1, LLV>), (unsigned)tiles, FIR_NT, h->smem, s, x, d_hist, y,     \
                h->d_taps_pp, tmap, tmap_out, gm, ep)
    if (h->vec == 2) {
        switch (L) {
        case 2: FIR_LL(2, 2); break;
        case 3: FIR_LL(2, 3); break;
        default: FIR_LL(2, 4); break;
        }
    } else {
        switch (L) {
        case 2: FIR_LL(1, 2); break;
        case 3: FIR_LL(1, 3); break;
        default: FIR_LL(1, 4); break;
        }
    }
#undef FIR_LL
    return B200_OK;
}

} // namespace b200

extern "C" {

int b200_fir_destroy(b200_fir* h)
{
    if (!h)
        return B200_OK;
    cudaFree(h->d_taps_pp);
    cudaFree(h->d_taps);
    cudaFree(h->d_hist[0]);
    cudaFree(h->d_hist[1]);
    ols_destroy(h->ols);
    ffa_destroy(h->ffa);
    delete h;
    return B200_OK;
}

int b200_fir_create(const b200_fir_params* p, b200_fir** out)
{
    if (!p || !out)
        return set_err(B200_ERR_ARG, "fir_create: null argument");
    *out = nullptr;
    if (!p->taps || p->n_taps < 1 || p->decimation < 1)
        return set_err(B200_ERR_ARG, "fir_create: need n_taps >= 1, decimation >= 1, taps != NULL");
    if (p->algorithm == 2)
        return set_err(B200_ERR_UNSUPPORTED, "fir_create: tensor-core algorithm not built yet");
    b200_fir* h = new b200_fir();
    h->T = p->n_taps;
    h->D = p->decimation;
    h->vec = p->is_complex ? 2 : 1;
    h->ep = fir_epilogue{ p->fuse_multiply_const ? 1 : 0, p->k_re, p->k_im };
    cudaGetDevice(&h->device);

    // full-rate real streams run as float pairs through the packed complex x real loop
    if (h->vec == 1 && h->D == 1 && !getenv("B200_FIR_REAL_SCALAR"))
        h->rp = 1;
    const int CH = FIR_ACC / (h->rp ? 2 : h->vec);
    const int MT = FIR_NT * (FIR_ACC / h->vec);
    // decimations that divide the 16 (32) window positions of a thread run in the full-rate kernel
    if (h->D > 1 && (FIR_ACC / h->vec) % h->D == 0 && !getenv("B200_FIR_PLANES"))
        h->dd = h->D;
    // ... and the other decimations up to 15 (3, 5, 6, 7, 9 ... 15) with D rows per thread
    if (!h->dd && h->D > 1 && h->D <= (getenv("B200_FIR_DG_MAX") ? atoi(getenv("B200_FIR_DG_MAX")) : 15) && !getenv("B200_FIR_PLANES"))
        h->dg = h->D;
    const int Dg = h->rp ? 2 : (h->dd || h->dg) ? 1 : h->D; // phases the plane / tap geometry is built for
    int tq = (h->T + Dg - 1) / Dg;
    h->TQ = (tq + CH - 1) / CH * CH; // whole 16- (32-) tap steps
    (void)MT;
    {
        int need = (h->dg ? fir_dg_nt(h->dg) * h->dg : FIR_NT) + h->TQ / CH; // rows per plane: per thread + one per tap step
        h->n_boxes = (need + 255) / 256;
        h->box_rows = (need + h->n_boxes - 1) / h->n_boxes;
        h->plane_rows = h->box_rows * h->n_boxes;
        if (const int gr = fir_dg_group(h->dg, h->vec)) { // padded layout: one TMA box per group, one spare row behind each
            h->n_boxes = (need + gr - 1) / gr;
            h->box_rows = gr;
            h->plane_rows = h->n_boxes * (gr + 1);
        }
    }
    h->smem = 16 + sizeof(float) * ((size_t)Dg * h->TQ + (size_t)Dg * (h->plane_rows * 32 + 8)) + 1024;
    if (const char* e = getenv("B200_FIR_TMA"))
        h->use_tma = atoi(e);
    h->algorithm = 1;
    if (h->smem > 200 * 1024 || p->algorithm == 4)
        h->algorithm = 4; // naive global-memory fallback
    // overlap-save fast convolution: complex streams, enough taps per output to pay for two
    // FFTs per block (crossover measured on B200: ~100 taps per output sample)
    {
        const bool can = ols_supported(h->T, h->D, h->vec == 1);
        bool want = (p->algorithm == 3);
        // measured crossovers on B200 (tools/fir_sweep.py, tools/olsd_sweep.py): the full-rate form
        // beats the direct kernel from ~96 taps per output; the polyphase form of a decimating
        // complex filter already from ~40 (even D, TMA-staged; at 32 the two are within 5 %) / ~64 (odd D)
        const int poly = can ? ols_polyphase(h->T, h->D, h->vec == 1) : 0;
        // full-rate filters: the direct kernel pays for taps in pairs of 16-tap steps (64, 96, 128 ...), so
        // 65..96 taps cost 1.5x 64 taps (151 vs 214 GS/s at 16 Mi samples) while overlap-save stays at
        // ~185 (tools/cross_ab.py): switch right after 64 taps, complex and real alike
        const int cross = poly == 1 ? 40 : poly == 2 ? 64 : h->D == 1 ? 65 : 96;
        if (p->algorithm == 0 && can && h->T / h->D >= cross)
            want = true;
        if (p->algorithm == 0 && h->dg && can) {
            // D = 3, 5, 6, 7, 9 ... 15 folded into the full-rate kernel with D rows per thread (tools/decim_ab.py
            // with DS=..., tools/dg_ab.py): complex 400-670 GS/s at 32 taps against 93-280 for overlap-save, real
            // 0.67-1.24 TS/s against ~200; its cost grows with T, overlap-save is flat: measured crossovers (taps)
            static const short tx_c[16] = { 0, 0, 0, 160, 0, 256, 96, 256, 0, 320, 128, 384, 160, 384, 160, 448 };
            static const short tx_r[16] = { 0, 0, 0, 256, 0, 512, 480, 576, 0, 640, 448, 640, 448, 512, 416, 576 };
            const int tx = (h->vec == 2 ? tx_c : tx_r)[h->dg];
            want = h->T > tx;
        } else if (p->algorithm == 0 && !h->dd && h->vec == 2 && h->D > 1 && can) {
            // decimations that do not fold (beyond 15, or B200_FIR_PLANES / B200_FIR_DG_MAX set): the phase-plane kernel
            // stages with per-sample copies and sits at 150-290 GS/s; the polyphase overlap-save is flat at
            // 130-280.  Even D = 6, 10, 12, 14 (TMA-staged polyphase form): from 64 taps; D = 3, 5, 7: beyond 96 taps;
            // larger D: from 192 taps, and whenever the phase planes would not fit shared memory
            if (poly == 1)
                want = h->T >= 64; // (shorter filters stay on the direct kernel: exact impulse response)
            else if (poly == 2)
                want = h->D <= 7 ? h->T > 96 : h->T >= 192;
            else
                want = h->T >= 192 || h->algorithm == 4;
        }
        if (p->algorithm == 0 && h->dd) {
            // decimation folded into the full-rate kernel (dd): short filters are HBM-bound there and
            // its cost grows with T, not T/D; measured crossovers against overlap-save
            // (tools/decim_ab.py, taps): complex 96 / 160 / 192 / 384 for D = 2 / 4 / 8 / 16, real
            // 256 / 512 / 768 / 768 / 768 for D = 2 / 4 / 8 / 16 / 32
            int tx;
            if (h->vec == 2)
                tx = h->dd == 2 ? 96 : h->dd == 4 ? 160 : h->dd == 8 ? 192 : 384;
            else
                tx = h->dd == 2 ? 256 : h->dd == 4 ? 512 : 768;
            want = can && h->T > tx;
        }
        if (const char* e = getenv("B200_FIR_ALGO"))
            if (p->algorithm == 0)
                want = atoi(e) == 3;
        if (want && !can && p->algorithm == 3) {
            b200_fir_destroy(h);
            return set_err(B200_ERR_UNSUPPORTED, "fir_create: overlap-save needs 2..32768 taps");
        }
        if (want && can) {
            int rc = ols_create(p->taps, h->T, h->D, h->vec == 1, h->ep.fuse, h->ep.kre, h->ep.kim, &h->ols);
            if (rc != B200_OK) {
                b200_fir_destroy(h);
                return rc;
            }
            h->algorithm = 3;
        }
    }

    // 2-parallel fast FIR (algorithm 5): full-rate filters in the FP32-bound range below the
    // overlap-save crossover execute 0.77x the FMAs of the direct form
    if (h->algorithm == 1 && h->D == 1) {
        const bool can5 = ffa_supported(h->T, h->D, h->vec == 1);
        bool want5 = p->algorithm == 5;
        // measured (tools/ffa_sweep.py): 0.70x the per-tap cost of the direct form but 3x its fixed
        // cost per tile, so it loses below ~100 taps (64 taps: 192 vs 214 GS/s) and overlap-save wins
        // above: never picked automatically, kept selectable (algorithm = 5 / B200_FIR_ALGO=5)
        if (const char* e = getenv("B200_FIR_ALGO"))
            if (p->algorithm == 0)
                want5 = atoi(e) == 5;
        if (p->algorithm == 5 && !can5) {
            b200_fir_destroy(h);
            return set_err(B200_ERR_UNSUPPORTED, "fir_create: the 2-parallel form needs decimation 1 and 4..2048 taps");
        }
        if (want5 && can5) {
            int rc = ffa_create(p->taps, h->T, h->vec == 1, h->ep.fuse, h->ep.kre, h->ep.kim, &h->ffa);
            if (rc != B200_OK) {
                b200_fir_destroy(h);
                return rc;
            }
            h->algorithm = 5;
        }
    } else if (p->algorithm == 5) {
        b200_fir_destroy(h);
        return set_err(B200_ERR_UNSUPPORTED, "fir_create: the 2-parallel form needs decimation 1");
    }

#define FIR_CUDA(call)                                                                   \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            b200_fir_destroy(h);                                                         \
            return set_err(e__ == cudaErrorMemoryAllocation ? B200_ERR_NOMEM : B200_ERR_CUDA, \
                           "fir_create: %s -> %s", #call, cudaGetErrorString(e__));      \
        }                                                                                \
    } while (0)

    std::vector<float> pp((size_t)Dg * h->TQ, 0.f);
    for (int ph = 0; ph < Dg; ph++)
        for (int qr = 0; qr < h->TQ; qr++) {
            long long k = (long long)(h->TQ - 1 - qr) * Dg + ph;
            pp[(size_t)ph * h->TQ + qr] = (k < h->T) ? p->taps[k] : 0.f;
        }
    FIR_CUDA(cudaMalloc(&h->d_taps_pp, pp.size() * sizeof(float)));
    FIR_CUDA(cudaMemcpy(h->d_taps_pp, pp.data(), pp.size() * sizeof(float), cudaMemcpyHostToDevice));
    FIR_CUDA(cudaMalloc(&h->d_taps, (size_t)h->T * sizeof(float)));
    FIR_CUDA(cudaMemcpy(h->d_taps, p->taps, (size_t)h->T * sizeof(float), cudaMemcpyHostToDevice));
    size_t hb = sizeof(float) * (size_t)h->vec * (size_t)(h->T > 1 ? h->T - 1 : 1);
    for (int i = 0; i < 2; i++) {
        FIR_CUDA(cudaMalloc(&h->d_hist[i], hb));
        FIR_CUDA(cudaMemset(h->d_hist[i], 0, hb));
    }
    if (h->algorithm == 1 && h->dg) {
#define FIR_DG_ATTR(V, DGV)                                                                                   \
    FIR_CUDA(cudaFuncSetAttribute(fir_direct_kernel<V, false, 1, 1, false, DGV>,                               \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024))
        FIR_DG_ATTR(2, 3);
        FIR_DG_ATTR(2, 5);
        FIR_DG_ATTR(2, 6);
        FIR_DG_ATTR(2, 7);
        FIR_DG_ATTR(2, 9);
        FIR_DG_ATTR(2, 10);
        FIR_DG_ATTR(2, 11);
        FIR_DG_ATTR(2, 12);
        FIR_DG_ATTR(2, 13);
        FIR_DG_ATTR(2, 14);
        FIR_DG_ATTR(2, 15);
        FIR_DG_ATTR(1, 3);
        FIR_DG_ATTR(1, 5);
        FIR_DG_ATTR(1, 6);
        FIR_DG_ATTR(1, 7);
        FIR_DG_ATTR(1, 9);
        FIR_DG_ATTR(1, 10);
        FIR_DG_ATTR(1, 11);
        FIR_DG_ATTR(1, 12);
        FIR_DG_ATTR(1, 13);
        FIR_DG_ATTR(1, 14);
        FIR_DG_ATTR(1, 15);
#undef FIR_DG_ATTR
    }
    if (h->algorithm == 1 && h->dd) {
#define FIR_DD_ATTR(V, DDV) \
    FIR_CUDA(cudaFuncSetAttribute(fir_direct_kernel<V, false, DDV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024))
        FIR_DD_ATTR(2, 2);
        FIR_DD_ATTR(2, 4);
        FIR_DD_ATTR(2, 8);
        FIR_DD_ATTR(2, 16);
        FIR_DD_ATTR(1, 2);
        FIR_DD_ATTR(1, 4);
        FIR_DD_ATTR(1, 8);
        FIR_DD_ATTR(1, 16);
        FIR_DD_ATTR(1, 32);
#undef FIR_DD_ATTR
    }
    if (h->algorithm == 1 && h->rp)
        FIR_CUDA(cudaFuncSetAttribute(fir_direct_kernel<2, false, 1, 1, true>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    if (h->algorithm == 1) {
        if (h->vec == 2)
        {
            FIR_CUDA(cudaFuncSetAttribute(fir_direct_kernel<2, true>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            FIR_CUDA(cudaFuncSetAttribute(fir_direct_kernel<2, false>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        }
        else
        {
            FIR_CUDA(cudaFuncSetAttribute(fir_direct_kernel<1, true>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            FIR_CUDA(cudaFuncSetAttribute(fir_direct_kernel<1, false>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        }
    }
#undef FIR_CUDA
    *out = h;
    return B200_OK;
}

int b200_fir_algorithm(const b200_fir* h) { return h ? h->algorithm : 0; }

int b200_fir_geometry(const b200_fir* h, int* decimation, int* item_bytes)
{
    if (!h)
        return set_err(B200_ERR_ARG, "fir_geometry: null handle");
    if (decimation)
        *decimation = h->D;
    if (item_bytes)
        *item_bytes = 4 * h->vec;
    return B200_OK;
}

int b200_fir_run(b200_fir* h, const void* d_in, void* d_out, int64_t n_in_items,
                 int64_t* n_consumed, int64_t* n_produced, b200_stream_t s)
{
    if (!h || n_in_items < 0 || (n_in_items > 0 && !d_in) || (n_in_items >= h->D && !d_out))
        return set_err(B200_ERR_ARG, "fir_run: bad argument");
    long long n_out = n_in_items / h->D;
    long long n_cons = n_out * h->D;
    int rc = fir_launch(h, h->d_hist[h->cur], d_in, d_out, n_in_items, n_out, cs(s));
    if (rc != B200_OK)
        return rc;
    if (h->T > 1 && n_cons > 0) {
        int Tm1 = h->T - 1;
        B200_LAUNCH(fir_hist_kernel, (Tm1 + 255) / 256, 256, 0, cs(s), (const float*)d_in,
                    h->d_hist[h->cur], h->d_hist[h->cur ^ 1], n_cons, Tm1, h->vec);
        h->cur ^= 1;
    }
    if (n_consumed)
        *n_consumed = n_cons;
    if (n_produced)
        *n_produced = n_out;
    return B200_OK;
}

int b200_fir_run_segment(b200_fir* h, const void* d_halo, const void* d_in, void* d_out,
                         int64_t n_in_items, int64_t* n_produced, b200_stream_t s)
{
    if (!h || n_in_items < 0 || (n_in_items > 0 && !d_in) || (n_in_items >= h->D && !d_out))
        return set_err(B200_ERR_ARG, "fir_run_segment: bad argument");
    long long n_out = n_in_items / h->D;
    int rc = fir_launch(h, (const float*)d_halo, d_in, d_out, n_in_items, n_out, cs(s));
    if (rc == B200_OK && n_produced)
        *n_produced = n_out;
    return rc;
}

int b200_fir_reset(b200_fir* h, b200_stream_t s)
{
    if (!h)
        return set_err(B200_ERR_ARG, "fir_reset: null handle");
    size_t hb = sizeof(float) * (size_t)h->vec * (size_t)(h->T > 1 ? h->T - 1 : 1);
    B200_CUDA(cudaMemsetAsync(h->d_hist[h->cur], 0, hb, cs(s)));
    return B200_OK;
}

int b200_fir_set_history(b200_fir* h, const void* d_hist, b200_stream_t s)
{
    if (!h || !d_hist)
        return set_err(B200_ERR_ARG, "fir_set_history: null argument");
    if (h->T > 1)
        B200_CUDA(cudaMemcpyAsync(h->d_hist[h->cur], d_hist,
                                  sizeof(float) * (size_t)h->vec * (size_t)(h->T - 1),
                                  cudaMemcpyDeviceToDevice, cs(s)));
    return B200_OK;
}

int b200_fir_get_history(b200_fir* h, void* d_hist, b200_stream_t s)
{
    if (!h || !d_hist)
        return set_err(B200_ERR_ARG, "fir_get_history: null argument");
    if (h->T > 1)
        B200_CUDA(cudaMemcpyAsync(d_hist, h->d_hist[h->cur],
                                  sizeof(float) * (size_t)h->vec * (size_t)(h->T - 1),
                                  cudaMemcpyDeviceToDevice, cs(s)));
    return B200_OK;
}

} // extern "C"
