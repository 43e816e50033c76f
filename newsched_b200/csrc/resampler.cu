// resampler.cu -- polyphase interpolating FIR / rational resampler (real taps; ccf / fff).
//
//   y[m] = sum_k h[k] xu[m*D - k],   xu = x with L-1 zeros inserted after every sample
//        = sum_q h[q*L + phi] x[j - q],   m*D = j*L + phi  (0 <= phi < L)
//
// interp_fir_filter is the D == 1 case.  Absent from the reference snapshot (SURVEY.md 0.1, row
// (f)4 of 8); plugs into gr::block::work (runtime/include/gnuradio/block.hpp:81-85) as a
// rate-changing block that consumes D items per L produced, so every call starts at phase 0.
// Only the products that meet a real sample are formed (T/L per output, never the inserted zeros).
//
// Kernel: one CTA per tile of consecutive outputs; the input span the tile touches and the
// phase-major tap table h_pp[phi][q] live in shared memory.  For an interpolator the lanes of a warp
// walk the L tap rows (row stride padded odd: conflict-free) while they share one or two input
// samples (shared-memory broadcast); the history (ceil(T/L)-1 input samples) is carried in the
// handle like the FIR's.
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "fir_interp.cuh"

extern "C" int b200_fir_destroy(b200_fir* h);

namespace b200 {

constexpr int RS_NT = 256;       // threads per CTA
constexpr int RS_SPAN = 6144;    // input samples staged per tile (48 KiB complex)
constexpr int RS_MAX_TAPS = 8192; // padded phase table, floats

struct rs_geom {
    int L, D, Tq, TqP; // taps per phase, padded row stride (odd)
    int MT;            // outputs per tile
    long long n_in, n_out;
};

template <int VEC>
__device__ __forceinline__ void rs_load(float (&v)[VEC], const float* __restrict__ x, const float* __restrict__ hist,
                                        int nh, long long j, long long n_in)
{
#pragma unroll
    for (int c = 0; c < VEC; c++)
        v[c] = 0.f;
    const float* src = nullptr;
    if (j >= 0) {
        if (j < n_in)
            src = x + j * VEC;
    } else if (hist && j >= -(long long)nh)
        src = hist + (nh + j) * VEC;
    if (src) {
#pragma unroll
        for (int c = 0; c < VEC; c++)
            v[c] = __ldg(src + c);
    }
}

template <int VEC>
__global__ void __launch_bounds__(RS_NT)
    resample_kernel(const float* __restrict__ x, const float* __restrict__ hist, float* __restrict__ y,
                    const float* __restrict__ taps_pp, rs_geom g)
{
    extern __shared__ __align__(16) float rs_sm[];
    float* sT = rs_sm;                     // [L][TqP]
    float* sX = rs_sm + ((g.L * g.TqP + 3) & ~3); // [span][VEC]
    const int tid = threadIdx.x;
    for (int i = tid; i < g.L * g.TqP; i += RS_NT)
        sT[i] = __ldg(taps_pp + i);
    const long long m0 = (long long)blockIdx.x * g.MT;
    const long long m1 = min(m0 + g.MT, g.n_out); // exclusive
    const long long j_lo = (m0 * g.D) / g.L - (g.Tq - 1);
    const long long j_hi = ((m1 - 1) * g.D) / g.L; // inclusive
    const int span = (int)(j_hi - j_lo + 1);
    const int nh = g.Tq - 1;
    for (int i = tid; i < span; i += RS_NT) {
        float v[VEC];
        rs_load<VEC>(v, x, hist, nh, j_lo + i, g.n_in);
#pragma unroll
        for (int c = 0; c < VEC; c++)
            sX[i * VEC + c] = v[c];
    }
    __syncthreads();
    for (long long m = m0 + tid; m < m1; m += RS_NT) {
        const long long i = m * g.D;
        const int phi = (int)(i % g.L);
        const int jj = (int)(i / g.L - j_lo); // index of x[j] in the tile; x[j-q] = sX[jj-q]
        const float* h = sT + phi * g.TqP;
        float a0[VEC], a1[VEC];
#pragma unroll
        for (int c = 0; c < VEC; c++)
            a0[c] = a1[c] = 0.f;
        int q = 0;
        for (; q + 2 <= g.Tq; q += 2) {
            const float h0 = h[q], h1 = h[q + 1];
#pragma unroll
            for (int c = 0; c < VEC; c++) {
                a0[c] = fmaf(h0, sX[(jj - q) * VEC + c], a0[c]);
                a1[c] = fmaf(h1, sX[(jj - q - 1) * VEC + c], a1[c]);
            }
        }
        if (q < g.Tq) {
            const float h0 = h[q];
#pragma unroll
            for (int c = 0; c < VEC; c++)
                a0[c] = fmaf(h0, sX[(jj - q) * VEC + c], a0[c]);
        }
#pragma unroll
        for (int c = 0; c < VEC; c++)
            __stcs(y + m * VEC + c, a0[c] + a1[c]);
    }
}

// Register-blocked form for D <= 4 and L <= 256.  Outputs that share a residue rho = m mod L share
// the tap row phi = rho*D mod L and their input index advances by exactly D per output, so a thread
// that owns R such outputs (m = m_base + r*L) runs a plain D-strided FIR: per chunk of 8 taps it
// loads (R-1)*D + 8 samples and 8 taps once and issues 8R packed FMAs on them (D = 1, R = 15:
// 120 FFMA2 per 30 LDS instead of 1 per 1).  Lanes walk the residues, i.e. adjacent lanes read the
// same or neighbouring samples (broadcast) and different tap rows (odd row stride: conflict-free).
// Outputs go back through shared memory so global stores are whole lines.
template <int VEC, int D, int R>
__global__ void __launch_bounds__(RS_NT)
    resample_rb_kernel(const float* __restrict__ x, const float* __restrict__ hist, float* __restrict__ y,
                       const float* __restrict__ taps_pp, rs_geom g, int tpt /* active threads, multiple of L */)
{
    constexpr int WL = (R - 1) * D + 8; // window samples per chunk
    extern __shared__ __align__(16) float rs_sm[];
    float* sT = rs_sm;               // [L][TqP], zero padded to a multiple of 8 taps
    float* sX = rs_sm + ((g.L * g.TqP + 3) & ~3); // [span][VEC], 16-byte aligned; reused for the output tile
    const int tid = threadIdx.x;
    for (int i = tid; i < g.L * g.TqP; i += RS_NT)
        sT[i] = __ldg(taps_pp + i);
    const int TqPad = (g.Tq + 7) & ~7;
    const long long MTt = (long long)tpt * R;
    const long long m0 = (long long)blockIdx.x * MTt; // multiple of L
    const long long j_lo = (m0 * D) / g.L - (TqPad - 1);
    const long long j_hi = ((m0 + MTt - 1) * D) / g.L;
    const int span = (int)(j_hi - j_lo + 1);
    const int nh = g.Tq - 1;
    if (j_lo >= 0 && j_lo + span <= g.n_in) {
        // interior tile: asynchronous global -> shared copies, all of a thread's in flight at once
        const float* src = x + j_lo * VEC;
        for (int i = tid; i < span; i += RS_NT) {
            if (VEC == 2)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(sX + 2 * i)), "l"(src + 2 * i)
                             : "memory");
            else
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(sX + i)), "l"(src + i)
                             : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else {
        for (int i = tid; i < span; i += RS_NT) {
            float v[VEC];
            rs_load<VEC>(v, x, hist, nh, j_lo + i, g.n_in);
#pragma unroll
            for (int c = 0; c < VEC; c++)
                sX[i * VEC + c] = v[c];
        }
    }
    __syncthreads();
    float acc[R * VEC];
#pragma unroll
    for (int l = 0; l < R * VEC; l++)
        acc[l] = 0.f;
    const int rho = tid % g.L, kb = tid / g.L;
    if (tid < tpt) {
        const long long i0 = (m0 + (long long)kb * R * g.L + rho) * D; // m_base * D
        const int phi = (int)(i0 % g.L);
        const int jj = (int)(i0 / g.L - j_lo); // x[j_0] = sX[jj]; output r uses x[j_0 + r D - q]
        const float* h = sT + phi * g.TqP;
        for (int q0 = 0; q0 < TqPad; q0 += 8) {
            float hv[8];
#pragma unroll
            for (int u = 0; u < 8; u++)
                hv[u] = h[q0 + u];
            float w[WL * VEC];
            const float* wp = sX + (jj - q0 - 7) * VEC; // w[i] = x[j_0 - q0 - 7 + i]
#pragma unroll
            for (int i = 0; i < WL; i++) {
                if (VEC == 2) {
                    const float2 t = *reinterpret_cast<const float2*>(wp + 2 * i);
                    w[2 * i] = t.x;
                    w[2 * i + 1] = t.y;
                } else
                    w[i] = wp[i];
            }
#pragma unroll
            for (int u = 0; u < 8; u++)
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const int i = r * D - u + 7;
                    if (VEC == 2) {
                        const float2 a = __ffma2_rn(make_float2(w[2 * i], w[2 * i + 1]), make_float2(hv[u], hv[u]),
                                                    make_float2(acc[2 * r], acc[2 * r + 1]));
                        acc[2 * r] = a.x;
                        acc[2 * r + 1] = a.y;
                    } else
                        acc[r] = fmaf(hv[u], w[i], acc[r]);
                }
        }
    }
    __syncthreads(); // input tile consumed: the output tile goes on top of it
    if (tid < tpt) {
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int o = (kb * R + r) * g.L + rho; // output index within the tile
            if (VEC == 2)
                *reinterpret_cast<float2*>(sX + 2 * o) = make_float2(acc[2 * r], acc[2 * r + 1]);
            else
                sX[o] = acc[r];
        }
    }
    __syncthreads();
    const long long left = g.n_out - m0;
    const int n_tile = (int)(left < MTt ? left : MTt);
    for (int i = tid; i < n_tile; i += RS_NT) {
        if (VEC == 2)
            __stcs(reinterpret_cast<float2*>(y) + m0 + i, *reinterpret_cast<const float2*>(sX + 2 * i));
        else
            __stcs(y + m0 + i, sX[i]);
    }
}

// new history = last nh samples of [old history | first n_cons input samples]
__global__ void rs_hist_kernel(const float* __restrict__ x, const float* __restrict__ h_old,
                               float* __restrict__ h_new, long long n_cons, int nh, int vec)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nh)
        return;
    const long long j = n_cons - nh + i; // sample index relative to x[0]
    for (int c = 0; c < vec; c++)
        h_new[i * vec + c] = j >= 0 ? x[j * vec + c] : h_old[(nh + j) * vec + c];
}

} // namespace b200

using namespace b200;

struct b200_resampler {
    int T = 0, L = 1, D = 1, vec = 2;
    rs_geom g{};
    size_t smem = 0;
    float* d_taps_pp = nullptr;
    float* d_hist[2] = { nullptr, nullptr };
    int cur = 0;
    b200_fir* fold = nullptr; // L <= 4 (D == 1) or coprime L, D <= 5: folded into the TMA-staged direct FIR kernel
    int rb_R = 0;   // > 0: register-blocked kernel with R outputs per thread
    int rb_tpt = 0; // its active threads per CTA (multiple of L)
    size_t rb_smem = 0;
};

template <int VEC>
static int rs_launch_rb(b200_resampler* h, const float* d_hist, const void* d_in, void* d_out, rs_geom g,
                        cudaStream_t s)
{
    const long long per = (long long)h->rb_tpt * h->rb_R;
    const long long tiles = (g.n_out + per - 1) / per;
    if (tiles > 0x7fffffffLL)
        return set_err(B200_ERR_ARG, "resampler: too many items for one call");
#define RS_GO(DD, RR)                                                                                     \
    B200_LAUNCH((resample_rb_kernel<VEC, DD, RR>), (unsigned)tiles, RS_NT, h->rb_smem, s, (const float*)d_in, \
                d_hist, (float*)d_out, h->d_taps_pp, g, h->rb_tpt)
    switch (h->D) {
    case 1: RS_GO(1, 15); break;
    case 2: RS_GO(2, 7); break;
    case 3: RS_GO(3, 6); break;
    default: RS_GO(4, 3); break;
    }
#undef RS_GO
    return B200_OK;
}

static int rs_launch(b200_resampler* h, const float* d_hist, const void* d_in, void* d_out, long long n_in,
                     long long n_out, cudaStream_t s)
{
    if (n_out <= 0)
        return B200_OK;
    if (h->fold)
        return fir_interp_launch(h->fold, d_hist, d_in, d_out, n_in, s);
    rs_geom g = h->g;
    g.n_in = n_in;
    g.n_out = n_out;
    if (h->rb_R > 0)
        return h->vec == 2 ? rs_launch_rb<2>(h, d_hist, d_in, d_out, g, s) : rs_launch_rb<1>(h, d_hist, d_in, d_out, g, s);
    const long long tiles = (n_out + g.MT - 1) / g.MT;
    if (tiles > 0x7fffffffLL)
        return set_err(B200_ERR_ARG, "resampler: too many items for one call");
    if (h->vec == 2)
        B200_LAUNCH(resample_kernel<2>, (unsigned)tiles, RS_NT, h->smem, s, (const float*)d_in, d_hist,
                    (float*)d_out, h->d_taps_pp, g);
    else
        B200_LAUNCH(resample_kernel<1>, (unsigned)tiles, RS_NT, h->smem, s, (const float*)d_in, d_hist,
                    (float*)d_out, h->d_taps_pp, g);
    return B200_OK;
}

extern "C" {

int b200_resampler_destroy(b200_resampler* h)
{
    if (!h)
        return B200_OK;
    cudaFree(h->d_taps_pp);
    cudaFree(h->d_hist[0]);
    cudaFree(h->d_hist[1]);
    if (h->fold)
        b200_fir_destroy(h->fold);
    delete h;
    return B200_OK;
}

int b200_resampler_create(const b200_resampler_params* p, b200_resampler** out)
{
    if (!p || !out)
        return set_err(B200_ERR_ARG, "resampler_create: null argument");
    *out = nullptr;
    if (!p->taps || p->n_taps < 1 || p->interpolation < 1 || p->decimation < 1)
        return set_err(B200_ERR_ARG, "resampler_create: need n_taps >= 1, interpolation >= 1, decimation >= 1");
    b200_resampler* h = new b200_resampler();
    h->T = p->n_taps;
    h->L = p->interpolation;
    h->D = p->decimation;
    h->vec = p->is_complex ? 2 : 1;
    rs_geom& g = h->g;
    g.L = h->L;
    g.D = h->D;
    g.Tq = (h->T + h->L - 1) / h->L;
    g.TqP = ((g.Tq + 7) & ~7) | 1; // rows zero-padded to whole 8-tap chunks, odd stride
    // outputs per tile: as many as keep the staged input span within RS_SPAN samples
    const long long room = (long long)(RS_SPAN - g.Tq - 2) * h->L / h->D;
    if ((long long)g.L * g.TqP > RS_MAX_TAPS || room < 32) {
        b200_resampler_destroy(h);
        return set_err(B200_ERR_UNSUPPORTED,
                       "resampler_create: %d taps / interpolation %d / decimation %d exceed the shared-memory tile "
                       "(use fir_filter for plain decimation)",
                       h->T, h->L, h->D);
    }
    g.MT = (int)std::min<long long>(4 * RS_NT, room);
    h->smem = sizeof(float) * ((size_t)g.L * g.TqP + 4 + (size_t)RS_SPAN * h->vec);
    // register-blocked kernel: D <= 4, L <= 256, and the tile's input span / output tile fit
    if (h->D <= 4 && h->L <= RS_NT && !getenv("B200_RESAMPLER_SIMPLE")) {
        // R*D is the distance (in samples) between the windows of adjacent threads of a residue: kept
        // off multiples of 16 samples (128 B) so that their LDS.64 window reads hit different banks
        static const int Rtab[5] = { 0, 15, 7, 6, 3 };
        const int R = Rtab[h->D];
        const int tpt = RS_NT / h->L * h->L;
        const long long mt = (long long)tpt * R;
        const long long span = (mt * h->D) / h->L + ((g.Tq + 7) & ~7) + 2;
        const long long cells = std::max(span, mt);
        if (cells <= 3 * RS_SPAN) {
            h->rb_R = R;
            h->rb_tpt = tpt;
            h->rb_smem = sizeof(float) * ((size_t)g.L * g.TqP + 4 + (size_t)cells * h->vec);
        }
    }
#define RS_CUDA(call)                                                                        \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            b200_resampler_destroy(h);                                                       \
            return set_err(e__ == cudaErrorMemoryAllocation ? B200_ERR_NOMEM : B200_ERR_CUDA, \
                           "resampler_create: %s -> %s", #call, cudaGetErrorString(e__));    \
        }                                                                                    \
    } while (0)
    std::vector<float> pp((size_t)g.L * g.TqP, 0.f);
    for (int phi = 0; phi < g.L; phi++)
        for (int q = 0; q < g.Tq; q++) {
            const long long k = (long long)q * g.L + phi;
            pp[(size_t)phi * g.TqP + q] = k < h->T ? p->taps[k] : 0.f;
        }
    RS_CUDA(cudaMalloc(&h->d_taps_pp, pp.size() * sizeof(float)));
    RS_CUDA(cudaMemcpy(h->d_taps_pp, pp.data(), pp.size() * sizeof(float), cudaMemcpyHostToDevice));
    const size_t hb = sizeof(float) * (size_t)h->vec * (size_t)std::max(g.Tq - 1, 1);
    for (int i = 0; i < 2; i++) {
        RS_CUDA(cudaMalloc(&h->d_hist[i], hb));
        RS_CUDA(cudaMemset(h->d_hist[i], 0, hb));
    }
    RS_CUDA(cudaFuncSetAttribute(resample_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    RS_CUDA(cudaFuncSetAttribute(resample_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
#define RS_ATTR(V, DD, RR) \
    RS_CUDA(cudaFuncSetAttribute(resample_rb_kernel<V, DD, RR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024))
    RS_ATTR(2, 1, 15);
    RS_ATTR(2, 2, 7);
    RS_ATTR(2, 3, 6);
    RS_ATTR(2, 4, 3);
    RS_ATTR(1, 1, 15);
    RS_ATTR(1, 2, 7);
    RS_ATTR(1, 3, 6);
    RS_ATTR(1, 4, 3);
#undef RS_ATTR
    if (fir_interp_supported(h->T, h->L, h->D, h->vec == 2)) {
        int rc = fir_interp_create(p->taps, h->T, h->L, h->D, h->vec == 2, &h->fold);
        if (rc != B200_OK) {
            b200_resampler_destroy(h);
            return rc;
        }
    }
#undef RS_CUDA
    // the uploads above went through the legacy default stream (cudaMemcpy / cudaMemset); the caller's streams are
    // non-blocking and not ordered against it, so finish them before the handle can be used
    if (cudaDeviceSynchronize() != cudaSuccess) {
        b200_resampler_destroy(h);
        return set_err(B200_ERR_CUDA, "resampler_create: %s", cudaGetErrorString(cudaGetLastError()));
    }
    *out = h;
    return B200_OK;
}

int b200_resampler_geometry(const b200_resampler* h, int* interpolation, int* decimation, int* item_bytes)
{
    if (!h)
        return set_err(B200_ERR_ARG, "resampler_geometry: null handle");
    if (interpolation)
        *interpolation = h->L;
    if (decimation)
        *decimation = h->D;
    if (item_bytes)
        *item_bytes = 4 * h->vec;
    return B200_OK;
}

int b200_resampler_run(b200_resampler* h, const void* d_in, void* d_out, int64_t n_in_items, int64_t* n_consumed,
                       int64_t* n_produced, b200_stream_t s)
{
    if (!h || n_in_items < 0 || (n_in_items > 0 && !d_in) || (n_in_items >= h->D && !d_out))
        return set_err(B200_ERR_ARG, "resampler_run: bad argument");
    const long long chunks = n_in_items / h->D;
    const long long n_cons = chunks * h->D, n_out = chunks * h->L;
    int rc = rs_launch(h, h->d_hist[h->cur], d_in, d_out, n_cons, n_out, cs(s));
    if (rc != B200_OK)
        return rc;
    const int nh = h->g.Tq - 1;
    if (nh > 0 && n_cons > 0) {
        B200_LAUNCH(rs_hist_kernel, (nh + 255) / 256, 256, 0, cs(s), (const float*)d_in, h->d_hist[h->cur],
                    h->d_hist[h->cur ^ 1], n_cons, nh, h->vec);
        h->cur ^= 1;
    }
    if (n_consumed)
        *n_consumed = n_cons;
    if (n_produced)
        *n_produced = n_out;
    return B200_OK;
}

int b200_resampler_run_segment(b200_resampler* h, const void* d_halo, const void* d_in, void* d_out,
                               int64_t n_in_items, int64_t* n_produced, b200_stream_t s)
{
    if (!h || n_in_items < 0 || (n_in_items > 0 && !d_in) || (n_in_items >= h->D && !d_out))
        return set_err(B200_ERR_ARG, "resampler_run_segment: bad argument");
    const long long chunks = n_in_items / h->D;
    int rc = rs_launch(h, (const float*)d_halo, d_in, d_out, chunks * h->D, chunks * h->L, cs(s));
    if (rc == B200_OK && n_produced)
        *n_produced = chunks * h->L;
    return rc;
}

int b200_resampler_reset(b200_resampler* h, b200_stream_t s)
{
    if (!h)
        return set_err(B200_ERR_ARG, "resampler_reset: null handle");
    const size_t hb = sizeof(float) * (size_t)h->vec * (size_t)std::max(h->g.Tq - 1, 1);
    B200_CUDA(cudaMemsetAsync(h->d_hist[h->cur], 0, hb, cs(s)));
    return B200_OK;
}

} // extern "C"
