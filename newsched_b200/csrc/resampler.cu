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
    float* sX = rs_sm + g.L * g.TqP;       // [span][VEC]
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
};

static int rs_launch(b200_resampler* h, const float* d_hist, const void* d_in, void* d_out, long long n_in,
                     long long n_out, cudaStream_t s)
{
    if (n_out <= 0)
        return B200_OK;
    rs_geom g = h->g;
    g.n_in = n_in;
    g.n_out = n_out;
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
    g.TqP = g.Tq | 1;
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
    h->smem = sizeof(float) * ((size_t)g.L * g.TqP + (size_t)RS_SPAN * h->vec);
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
#undef RS_CUDA
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
