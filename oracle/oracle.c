/*
 * oracle.c -- CPU restatement of the newsched data-parallel block hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under newsched_b200/ or include/ may call,
 * link or load this file.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, as the checker or as
 * the timed CPU baseline -- never as the product path.
 *
 * Parity status per function:
 *   orc_copy                 PINNED   follows blocklib/blocks/include/gnuradio/blocklib/blocks/copy.hpp:33-44;
 *                                     golden = ramp (i,-i) of schedulers/mt/test/cuda/qa_scheduler_mt_cuda_copy.cpp:24-28
 *   orc_multiply_const_*     PINNED for k=1 (schedulers/mt/test/qa_scheduler_mt.cpp:86-88,128-132),
 *                            "parity unpinned" for k != 1: the arithmetic lives in VOLK
 *                            (volk >= 2.2, CI pins v2.2.1; call sites blocklib/blocks/lib/multiply_const.cpp:27,42)
 *                            which is not vendored; restated as the IEEE fp32 non-fused product.
 *   orc_fir_* / orc_fft_* / orc_complex_to_mag / orc_pfb_channelizer
 *                            "parity unpinned": these blocks do not exist in the mounted
 *                            reference snapshot.  The definitions restate the
 *                            GNU Radio block semantics fixed in SURVEY.md 8(c) and are cross
 *                            checked against numpy/scipy in tests/test_oracle.py.
 *
 * Build: see oracle/Makefile (gcc -O3 -march=native -fopenmp -shared).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

ORC_API int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

ORC_API void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0)
        omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------ copy
 * blocks::copy::work, copy.hpp:33-44: memcpy(out, in, n_items*itemsize). */
ORC_API void orc_copy(void* out, const void* in, int64_t n_items, int64_t itemsize)
{
    memcpy(out, in, (size_t)(n_items * itemsize));
}

/* -------------------------------------------------------- multiply_const
 * multiply_const<float>::work, multiply_const.cpp:19-31 -> volk_32f_s32f_multiply_32f:
 * out[i] = in[i] * k over n_items*vlen scalars. */
ORC_API void orc_multiply_const_ff(float* out, const float* in, float k, int64_t n)
{
    for (int64_t i = 0; i < n; i++)
        out[i] = in[i] * k;
}

/* multiply_const<gr_complex>::work, multiply_const.cpp:35-47 ->
 * volk_32fc_s32fc_multiply_32fc: (a+bi)(c+di) = (ac-bd) + (ad+bc)i, each product
 * rounded to fp32 before the add (no FMA contraction: built with -ffp-contract=off). */
ORC_API void
orc_multiply_const_cc(float* out, const float* in, float kre, float kim, int64_t n)
{
    for (int64_t i = 0; i < n; i++) {
        float a = in[2 * i], b = in[2 * i + 1];
        float ac = a * kre, bd = b * kim, ad = a * kim, bc = b * kre;
        out[2 * i] = ac - bd;
        out[2 * i + 1] = ad + bc;
    }
}

/* generic template path, multiply_const.cpp:54-81: *optr++ = *iptr++ * d_k with the
 * C++ integer promotions: int16 result truncated back to 16 bits, int32 wraps. */
ORC_API void orc_multiply_const_ss(int16_t* out, const int16_t* in, int16_t k, int64_t n)
{
    for (int64_t i = 0; i < n; i++)
        out[i] = (int16_t)((int32_t)in[i] * (int32_t)k);
}

ORC_API void orc_multiply_const_ii(int32_t* out, const int32_t* in, int32_t k, int64_t n)
{
    for (int64_t i = 0; i < n; i++)
        out[i] = (int32_t)((uint32_t)in[i] * (uint32_t)k);
}

/* two-input blocks (GNU Radio multiply_XX / add_XX semantics; absent from the snapshot):
 * element-wise product / sum, complex product with each partial product rounded to fp32 */
ORC_API void orc_multiply_ff(float* out, const float* a, const float* b, int64_t n)
{
    for (int64_t i = 0; i < n; i++)
        out[i] = a[i] * b[i];
}
ORC_API void orc_multiply_cc(float* out, const float* a, const float* b, int64_t n)
{
    for (int64_t i = 0; i < n; i++) {
        float ar = a[2 * i], ai = a[2 * i + 1], br = b[2 * i], bi = b[2 * i + 1];
        float p0 = ar * br, p1 = ai * bi, p2 = ar * bi, p3 = ai * br;
        out[2 * i] = p0 - p1;
        out[2 * i + 1] = p2 + p3;
    }
}
ORC_API void orc_add_f(float* out, const float* a, const float* b, int64_t n_floats)
{
    for (int64_t i = 0; i < n_floats; i++)
        out[i] = a[i] + b[i];
}

/* -------------------------------------------------------- complex_to_mag
 * y = sqrtf(re*re + im*im), fp32, not hypot (volk_32fc_magnitude_32f
 * semantics). */
ORC_API void orc_complex_to_mag(float* out, const float* in, int64_t n)
{
    for (int64_t i = 0; i < n; i++) {
        float re = in[2 * i], im = in[2 * i + 1];
        float rr = re * re, ii = im * im;
        out[i] = sqrtf(rr + ii);
    }
}

ORC_API void orc_complex_to_mag_squared(float* out, const float* in, int64_t n)
{
    for (int64_t i = 0; i < n; i++) {
        float re = in[2 * i], im = in[2 * i + 1];
        float rr = re * re, ii = im * im;
        out[i] = rr + ii;
    }
}

/* ------------------------------------------------------------ fir_filter
 * y[m] = sum_{k<T} h[k] * x[m*D - k], m = 0 .. floor(n_in/D)-1,
 * decimation phase 0, x[n] for n < 0 taken from `hist` (the T-1 samples that
 * precede x[0], oldest first; NULL means zeros).
 *
 * Two flavours: *_f32 accumulates in fp32 over 8 lanes the way a SIMD dot product
 * does (the order a CPU block would use); *_f64 accumulates in double = "truth".
 * Returns the number of outputs written. */
static inline float fir_x_re(const float* x, const float* hist, int T, int64_t n, int c, int vec)
{
    if (n >= 0)
        return x[n * vec + c];
    if (!hist)
        return 0.0f;
    int64_t j = (int64_t)(T - 1) + n; /* hist[T-1-1] is x[-1] */
    if (j < 0)
        return 0.0f;
    return hist[j * vec + c];
}

static int64_t fir_f64(float* out,
                       const float* in,
                       int64_t n_in,
                       const float* taps,
                       int T,
                       int D,
                       const float* hist,
                       int vec,
                       int mt)
{
    int64_t n_out = n_in / D;
#pragma omp parallel for schedule(static) if (mt)
    for (int64_t m = 0; m < n_out; m++) {
        double acc[2] = { 0.0, 0.0 };
        int64_t base = m * D;
        for (int k = 0; k < T; k++) {
            int64_t n = base - k;
            for (int c = 0; c < vec; c++)
                acc[c] += (double)taps[k] * (double)fir_x_re(in, hist, T, n, c, vec);
        }
        for (int c = 0; c < vec; c++)
            out[m * vec + c] = (float)acc[c];
    }
    return n_out;
}

/* fp32, 8 partial sums per component (AVX-style), taps reversed walk. */
static int64_t fir_f32(float* out,
                       const float* in,
                       int64_t n_in,
                       const float* taps,
                       int T,
                       int D,
                       const float* hist,
                       int vec,
                       int mt)
{
    int64_t n_out = n_in / D;
    /* reversed taps so the inner loop walks x forward: y[m] = sum_j hr[j] x[mD-(T-1)+j] */
    float* hr = (float*)malloc(sizeof(float) * (size_t)T);
    for (int j = 0; j < T; j++)
        hr[j] = taps[T - 1 - j];
#pragma omp parallel for schedule(static) if (mt)
    for (int64_t m = 0; m < n_out; m++) {
        int64_t start = m * D - (T - 1);
        float accs[2][8];
        memset(accs, 0, sizeof(accs));
        if (start >= 0) {
            const float* xp = in + start * vec;
            if (vec == 2) {
                int j = 0;
                for (; j + 8 <= T; j += 8)
                    for (int l = 0; l < 8; l++) {
                        accs[0][l] += hr[j + l] * xp[2 * (j + l)];
                        accs[1][l] += hr[j + l] * xp[2 * (j + l) + 1];
                    }
                for (; j < T; j++) {
                    accs[0][j & 7] += hr[j] * xp[2 * j];
                    accs[1][j & 7] += hr[j] * xp[2 * j + 1];
                }
            } else {
                int j = 0;
                for (; j + 8 <= T; j += 8)
                    for (int l = 0; l < 8; l++)
                        accs[0][l] += hr[j + l] * xp[j + l];
                for (; j < T; j++)
                    accs[0][j & 7] += hr[j] * xp[j];
            }
        } else {
            for (int j = 0; j < T; j++)
                for (int c = 0; c < vec; c++)
                    accs[c][j & 7] += hr[j] * fir_x_re(in, hist, T, start + j, c, vec);
        }
        for (int c = 0; c < vec; c++) {
            float s = ((accs[c][0] + accs[c][4]) + (accs[c][1] + accs[c][5])) +
                      ((accs[c][2] + accs[c][6]) + (accs[c][3] + accs[c][7]));
            out[m * vec + c] = s;
        }
    }
    free(hr);
    return n_out;
}

ORC_API int64_t orc_fir_ccf_f64(float* out, const float* in, int64_t n_in, const float* taps,
                                int T, int D, const float* hist)
{
    return fir_f64(out, in, n_in, taps, T, D, hist, 2, 0);
}
ORC_API int64_t orc_fir_fff_f64(float* out, const float* in, int64_t n_in, const float* taps,
                                int T, int D, const float* hist)
{
    return fir_f64(out, in, n_in, taps, T, D, hist, 1, 0);
}
ORC_API int64_t orc_fir_ccf_f32(float* out, const float* in, int64_t n_in, const float* taps,
                                int T, int D, const float* hist)
{
    return fir_f32(out, in, n_in, taps, T, D, hist, 2, 0);
}
ORC_API int64_t orc_fir_fff_f32(float* out, const float* in, int64_t n_in, const float* taps,
                                int T, int D, const float* hist)
{
    return fir_f32(out, in, n_in, taps, T, D, hist, 1, 0);
}
/* all-host-cores variants: outputs are independent, split over m (time segments). */
ORC_API int64_t orc_fir_ccf_f64_mt(float* out, const float* in, int64_t n_in, const float* taps,
                                   int T, int D, const float* hist)
{
    return fir_f64(out, in, n_in, taps, T, D, hist, 2, 1);
}
ORC_API int64_t orc_fir_ccf_f32_mt(float* out, const float* in, int64_t n_in, const float* taps,
                                   int T, int D, const float* hist)
{
    return fir_f32(out, in, n_in, taps, T, D, hist, 2, 1);
}
ORC_API int64_t orc_fir_fff_f32_mt(float* out, const float* in, int64_t n_in, const float* taps,
                                   int T, int D, const float* hist)
{
    return fir_f32(out, in, n_in, taps, T, D, hist, 1, 1);
}

/* ---------------------------------------------------------------- resampler
 * interp_fir_filter / rational_resampler (absent from the reference
 * snapshot -- "parity unpinned", cross-checked against scipy.signal.upfirdn in tests/test_oracle.py):
 *   y[m] = sum_k h[k] xu[m D - k],  xu[i] = x[i / L] if L divides i else 0,
 * i.e. y[m] = sum_q h[q L + phi] x[j - q] with m D = j L + phi.  floor(n_in / D) groups of D inputs
 * give L outputs each.  hist = the ceil(T/L)-1 samples before in[0] (oldest first) or NULL. */
static int64_t resample_f64(float* out, const float* in, int64_t n_in, const float* taps, int T, int L, int D,
                            const float* hist, int vec)
{
    const int64_t n_out = (n_in / D) * L;
    const int Tq = (T + L - 1) / L, nh = Tq - 1;
    for (int64_t m = 0; m < n_out; m++) {
        const int64_t i = m * D, j = i / L;
        const int phi = (int)(i % L);
        double acc[2] = { 0.0, 0.0 };
        for (int q = 0; (int64_t)q * L + phi < T; q++) {
            const int64_t s = j - q;
            const double h = taps[(int64_t)q * L + phi];
            for (int c = 0; c < vec; c++) {
                double xv = 0.0;
                if (s >= 0)
                    xv = in[s * vec + c];
                else if (hist && s >= -(int64_t)nh)
                    xv = hist[(nh + s) * vec + c];
                acc[c] += h * xv;
            }
        }
        for (int c = 0; c < vec; c++)
            out[m * vec + c] = (float)acc[c];
    }
    return n_out;
}
ORC_API int64_t orc_resample_ccf_f64(float* out, const float* in, int64_t n_in, const float* taps, int T, int L,
                                     int D, const float* hist)
{
    return resample_f64(out, in, n_in, taps, T, L, D, hist, 2);
}
ORC_API int64_t orc_resample_fff_f64(float* out, const float* in, int64_t n_in, const float* taps, int T, int L,
                                     int D, const float* hist)
{
    return resample_f64(out, in, n_in, taps, T, L, D, hist, 1);
}

/* ---------------------------------------------------------------- window
 * 4-term 92 dB Blackman-Harris, symmetric:
 * w[n] = a0 - a1 cos(2 pi n/(N-1)) + a2 cos(4 pi n/(N-1)) - a3 cos(6 pi n/(N-1)). */
ORC_API void orc_window_blackmanharris(float* w, int N)
{
    const double a0 = 0.35875, a1 = 0.48829, a2 = 0.14128, a3 = 0.01168;
    for (int n = 0; n < N; n++) {
        double t = (N > 1) ? (double)n / (double)(N - 1) : 0.0;
        w[n] = (float)(a0 - a1 * cos(2.0 * M_PI * t) + a2 * cos(4.0 * M_PI * t) -
                       a3 * cos(6.0 * M_PI * t));
    }
}

/* ------------------------------------------------------------------- fft
 * SURVEY.md 8(c) (GNU Radio fft_vcc semantics): per item of N complex64,
 *   forward: t[n] = x[n]*w[n];                    X[k] = sum_n t[n] e^{-j 2 pi k n / N};
 *            shift -> out[j] = X[(j + N/2) mod N] (DC lands on index N/2)
 *   reverse: shift -> u[n] = x[(n + N/2) mod N] else u = x;  t[n] = u[n]*w[n];
 *            X[k] = sum_n t[n] e^{+j 2 pi k n / N}
 * No 1/N scaling in either direction.  window == NULL means no window.
 * N must be a power of two.  Computed in double (radix-2 DIT) = "truth". */
static void fft_pow2_f64(double* re, double* im, int N, int sign)
{
    /* bit reversal */
    for (int i = 1, j = 0; i < N; i++) {
        int bit = N >> 1;
        for (; j & bit; bit >>= 1)
            j ^= bit;
        j ^= bit;
        if (i < j) {
            double t = re[i]; re[i] = re[j]; re[j] = t;
            t = im[i]; im[i] = im[j]; im[j] = t;
        }
    }
    for (int len = 2; len <= N; len <<= 1) {
        int half = len >> 1;
        for (int j = 0; j < half; j++) {
            double ang = sign * 2.0 * M_PI * (double)j / (double)len;
            double wr = cos(ang), wi = sin(ang);
            for (int i = j; i < N; i += len) {
                int k = i + half;
                double tr = re[k] * wr - im[k] * wi;
                double ti = re[k] * wi + im[k] * wr;
                re[k] = re[i] - tr; im[k] = im[i] - ti;
                re[i] += tr; im[i] += ti;
            }
        }
    }
}

ORC_API int orc_fft_f64(float* out, const float* in, int64_t n_vec, int N, int forward,
                        const float* window, int shift)
{
    if (N < 1 || (N & (N - 1)))
        return -1;
    int err = 0;
#pragma omp parallel
    {
        double* re = (double*)malloc(sizeof(double) * (size_t)N);
        double* im = (double*)malloc(sizeof(double) * (size_t)N);
        if (!re || !im) {
#pragma omp atomic write
            err = -2;
        } else {
#pragma omp for schedule(static)
            for (int64_t v = 0; v < n_vec; v++) {
                const float* x = in + v * 2 * (int64_t)N;
                float* y = out + v * 2 * (int64_t)N;
                for (int n = 0; n < N; n++) {
                    int src = (!forward && shift) ? ((n + N / 2) % N) : n;
                    double w = window ? (double)window[n] : 1.0;
                    /* the block multiplies in fp32 (volk_32fc_32f_multiply_32fc) */
                    re[n] = (double)(float)(x[2 * src] * (float)w);
                    im[n] = (double)(float)(x[2 * src + 1] * (float)w);
                }
                fft_pow2_f64(re, im, N, forward ? -1 : +1);
                for (int j = 0; j < N; j++) {
                    int k = (forward && shift) ? ((j + N / 2) % N) : j;
                    y[2 * j] = (float)re[k];
                    y[2 * j + 1] = (float)im[k];
                }
            }
        }
        free(re);
        free(im);
    }
    return err;
}

/* fp32 FFT "as a CPU block would compute it": iterative radix-4 Stockham autosort with
 * a precomputed fp32 twiddle table, vectorisable inner loops; used as the timed CPU
 * baseline (bench.py cpu_baseline / --impl reference) and as a second checker. */
typedef struct {
    int N;
    int log2n;
    float* twr; /* N/2 twiddles, forward sign */
    float* twi;
} orc_fft_plan;

ORC_API orc_fft_plan* orc_fft_plan_create(int N)
{
    if (N < 1 || (N & (N - 1)))
        return NULL;
    orc_fft_plan* p = (orc_fft_plan*)calloc(1, sizeof(*p));
    p->N = N;
    p->log2n = 0;
    while ((1 << p->log2n) < N)
        p->log2n++;
    p->twr = (float*)malloc(sizeof(float) * (size_t)N);
    p->twi = (float*)malloc(sizeof(float) * (size_t)N);
    for (int j = 0; j < N; j++) { /* full circle: the radix-4 passes need W^{3p} */
        double ang = -2.0 * M_PI * (double)j / (double)N;
        p->twr[j] = (float)cos(ang);
        p->twi[j] = (float)sin(ang);
    }
    return p;
}

ORC_API void orc_fft_plan_destroy(orc_fft_plan* p)
{
    if (!p)
        return;
    free(p->twr);
    free(p->twi);
    free(p);
}

/* Stockham autosort FFT, split re/im planes, radix-4 passes (one radix-2 pass first when log2 N
 * is odd).  Each pass reads x and writes y, then the roles swap; the return value says where the
 * result ended up (0 = the buffers passed as x, 1 = the buffers passed as y).
 * n = current transform length, s = stride (number of interleaved sub-sequences). */
static int stockham_f32(const orc_fft_plan* p, float* xr, float* xi, float* yr, float* yi,
                        float sign)
{
    const int N = p->N;
    int s = 1, n = N, where = 0;
    float* t;
    if (p->log2n & 1) { /* radix-2 pass */
        int m = n >> 1;
        int tstep = N / n;
        for (int q = 0; q < m; q++) {
            float wr = p->twr[q * tstep], wi = sign * p->twi[q * tstep];
            float a_r = xr[q], a_i = xi[q], b_r = xr[q + m], b_i = xi[q + m];
            float sr = a_r - b_r, si = a_i - b_i;
            yr[2 * q] = a_r + b_r;
            yi[2 * q] = a_i + b_i;
            yr[2 * q + 1] = sr * wr - si * wi;
            yi[2 * q + 1] = sr * wi + si * wr;
        }
        t = xr; xr = yr; yr = t;
        t = xi; xi = yi; yi = t;
        where ^= 1;
        n >>= 1;
        s <<= 1;
    }
    for (; n > 1; n >>= 2, s <<= 2) {
        const int m = n >> 2;
        const int tstep = N / n;
        for (int q = 0; q < m; q++) {
            const float w1r = p->twr[q * tstep], w1i = sign * p->twi[q * tstep];
            const float w2r = p->twr[2 * q * tstep], w2i = sign * p->twi[2 * q * tstep];
            const float w3r = p->twr[3 * q * tstep], w3i = sign * p->twi[3 * q * tstep];
            const float *ar = xr + s * q, *ai = xi + s * q;
            const float *br = ar + s * m, *bi = ai + s * m;
            const float *cr = br + s * m, *ci = bi + s * m;
            const float *dr = cr + s * m, *di = ci + s * m;
            float *o0r = yr + s * 4 * q, *o0i = yi + s * 4 * q;
            float *o1r = o0r + s, *o1i = o0i + s, *o2r = o1r + s, *o2i = o1i + s, *o3r = o2r + s, *o3i = o2i + s;
            for (int j = 0; j < s; j++) {
                float apcr = ar[j] + cr[j], apci = ai[j] + ci[j];
                float amcr = ar[j] - cr[j], amci = ai[j] - ci[j];
                float bpdr = br[j] + dr[j], bpdi = bi[j] + di[j];
                float bmdr = br[j] - dr[j], bmdi = bi[j] - di[j];
                /* forward (sign=+1): -j (b-d) = (bmdi, -bmdr); reverse: +j (b-d) */
                float jr = sign * bmdi, ji = -sign * bmdr;
                float t1r = amcr + jr, t1i = amci + ji;
                float t2r = apcr - bpdr, t2i = apci - bpdi;
                float t3r = amcr - jr, t3i = amci - ji;
                o0r[j] = apcr + bpdr;
                o0i[j] = apci + bpdi;
                o1r[j] = t1r * w1r - t1i * w1i;
                o1i[j] = t1r * w1i + t1i * w1r;
                o2r[j] = t2r * w2r - t2i * w2i;
                o2i[j] = t2r * w2i + t2i * w2r;
                o3r[j] = t3r * w3r - t3i * w3i;
                o3i[j] = t3r * w3i + t3i * w3r;
            }
        }
        t = xr; xr = yr; yr = t;
        t = xi; xi = yi; yi = t;
        where ^= 1;
    }
    return where;
}

ORC_API int orc_fft_f32(const orc_fft_plan* p, float* out, const float* in, int64_t n_vec,
                        int forward, const float* window, int shift, int mag, int mt)
{
    if (!p)
        return -1;
    const int N = p->N;
    int err = 0;
#pragma omp parallel if (mt)
    {
        float* buf = (float*)malloc(sizeof(float) * 4 * (size_t)N);
        if (!buf) {
#pragma omp atomic write
            err = -2;
        } else {
            float *xr = buf, *xi = buf + N, *yr = buf + 2 * N, *yi = buf + 3 * N;
#pragma omp for schedule(static)
            for (int64_t v = 0; v < n_vec; v++) {
                const float* x = in + v * 2 * (int64_t)N;
                for (int n = 0; n < N; n++) {
                    int src = (!forward && shift) ? ((n + N / 2) % N) : n;
                    float w = window ? window[n] : 1.0f;
                    xr[n] = x[2 * src] * w;
                    xi[n] = x[2 * src + 1] * w;
                }
                int where = stockham_f32(p, xr, xi, yr, yi, forward ? 1.0f : -1.0f);
                const float* rr = where ? yr : xr;
                const float* ri = where ? yi : xi;
                if (mag) {
                    float* y = out + v * (int64_t)N;
                    for (int j = 0; j < N; j++) {
                        int k = (forward && shift) ? ((j + N / 2) % N) : j;
                        y[j] = sqrtf(rr[k] * rr[k] + ri[k] * ri[k]);
                    }
                } else {
                    float* y = out + v * 2 * (int64_t)N;
                    for (int j = 0; j < N; j++) {
                        int k = (forward && shift) ? ((j + N / 2) % N) : j;
                        y[2 * j] = rr[k];
                        y[2 * j + 1] = ri[k];
                    }
                }
            }
        }
        free(buf);
    }
    return err;
}

/* --------------------------------------------------- polyphase channelizer
 * critically sampled analysis bank, M channels, prototype taps
 * h[0..T), T = M*P, branch p_i[r] = h[i + r*M];
 *   u_i[t] = sum_{r<P} p_i[r] * x[(t - r)*M + (M-1-i)]      (zeros before stream start,
 *                                                             or `hist` = the (P-1)*M samples before x[0])
 *   y_c[t] = sum_{i<M} u_i[t] * e^{+j 2 pi i c / M}          (un-normalised reverse DFT)
 * Output layout: out[t*M + c] (one M-vector per output time), t = 0 .. floor(n_in/M)-1.
 * Truth flavour: double accumulate, direct O(M^2) DFT per frame. */
ORC_API int64_t orc_pfb_channelizer_f64(float* out, const float* in, int64_t n_in,
                                        const float* taps, int M, int P, const float* hist)
{
    int64_t n_t = n_in / M;
    int64_t nh = (int64_t)(P - 1) * M;
    double* cs = (double*)malloc(sizeof(double) * 2 * (size_t)M);
    for (int i = 0; i < M; i++) {
        cs[2 * i] = cos(2.0 * M_PI * (double)i / (double)M);
        cs[2 * i + 1] = sin(2.0 * M_PI * (double)i / (double)M);
    }
#pragma omp parallel
    {
        double* u = (double*)malloc(sizeof(double) * 2 * (size_t)M);
#pragma omp for schedule(static)
        for (int64_t t = 0; t < n_t; t++) {
            for (int i = 0; i < M; i++) {
                double ar = 0.0, ai = 0.0;
                for (int r = 0; r < P; r++) {
                    int64_t n = (t - r) * M + (M - 1 - i);
                    double xr = 0.0, xi = 0.0;
                    if (n >= 0) {
                        xr = in[2 * n];
                        xi = in[2 * n + 1];
                    } else if (hist && nh + n >= 0) {
                        xr = hist[2 * (nh + n)];
                        xi = hist[2 * (nh + n) + 1];
                    }
                    double h = taps[i + r * M];
                    ar += h * xr;
                    ai += h * xi;
                }
                u[2 * i] = ar;
                u[2 * i + 1] = ai;
            }
            for (int c = 0; c < M; c++) {
                double yr = 0.0, yi = 0.0;
                for (int i = 0; i < M; i++) {
                    int idx = (int)(((int64_t)i * c) % M);
                    double wr = cs[2 * idx], wi = cs[2 * idx + 1];
                    yr += u[2 * i] * wr - u[2 * i + 1] * wi;
                    yi += u[2 * i] * wi + u[2 * i + 1] * wr;
                }
                out[2 * (t * M + c)] = (float)yr;
                out[2 * (t * M + c) + 1] = (float)yi;
            }
        }
        free(u);
    }
    free(cs);
    return n_t;
}

/* fp32 flavour used as the timed CPU baseline: fp32 branch filters + fp32 radix-2 FFT
 * across branches (M power of two). */
ORC_API int64_t orc_pfb_channelizer_f32(float* out, const float* in, int64_t n_in,
                                        const float* taps, int M, int P, const float* hist,
                                        int mt)
{
    int64_t n_t = n_in / M;
    int64_t nh = (int64_t)(P - 1) * M;
    orc_fft_plan* plan = orc_fft_plan_create(M);
    if (!plan)
        return -1;
#pragma omp parallel if (mt)
    {
        float* buf = (float*)malloc(sizeof(float) * 4 * (size_t)M);
        float *xr = buf, *xi = buf + M, *yr = buf + 2 * M, *yi = buf + 3 * M;
#pragma omp for schedule(static)
        for (int64_t t = 0; t < n_t; t++) {
            for (int i = 0; i < M; i++) {
                float ar = 0.0f, ai = 0.0f;
                for (int r = 0; r < P; r++) {
                    int64_t n = (t - r) * M + (M - 1 - i);
                    float vr = 0.0f, vi = 0.0f;
                    if (n >= 0) {
                        vr = in[2 * n];
                        vi = in[2 * n + 1];
                    } else if (hist && nh + n >= 0) {
                        vr = hist[2 * (nh + n)];
                        vi = hist[2 * (nh + n) + 1];
                    }
                    float h = taps[i + r * M];
                    ar += h * vr;
                    ai += h * vi;
                }
                xr[i] = ar;
                xi[i] = ai;
            }
            int where = stockham_f32(plan, xr, xi, yr, yi, -1.0f); /* reverse DFT: e^{+j} */
            const float* rr = where ? yr : xr;
            const float* ri = where ? yi : xi;
            for (int c = 0; c < M; c++) {
                out[2 * (t * M + c)] = rr[c];
                out[2 * (t * M + c) + 1] = ri[c];
            }
        }
        free(buf);
    }
    orc_fft_plan_destroy(plan);
    return n_t;
}

/* --------------------------------------------------------- multi-thread elementwise
 * (timed CPU baseline helpers) */
ORC_API void orc_multiply_const_cc_mt(float* out, const float* in, float kre, float kim,
                                      int64_t n)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        float a = in[2 * i], b = in[2 * i + 1];
        float ac = a * kre, bd = b * kim, ad = a * kim, bc = b * kre;
        out[2 * i] = ac - bd;
        out[2 * i + 1] = ad + bc;
    }
}
