// elementwise.cu -- copy, multiply_const (ff/cc/ss/ii), complex_to_mag(_squared).
//
// All five are pure HBM streams (16 B/sample for c64 copy / multiply_const, 12 B/sample for
// complex_to_mag), so the kernels are nothing but wide coalesced traffic:
// one 16-byte load per thread per step, EW_UNROLL independent steps in flight per thread,
// block-contiguous spans so every warp instruction covers whole 128-byte lines.
// Misaligned heads / tails (the ring advances in items, not in 16-byte units) are peeled
// into a scalar launch so the body always runs the 16-byte path.
//
// Reference semantics: blocks::copy::work (copy.hpp:33-44),
// blocks::multiply_const<T>::work (multiply_const.cpp:19-81).  The reference CUDA versions
// (blocklib/cuda/lib/copy.cu:6-28: one launch per 8 KiB item, component-wise 4-byte stores;
// multiply_const.cu:1-17: 64-thread blocks, uninitialised k) are what this replaces.
#include "common.cuh"

namespace b200 {

constexpr int EW_THREADS = 256;
constexpr int EW_UNROLL = 4;

// ---- ops: vin is always a 16-byte vector ------------------------------------------------
struct op_copy16 {
    using ein = uint8_t;
    using eout = uint8_t;
    using vin = uint4;
    using vout = uint4;
    static constexpr int EV = 16;
    __device__ __forceinline__ vout vec(vin v) const { return v; }
    __device__ __forceinline__ eout one(ein v) const { return v; }
};
struct op_mul_ff {
    using ein = float;
    using eout = float;
    using vin = float4;
    using vout = float4;
    static constexpr int EV = 4;
    float k;
    __device__ __forceinline__ vout vec(vin v) const
    {
        return make_float4(__fmul_rn(v.x, k), __fmul_rn(v.y, k), __fmul_rn(v.z, k), __fmul_rn(v.w, k));
    }
    __device__ __forceinline__ eout one(ein v) const { return __fmul_rn(v, k); }
};
struct op_mul_cc {
    using ein = float2;
    using eout = float2;
    using vin = float4;
    using vout = float4;
    static constexpr int EV = 2;
    float kre, kim;
    __device__ __forceinline__ vout vec(vin v) const
    {
        float2 a = cmul_nofma(make_float2(v.x, v.y), kre, kim);
        float2 b = cmul_nofma(make_float2(v.z, v.w), kre, kim);
        return make_float4(a.x, a.y, b.x, b.y);
    }
    __device__ __forceinline__ eout one(ein v) const { return cmul_nofma(v, kre, kim); }
};
struct op_mul_ss {
    using ein = int16_t;
    using eout = int16_t;
    using vin = uint4;
    using vout = uint4;
    static constexpr int EV = 8;
    int16_t k;
    __device__ __forceinline__ uint32_t pair(uint32_t w) const
    {
        int32_t lo = (int32_t)(int16_t)(w & 0xffffu) * (int32_t)k;
        int32_t hi = (int32_t)(int16_t)(w >> 16) * (int32_t)k;
        return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16);
    }
    __device__ __forceinline__ vout vec(vin v) const
    {
        return make_uint4(pair(v.x), pair(v.y), pair(v.z), pair(v.w));
    }
    __device__ __forceinline__ eout one(ein v) const { return (int16_t)((int32_t)v * (int32_t)k); }
};
struct op_mul_ii {
    using ein = int32_t;
    using eout = int32_t;
    using vin = uint4;
    using vout = uint4;
    static constexpr int EV = 4;
    uint32_t k;
    __device__ __forceinline__ vout vec(vin v) const
    {
        return make_uint4(v.x * k, v.y * k, v.z * k, v.w * k);
    }
    __device__ __forceinline__ eout one(ein v) const { return (int32_t)((uint32_t)v * k); }
};
template <bool SQUARED>
struct op_mag {
    using ein = float2;
    using eout = float;
    using vin = float4;
    using vout = float2;
    static constexpr int EV = 2;
    __device__ __forceinline__ float m(float re, float im) const
    {
        float s = __fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im));
        return SQUARED ? s : __fsqrt_rn(s);
    }
    __device__ __forceinline__ vout vec(vin v) const { return make_float2(m(v.x, v.y), m(v.z, v.w)); }
    __device__ __forceinline__ eout one(ein v) const { return m(v.x, v.y); }
};

template <class Op>
__global__ void __launch_bounds__(EW_THREADS)
    ew_vec_kernel(const typename Op::vin* __restrict__ in, typename Op::vout* __restrict__ out,
                  size_t nvec, Op op)
{
    size_t base = (size_t)blockIdx.x * (EW_THREADS * EW_UNROLL) + threadIdx.x;
    typename Op::vin v[EW_UNROLL];
#pragma unroll
    for (int u = 0; u < EW_UNROLL; u++) {
        size_t i = base + (size_t)u * EW_THREADS;
        if (i < nvec)
            v[u] = __ldcs(in + i);
    }
#pragma unroll
    for (int u = 0; u < EW_UNROLL; u++) {
        size_t i = base + (size_t)u * EW_THREADS;
        if (i < nvec)
            __stcs(out + i, op.vec(v[u]));
    }
}

// scalar path for misaligned heads/tails (and fully misaligned buffers): two index ranges
// [0, n0) and [s1, s1+n1) in one launch.
template <class Op>
__global__ void __launch_bounds__(EW_THREADS)
    ew_scalar_kernel(const typename Op::ein* __restrict__ in, typename Op::eout* __restrict__ out,
                     size_t n0, size_t s1, size_t n1, Op op)
{
    size_t i = (size_t)blockIdx.x * EW_THREADS + threadIdx.x;
    if (i < n0)
        out[i] = op.one(in[i]);
    else if (i < n0 + n1) {
        size_t j = s1 + (i - n0);
        out[j] = op.one(in[j]);
    }
}

template <class Op>
static int ew_launch(const void* d_in, void* d_out, size_t n, Op op, cudaStream_t s)
{
    using ein = typename Op::ein;
    using eout = typename Op::eout;
    if (n == 0)
        return B200_OK;
    if (!d_in || !d_out)
        return set_err(B200_ERR_ARG, "elementwise: null pointer");
    uintptr_t pi = (uintptr_t)d_in, po = (uintptr_t)d_out;
    if (pi % alignof(ein) || po % alignof(eout))
        return set_err(B200_ERR_ARG, "elementwise: pointer not aligned to the item type");
    // smallest head (in elements) after which both streams are vector aligned
    size_t head = (size_t)-1;
    for (size_t h = 0; h < 2 * (size_t)Op::EV; h++) {
        if ((pi + h * sizeof(ein)) % sizeof(typename Op::vin) == 0 &&
            (po + h * sizeof(eout)) % sizeof(typename Op::vout) == 0) {
            head = h;
            break;
        }
    }
    const ein* in = reinterpret_cast<const ein*>(d_in);
    eout* out = reinterpret_cast<eout*>(d_out);
    if (head == (size_t)-1 || head >= n) {
        size_t blocks = (n + EW_THREADS - 1) / EW_THREADS;
        B200_LAUNCH((ew_scalar_kernel<Op>), (unsigned)blocks, EW_THREADS, 0, s, in, out, n, (size_t)0,
                    (size_t)0, op);
        return B200_OK;
    }
    size_t nvec = (n - head) / Op::EV;
    size_t tail_start = head + nvec * Op::EV;
    size_t tail = n - tail_start;
    if (nvec) {
        size_t per_block = (size_t)EW_THREADS * EW_UNROLL;
        size_t blocks = (nvec + per_block - 1) / per_block;
        if (blocks > 0x7fffffffull)
            return set_err(B200_ERR_ARG, "elementwise: too many items for one call");
        B200_LAUNCH((ew_vec_kernel<Op>), (unsigned)blocks, EW_THREADS, 0, s,
                    reinterpret_cast<const typename Op::vin*>(in + head),
                    reinterpret_cast<typename Op::vout*>(out + head), nvec, op);
    }
    if (head + tail) {
        size_t blocks = (head + tail + EW_THREADS - 1) / EW_THREADS;
        B200_LAUNCH((ew_scalar_kernel<Op>), (unsigned)blocks, EW_THREADS, 0, s, in, out, head,
                    tail_start, tail, op);
    }
    return B200_OK;
}

// narrower copies for buffers whose relative alignment is below 16 bytes
template <typename V>
__global__ void __launch_bounds__(EW_THREADS)
    copy_narrow_kernel(const V* __restrict__ in, V* __restrict__ out, size_t n)
{
    size_t i = (size_t)blockIdx.x * EW_THREADS + threadIdx.x;
    if (i < n)
        out[i] = in[i];
}

template <typename V>
static int copy_narrow(const uint8_t* in, uint8_t* out, size_t n_bytes, cudaStream_t s)
{
    // head bytes until `in` (and therefore `out`) is V-aligned, body in V, tail in bytes
    size_t head = (sizeof(V) - ((uintptr_t)in % sizeof(V))) % sizeof(V);
    if (head > n_bytes)
        head = n_bytes;
    size_t nv = (n_bytes - head) / sizeof(V);
    size_t tail_start = head + nv * sizeof(V);
    size_t tail = n_bytes - tail_start;
    if (nv) {
        size_t blocks = (nv + EW_THREADS - 1) / EW_THREADS;
        B200_LAUNCH((copy_narrow_kernel<V>), (unsigned)blocks, EW_THREADS, 0, s,
                    reinterpret_cast<const V*>(in + head), reinterpret_cast<V*>(out + head), nv);
    }
    if (head + tail) {
        op_copy16 op;
        B200_LAUNCH((ew_scalar_kernel<op_copy16>), 1, EW_THREADS, 0, s, in, out, head, tail_start,
                    tail, op);
    }
    return B200_OK;
}

// ---- two-input ops: out[i] = f(a[i], b[i]) over float lanes (complex = 2 lanes) ------------------
// MODE 0: a+b (ff and cc are the same thing lane-wise), 1: a*b real, 2: a*b complex
template <int MODE>
__device__ __forceinline__ float4 bin4(float4 a, float4 b)
{
    if (MODE == 0)
        return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
    if (MODE == 1)
        return make_float4(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y), __fmul_rn(a.z, b.z), __fmul_rn(a.w, b.w));
    float2 p = cmul_nofma(make_float2(a.x, a.y), b.x, b.y);
    float2 q = cmul_nofma(make_float2(a.z, a.w), b.z, b.w);
    return make_float4(p.x, p.y, q.x, q.y);
}

template <int MODE>
__global__ void __launch_bounds__(EW_THREADS)
    ew_binary_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                     size_t n_floats, int vec_ok)
{
    const size_t nvec = vec_ok ? n_floats / 4 : 0;
    size_t base = (size_t)blockIdx.x * (EW_THREADS * EW_UNROLL) + threadIdx.x;
    float4 va[EW_UNROLL], vb[EW_UNROLL];
#pragma unroll
    for (int u = 0; u < EW_UNROLL; u++) {
        size_t i = base + (size_t)u * EW_THREADS;
        if (i < nvec) {
            va[u] = __ldcs(reinterpret_cast<const float4*>(a) + i);
            vb[u] = __ldcs(reinterpret_cast<const float4*>(b) + i);
        }
    }
#pragma unroll
    for (int u = 0; u < EW_UNROLL; u++) {
        size_t i = base + (size_t)u * EW_THREADS;
        if (i < nvec)
            __stcs(reinterpret_cast<float4*>(out) + i, bin4<MODE>(va[u], vb[u]));
    }
    // scalar remainder (or everything when the buffers are not 16-byte aligned): complex pairs
    // stay together because the lane count is even for MODE 2
    const size_t step = (MODE == 2) ? 2 : 1;
    for (size_t j = nvec * 4 + ((size_t)blockIdx.x * EW_THREADS + threadIdx.x) * step; j < n_floats;
         j += (size_t)gridDim.x * EW_THREADS * step) {
        if (MODE == 0)
            out[j] = __fadd_rn(a[j], b[j]);
        else if (MODE == 1)
            out[j] = __fmul_rn(a[j], b[j]);
        else {
            float2 p = cmul_nofma(make_float2(a[j], a[j + 1]), b[j], b[j + 1]);
            out[j] = p.x;
            out[j + 1] = p.y;
        }
    }
}

template <int MODE>
static int ew_binary(const void* a, const void* b, void* out, size_t n_floats, cudaStream_t s)
{
    if (n_floats == 0)
        return B200_OK;
    if (!a || !b || !out)
        return set_err(B200_ERR_ARG, "elementwise: null pointer");
    const uintptr_t al = (MODE == 2) ? 8 : 4;
    if ((uintptr_t)a % al || (uintptr_t)b % al || (uintptr_t)out % al)
        return set_err(B200_ERR_ARG, "elementwise: pointer not aligned to the item type");
    const int vec_ok = ((uintptr_t)a % 16 == 0 && (uintptr_t)b % 16 == 0 && (uintptr_t)out % 16 == 0) ? 1 : 0;
    const size_t per_block = (size_t)EW_THREADS * EW_UNROLL * 4;
    size_t blocks = vec_ok ? (n_floats + per_block - 1) / per_block : (n_floats + EW_THREADS - 1) / EW_THREADS;
    if (blocks > 0x7fffffffull)
        return set_err(B200_ERR_ARG, "elementwise: too many items for one call");
    B200_LAUNCH((ew_binary_kernel<MODE>), (unsigned)blocks, EW_THREADS, 0, s, (const float*)a, (const float*)b,
                (float*)out, n_floats, vec_ok);
    return B200_OK;
}

} // namespace b200

using namespace b200;

extern "C" {

int b200_copy(void* d_out, const void* d_in, size_t n_bytes, b200_stream_t s)
{
    if (n_bytes == 0)
        return B200_OK;
    if (!d_in || !d_out)
        return set_err(B200_ERR_ARG, "copy: null pointer");
    uintptr_t rel = ((uintptr_t)d_in ^ (uintptr_t)d_out);
    const uint8_t* in = (const uint8_t*)d_in;
    uint8_t* out = (uint8_t*)d_out;
    if (rel % 16 == 0)
        return ew_launch<op_copy16>(d_in, d_out, n_bytes, op_copy16{}, cs(s));
    if (rel % 8 == 0)
        return copy_narrow<uint2>(in, out, n_bytes, cs(s));
    if (rel % 4 == 0)
        return copy_narrow<uint32_t>(in, out, n_bytes, cs(s));
    if (rel % 2 == 0)
        return copy_narrow<uint16_t>(in, out, n_bytes, cs(s));
    return copy_narrow<uint8_t>(in, out, n_bytes, cs(s));
}

int b200_multiply_const_ff(float* d_out, const float* d_in, float k, size_t n, b200_stream_t s)
{
    return ew_launch<op_mul_ff>(d_in, d_out, n, op_mul_ff{ k }, cs(s));
}
int b200_multiply_const_cc(void* d_out, const void* d_in, float k_re, float k_im, size_t n,
                           b200_stream_t s)
{
    return ew_launch<op_mul_cc>(d_in, d_out, n, op_mul_cc{ k_re, k_im }, cs(s));
}
int b200_multiply_const_ss(int16_t* d_out, const int16_t* d_in, int16_t k, size_t n, b200_stream_t s)
{
    return ew_launch<op_mul_ss>(d_in, d_out, n, op_mul_ss{ k }, cs(s));
}
int b200_multiply_const_ii(int32_t* d_out, const int32_t* d_in, int32_t k, size_t n, b200_stream_t s)
{
    return ew_launch<op_mul_ii>(d_in, d_out, n, op_mul_ii{ (uint32_t)k }, cs(s));
}
int b200_multiply_ff(float* d_out, const float* d_a, const float* d_b, size_t n, b200_stream_t s)
{
    return ew_binary<1>(d_a, d_b, d_out, n, cs(s));
}
int b200_multiply_cc(void* d_out, const void* d_a, const void* d_b, size_t n, b200_stream_t s)
{
    return ew_binary<2>(d_a, d_b, d_out, 2 * n, cs(s));
}
int b200_add_ff(float* d_out, const float* d_a, const float* d_b, size_t n, b200_stream_t s)
{
    return ew_binary<0>(d_a, d_b, d_out, n, cs(s));
}
int b200_add_cc(void* d_out, const void* d_a, const void* d_b, size_t n, b200_stream_t s)
{
    return ew_binary<0>(d_a, d_b, d_out, 2 * n, cs(s));
}
int b200_complex_to_mag(float* d_out, const void* d_in, size_t n, b200_stream_t s)
{
    return ew_launch<op_mag<false>>(d_in, d_out, n, op_mag<false>{}, cs(s));
}
int b200_complex_to_mag_squared(float* d_out, const void* d_in, size_t n, b200_stream_t s)
{
    return ew_launch<op_mag<true>>(d_in, d_out, n, op_mag<true>{}, cs(s));
}

} // extern "C"
