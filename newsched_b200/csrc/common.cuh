// common.cuh -- shared helpers for libb200dsp (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/b200dsp.h"

namespace b200 {

// thread-local error text behind b200_last_error()
char* err_buf();
int set_err(int code, const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

inline cudaStream_t cs(b200_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

#define B200_CUDA(call)                                                                  \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess)                                                          \
            return ::b200::set_err(B200_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__,  \
                                   #call, cudaGetErrorString(e__));                      \
    } while (0)

// every kernel launch goes through this so b200_launch_count() is the library's own claim
#define B200_LAUNCH(kernel, grid, block, smem, stream, ...)                              \
    do {                                                                                 \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                      \
        ::b200::g_launches.fetch_add(1, std::memory_order_relaxed);                      \
        cudaError_t e__ = cudaPeekAtLastError();                                         \
        if (e__ != cudaSuccess)                                                          \
            return ::b200::set_err(B200_ERR_CUDA, "%s:%d launch %s -> %s", __FILE__,     \
                                   __LINE__, #kernel, cudaGetErrorString(e__));          \
    } while (0)

int sm_count();

// non-fused complex product (matches the oracle / VOLK generic kernel): each product is
// rounded to fp32 before the add.
__device__ __forceinline__ float2 cmul_nofma(float2 a, float kre, float kim)
{
    float ac = __fmul_rn(a.x, kre), bd = __fmul_rn(a.y, kim);
    float ad = __fmul_rn(a.x, kim), bc = __fmul_rn(a.y, kre);
    return make_float2(__fsub_rn(ac, bd), __fadd_rn(ad, bc));
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
    return make_float2(fmaf(-a.y, b.y, a.x * b.x), fmaf(a.x, b.y, a.y * b.x));
}

__device__ __forceinline__ float mag_nofma(float2 a)
{
    return __fsqrt_rn(__fadd_rn(__fmul_rn(a.x, a.x), __fmul_rn(a.y, a.y)));
}

} // namespace b200
