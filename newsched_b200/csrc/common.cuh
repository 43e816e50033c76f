// common.cuh -- shared helpers for libb200dsp (sm_100a only).
#pragma once
#include <cuda.h>
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

// Launch with programmatic stream serialization (programmatic dependent launch): the kernel's CTAs may become
// resident while the previous kernel in the stream drains.  The KERNEL must order itself: griddepcontrol.wait
// before it touches anything a predecessor in the stream may have written or may still read.
#define B200_LAUNCH_PDL(kernel, grid_, block_, smem_, stream_, ...)                          \
    do {                                                                                 \
        cudaLaunchConfig_t cfg__{};                                                      \
        cfg__.gridDim = dim3(grid_);                                                     \
        cfg__.blockDim = dim3(block_);                                                   \
        cfg__.dynamicSmemBytes = (smem_);                                                \
        cfg__.stream = (stream_);                                                        \
        cudaLaunchAttribute attr__[1];                                                   \
        attr__[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;               \
        attr__[0].val.programmaticStreamSerializationAllowed = 1;                        \
        cfg__.attrs = attr__;                                                            \
        cfg__.numAttrs = 1;                                                              \
        cudaError_t e__ = cudaLaunchKernelEx(&cfg__, kernel, __VA_ARGS__);               \
        ::b200::g_launches.fetch_add(1, std::memory_order_relaxed);                      \
        if (e__ != cudaSuccess)                                                          \
            return ::b200::set_err(B200_ERR_CUDA, "%s:%d launch %s -> %s", __FILE__,     \
                                   __LINE__, #kernel, cudaGetErrorString(e__));          \
    } while (0)

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

int sm_count();

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library does not link libcuda);
// nullptr when the driver does not provide it
typedef CUresult (*tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
tmap_encode_fn tmap_encode_tiled();

// non-fused complex product (matches the oracle / VOLK generic kernel): each product is
// rounded to fp32 before the add.
__device__ __forceinline__ float2 cmul_nofma(float2 a, float kre, float kim)
{
    float ac = __fmul_rn(a.x, kre), bd = __fmul_rn(a.y, kim);
    float ad = __fmul_rn(a.x, kim), bc = __fmul_rn(a.y, kre);
    return make_float2(__fsub_rn(ac, bd), __fadd_rn(ad, bc));
}

// Complex products can be written as TWO packed instructions: FMUL2 b * a.x (scalar broadcast),
// then FFMA2 with the operand (-b.y, b.x), which the assembler expresses as a lane-swap + negate
// modifier on b (SASS: FMUL2 R, b.F32x2.HI_LO, a.x.F32 ; FFMA2 R, -b.F32x2.LO_HI.NP, a.y.F32, R):
// half the issue slots of the 2 FMUL + 2 FFMA form, no extra registers.  Measured on the B200
// (tools/pk_ab.py): 14 % fewer instructions in the FFT-4096 kernel, and 5 % SLOWER
// (fused window+FFT+|.| 449 -> 425 GS/s, overlap-save 182 -> 173): the packed forms serialise
// product and accumulate on one pipe where the scalar pair dual-issues.  Default off.
#ifndef B200_PK_CMUL
#define B200_PK_CMUL 0
#endif
__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
#if B200_PK_CMUL
    const float2 r = __fmul2_rn(b, make_float2(a.x, a.x));
    return __ffma2_rn(make_float2(-b.y, b.x), make_float2(a.y, a.y), r);
#else
    return make_float2(fmaf(-a.y, b.y, a.x * b.x), fmaf(a.x, b.y, a.y * b.x));
#endif
}
// a * conj(b)
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b)
{
#if B200_PK_CMUL
    const float2 r = __fmul2_rn(make_float2(b.x, -b.y), make_float2(a.x, a.x));
    return __ffma2_rn(make_float2(b.y, b.x), make_float2(a.y, a.y), r);
#else
    return make_float2(fmaf(a.y, b.y, a.x * b.x), fmaf(-a.x, b.y, a.y * b.x));
#endif
}
// acc + a * b
__device__ __forceinline__ float2 cmac(float2 a, float2 b, float2 acc)
{
#if B200_PK_CMUL
    acc = __ffma2_rn(b, make_float2(a.x, a.x), acc);
    return __ffma2_rn(make_float2(-b.y, b.x), make_float2(a.y, a.y), acc);
#else
    return make_float2(fmaf(a.x, b.x, fmaf(-a.y, b.y, acc.x)), fmaf(a.x, b.y, fmaf(a.y, b.x, acc.y)));
#endif
}

__device__ __forceinline__ float mag_nofma(float2 a)
{
    return __fsqrt_rn(__fadd_rn(__fmul_rn(a.x, a.x), __fmul_rn(a.y, a.y)));
}

// ---- mbarrier / bulk-copy (TMA) helpers: raw PTX, sm_90+ ---------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// 1-D bulk copy global -> shared, completion signalled on an mbarrier (bytes % 16 == 0, both
// addresses 16-byte aligned)
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                              uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                 "selp.u32 %0, 1, 0, p;\n"
                 "}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}
// bounded wait: a broken pipeline traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); spins++)
        if (spins > (1u << 28))
            __trap();
}

} // namespace b200
