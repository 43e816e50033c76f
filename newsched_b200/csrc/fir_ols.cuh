// fir_ols.cuh -- internal interface of the overlap-save FIR (fir_ols.cu), used by fir.cu
#pragma once
#include <cuda_runtime.h>

namespace b200 {
struct ols_plan;
bool ols_supported(int n_taps, int decimation, int real);
// 0 = full-rate form, 1 = polyphase form staged by TMA (even D), 2 = polyphase with direct loads (odd D)
int ols_polyphase(int n_taps, int decimation, int real);
int ols_create(const float* taps, int n_taps, int decimation, int real, int fuse, float kre, float kim,
               ols_plan** out);
void ols_destroy(ols_plan* p);
int ols_launch(ols_plan* p, const float* d_hist, const void* d_in, void* d_out, long long n_in,
               long long n_out, cudaStream_t s);
} // namespace b200
