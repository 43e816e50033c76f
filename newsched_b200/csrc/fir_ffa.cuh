// fir_ffa.cuh -- internal interface of the 2-parallel fast FIR form (fir_ffa.cu), used by fir.cu
#pragma once
#include <cuda_runtime.h>

namespace b200 {
struct ffa_plan;
bool ffa_supported(int n_taps, int decimation, int real);
int ffa_create(const float* taps, int n_taps, int real, int fuse, float kre, float kim, ffa_plan** out);
void ffa_destroy(ffa_plan* p);
int ffa_launch(ffa_plan* p, const float* d_hist, const void* d_in, void* d_out, long long n_in,
               long long n_out, cudaStream_t s);
} // namespace b200
