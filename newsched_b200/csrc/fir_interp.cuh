// fir_interp.cuh -- interpolation by L folded into the TMA-staged direct FIR kernel (fir.cu), used by
// resampler.cu for interp_fir_filter with small L
#pragma once
#include <cuda_runtime.h>

struct b200_fir;
namespace b200 {
bool fir_interp_supported(int n_taps, int interpolation, int is_complex);
// handle without history buffers: the caller passes the ceil(T/L)-1 samples of history per launch
int fir_interp_create(const float* taps, int n_taps, int interpolation, int is_complex, b200_fir** out);
int fir_interp_launch(b200_fir* h, const float* d_hist, const void* d_in, void* d_out, long long n_in,
                      cudaStream_t s);
} // namespace b200
