// fir_interp.cuh -- interpolation by L and rational resampling by L / M folded into the TMA-staged direct
// FIR kernel (fir.cu), used by resampler.cu for interp_fir_filter / rational_resampler with small L, M
#pragma once
#include <cuda_runtime.h>

struct b200_fir;
namespace b200 {
// decimation == 1: interpolators (L = 2, 3, 4); decimation > 1: the rational ratios with coprime L, M <= 5
// (5/4 for long phases only, 5/2 never: see fir_ratio_min_tq)
// (L passes x M rows per thread, see fir_passes_ll_dg in fir.cu)
bool fir_interp_supported(int n_taps, int interpolation, int decimation, int is_complex);
// handle without history buffers: the caller passes the ceil(T/L)-1 samples of history per launch
int fir_interp_create(const float* taps, int n_taps, int interpolation, int decimation, int is_complex,
                      b200_fir** out);
// n_in = whole groups of `decimation` inputs; produces n_in / decimation * interpolation outputs
int fir_interp_launch(b200_fir* h, const float* d_hist, const void* d_in, void* d_out, long long n_in,
                      cudaStream_t s);
} // namespace b200
