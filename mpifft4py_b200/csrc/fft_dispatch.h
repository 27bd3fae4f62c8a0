// Internal interface between the kernel translation units and the C-ABI layer.
#pragma once
#include <cuda_runtime.h>
#include "fft_kernels.cuh"

namespace b200fft {
// return cudaError_t as int; -1 when the length has no plan / does not fit shared memory
int launch_strided_f64(int n, const StridedParams<double>& p, cudaStream_t st);
int launch_strided_f32(int n, const StridedParams<float>& p, cudaStream_t st);
int launch_rowc2c_f64(int n, const StridedParams<double>& p, cudaStream_t st);
int launch_rowc2c_f32(int n, const StridedParams<float>& p, cudaStream_t st);
int launch_r2c_f64(int h, const RowParams<double>& p, cudaStream_t st);
int launch_r2c_f32(int h, const RowParams<float>& p, cudaStream_t st);
int launch_c2r_f64(int h, const RowParams<double>& p, cudaStream_t st);
int launch_c2r_f32(int h, const RowParams<float>& p, cudaStream_t st);
// peer-mapped transports: write `value` into n (<= 16) flag words (peer-mapped device memory) from one tiny kernel
int launch_post_flags(unsigned* const* words, int n, unsigned value, cudaStream_t st);
// does the last stage of the plan for complex length n hold the factor 3 (fold-capable)?
bool plan_exists(int n);
}  // namespace b200fft
