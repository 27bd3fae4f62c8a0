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
constexpr int FUSE_CTL_WORDS = 4097;  // control words of a fused launch: queue head + up to 4096 group counters
// z + y (inverse_order: y + z) passes of a single-rank slab plan as one persistent kernel through L2; H = half the
// real row length, NY = column length, ppg = planes (batch entries of the strided pass) per group;
// -1: no kernel for this size pair, -3: too many / too small groups
int launch_fused_zy_f64(int H, int NY, const RowParams<double>& pr, const StridedParams<double>& ps, int inverse_order, int ppg,
                        unsigned* ctl, cudaStream_t st);
int launch_fused_zy_f32(int H, int NY, const RowParams<float>& pr, const StridedParams<float>& ps, int inverse_order, int ppg,
                        unsigned* ctl, cudaStream_t st);
// does the last stage of the plan for complex length n hold the factor 3 (fold-capable)?
bool plan_exists(int n);
// tuning switch (B200FFT_VARIANT environment variable, b200fft_set_variant): 0 = default kernels
int kernel_variant();
// switches that combine: variant = 100 + bits
enum { VAR_CLUSTER_FAR = 1, VAR_CLUSTER_LONG = 2, VAR_ROW_BARRIERS = 4, VAR_C2R_DIRECT = 8, VAR_ROW_OCC4 = 16, VAR_R2C_PAIRED = 32, VAR_C2R_PAIRED = 64, VAR_STRIDED_DIRECT = 128 };
inline bool variant_in_range(int v) { return v >= 100 && v < 356; }
inline bool variant_has(int flag) {
  const int v = kernel_variant();
  return variant_in_range(v) && ((v - 100) & flag) != 0;
}
}  // namespace b200fft
