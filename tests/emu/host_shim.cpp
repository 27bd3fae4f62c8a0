// libb200fft built for the HOST -- TEST INFRASTRUCTURE ONLY (never loaded by the product).
// The real C-ABI layer (mpifft4py_b200/csrc/b200fft.cu, compiled as C++ against tests/emu/cuda_shim) with
// its kernel launchers bound to the CPU emulator below.  What runs is the product's own plan objects,
// program execution loop, descriptor construction, fused-launch fallback logic and timing records; what
// is emulated is the kernels (same phase bodies, tests/emu/emu.cpp) and the CUDA runtime (inert).
#include "emu.cpp"

#include "../../mpifft4py_b200/csrc/fft_dispatch.h"

namespace b200fft {
template <class real>
static int h_strided(int n, const StridedParams<real>& p) {
  switch (n) {
#define X(nn, ...) \
  case nn:         \
    return emulate<StridedK<real, Plan<__VA_ARGS__>>>(p);
    B200FFT_PLANS(X)
#undef X
    default:
      return -1;
  }
}
template <class real>
static int h_rowc2c(int n, const StridedParams<real>& p) {
  switch (n) {
#define X(nn, ...) \
  case nn:         \
    return emulate<RowC2CK<real, Plan<__VA_ARGS__>>>(p);
    B200FFT_PLANS(X)
#undef X
    default:
      return -1;
  }
}
template <class real, bool FWD>
static int h_rows(int h, const RowParams<real>& p) {
  switch (h) {
#define X(nn, ...)                                                   \
  case nn:                                                           \
    if (FWD) return emulate<R2CK<real, Plan<__VA_ARGS__>>>(p);       \
    else return emulate<C2RK<real, Plan<__VA_ARGS__>>>(p);
    B200FFT_ROW_PLANS(X)
#undef X
    default:
      return -1;
  }
}
template <class real>
static int h_fused(int H, int NY, const RowParams<real>& pr, const StridedParams<real>& ps, int inverse_order, int ppg) {
#define X(h, ny, PR, PC)                                                                                   \
  if (H == h && NY == ny)                                                                                  \
    return inverse_order ? emulate_fused_pair<StridedK<real, PC>, C2RK<real, PR>>(ps, pr, ps.B, ppg)        \
                         : emulate_fused_pair<R2CK<real, PR>, StridedK<real, PC>>(pr, ps, ps.B, ppg);
  B200FFT_FUSED_PAIRS(X)
#undef X
  return -1;
}

int launch_strided_f64(int n, const StridedParams<double>& p, cudaStream_t) { return h_strided<double>(n, p); }
int launch_strided_f32(int n, const StridedParams<float>& p, cudaStream_t) { return h_strided<float>(n, p); }
int launch_rowc2c_f64(int n, const StridedParams<double>& p, cudaStream_t) { return h_rowc2c<double>(n, p); }
int launch_rowc2c_f32(int n, const StridedParams<float>& p, cudaStream_t) { return h_rowc2c<float>(n, p); }
int launch_r2c_f64(int h, const RowParams<double>& p, cudaStream_t) { return h_rows<double, true>(h, p); }
int launch_r2c_f32(int h, const RowParams<float>& p, cudaStream_t) { return h_rows<float, true>(h, p); }
int launch_c2r_f64(int h, const RowParams<double>& p, cudaStream_t) { return h_rows<double, false>(h, p); }
int launch_c2r_f32(int h, const RowParams<float>& p, cudaStream_t) { return h_rows<float, false>(h, p); }
int launch_fused_zy_f64(int H, int NY, const RowParams<double>& pr, const StridedParams<double>& ps, int inverse_order, int ppg,
                        unsigned*, cudaStream_t) {
  return h_fused<double>(H, NY, pr, ps, inverse_order, ppg);
}
int launch_fused_zy_f32(int H, int NY, const RowParams<float>& pr, const StridedParams<float>& ps, int inverse_order, int ppg,
                        unsigned*, cudaStream_t) {
  return h_fused<float>(H, NY, pr, ps, inverse_order, ppg);
}
}  // namespace b200fft

#include "../../mpifft4py_b200/csrc/b200fft.cu"
