// CPU emulator of the CUDA FFT kernels -- TEST INFRASTRUCTURE ONLY (never loaded by the product).
//
// Compiles mpifft4py_b200/csrc/fft_kernels.cuh with g++ (no CUDA) and executes the very same
// phase bodies the __global__ kernels run: all threads of a block are stepped through phase s
// before any thread starts phase s+1, which is exactly what __syncthreads() enforces on the GPU.
// This checks radix plans, digit-reversal, swizzles and every fused index map (pad, truncate +
// fold, peer chunks, masks) in the `-m "not gpu"` suite; the `-m gpu` suite checks the real thing.
#include <cstdio>
#include <map>
#include <vector>

#include "../../mpifft4py_b200/csrc/desc_convert.h"
#include "../../mpifft4py_b200/csrc/fft_plans.h"

using namespace b200fft;

template <class K, int s>
static void emu_phases(const typename K::Params& p, std::vector<unsigned char>& sm, int bx, int by) {
  for (int tid = 0; tid < K::NT; ++tid) K::template phase<s>(p, sm.data(), tid, bx, by);
  if constexpr (s + 1 < K::NPHASE) emu_phases<K, s + 1>(p, sm, bx, by);
}

template <class K>
static int emulate(const typename K::Params& p) {
  const unsigned long long nblk = K::blocks(p);
  std::vector<unsigned char> sm((size_t)(K::SMEM > 0 ? K::SMEM : 16) + 256);
  for (unsigned long long b = 0; b < nblk; ++b) {
    int bx, by;
    K::decode(p, (unsigned)b, bx, by);
    // poison shared memory so that reads of unwritten slots show up
    for (auto& c : sm) c = 0x7f;
    emu_phases<K, 0>(p, sm, bx, by);
  }
  return 0;
}

template <class real>
static const cx<real>* table(int len) {
  static std::map<int, std::vector<cx<real>>> cache;
  auto it = cache.find(len);
  if (it == cache.end()) it = cache.emplace(len, make_twiddles<real>(len)).first;
  return it->second.data();
}

template <class real>
static int strided(const b200fft_strided_desc_t& d) {
  auto p = convert_strided<real>(d, table<real>(d.n), 1);
  switch (d.n) {
#define X(n, ...) \
  case n:         \
    return emulate<StridedK<real, Plan<__VA_ARGS__>>>(p);
    B200FFT_PLANS(X)
#undef X
    default:
      return -1;
  }
}

template <class real, bool FWD>
static int rows(const b200fft_rows_desc_t& d) {
  auto p = convert_rows<real>(d, table<real>(d.n), 1, FWD);
  switch (d.n / 2) {
#define X(n, ...)                                                     \
  case n:                                                             \
    if (FWD) return emulate<R2CK<real, Plan<__VA_ARGS__>>>(p);        \
    else return emulate<C2RK<real, Plan<__VA_ARGS__>>>(p);
    B200FFT_PLANS(X)
#undef X
    default:
      return -1;
  }
}

extern "C" {
int emu_exec_strided(const b200fft_strided_desc_t* d) {
  if (const char* e = check_strided(*d)) { std::fprintf(stderr, "emu: %s\n", e); return 1; }
  return d->precision == B200FFT_DOUBLE ? strided<double>(*d) : strided<float>(*d);
}
int emu_exec_r2c(const b200fft_rows_desc_t* d) {
  if (const char* e = check_rows(*d)) { std::fprintf(stderr, "emu: %s\n", e); return 1; }
  return d->precision == B200FFT_DOUBLE ? rows<double, true>(*d) : rows<float, true>(*d);
}
int emu_exec_c2r(const b200fft_rows_desc_t* d) {
  if (const char* e = check_rows(*d)) { std::fprintf(stderr, "emu: %s\n", e); return 1; }
  return d->precision == B200FFT_DOUBLE ? rows<double, false>(*d) : rows<float, false>(*d);
}
// kernel launch geometry, for DESIGN.md / tests
int emu_strided_config(int precision, int n, int* T, int* TC, int* smem) {
  switch (n) {
#define X(nn, ...)                                                                        \
  case nn:                                                                                \
    if (precision == B200FFT_DOUBLE) {                                                    \
      using C = StridedCfg<double, Plan<__VA_ARGS__>>;                                    \
      *T = C::T; *TC = C::TC; *smem = C::SMEM;                                            \
    } else {                                                                              \
      using C = StridedCfg<float, Plan<__VA_ARGS__>>;                                     \
      *T = C::T; *TC = C::TC; *smem = C::SMEM;                                            \
    }                                                                                     \
    return 0;
    B200FFT_PLANS(X)
#undef X
    default:
      return -1;
  }
}
}
