// Radix plans: every complex length the engine supports, 2^k and 3*2^k (3/2-rule padded sizes).
// The factor 3 always sits in the LAST stage so that the Nyquist fold of the 3/2-rule
// truncation (modes +N/2 and -N/2 of the padded spectrum) is local to one thread.
// X(n, radices...)
#pragma once
#define B200FFT_PLANS(X)                                                        \
  X(2, 2) X(4, 4) X(8, 8) X(16, 16) X(32, 4, 8) X(64, 8, 8) X(128, 16, 8)       \
  X(256, 16, 16) X(512, 8, 8, 8) X(1024, 16, 8, 8) X(2048, 16, 16, 8)           \
  X(4096, 16, 16, 16) X(8192, 16, 8, 8, 8) X(16384, 16, 16, 8, 8)               \
  X(3, 3) X(6, 6) X(12, 12) X(24, 2, 12) X(48, 4, 12) X(96, 8, 12)              \
  X(192, 16, 12) X(384, 4, 8, 12) X(768, 8, 8, 12) X(1536, 16, 8, 12)           \
  X(3072, 16, 16, 12) X(6144, 8, 8, 8, 12) X(12288, 16, 8, 8, 12)

// Row (R2C / C2R) kernels: same radix plans.  (Taking the factor 3 in the first stage removes the
// shared-memory bank conflicts of the 12-long runs in the middle stage but measured 3-5% slower
// for n = 1536 double -- the radix-12 twiddle tree costs more FP64 than the conflicts cost LSU.)
#define B200FFT_ROW_PLANS(X) B200FFT_PLANS(X)

// Precision-specific overrides of the strided pass: X(n, min CTAs per SM (0 = auto), tile row bytes
// (0 = auto), radices...).  None at present: with chained twiddle powers the double-precision
// (16, 8, 8) plan fits the 80-register budget of three CTAs per SM and measured 10% faster than
// four stages of radix <= 8 (fewer shared-memory round trips; the pass is LSU-bound).
#define B200FFT_STRIDED_F64_PLANS(X)
#define B200FFT_STRIDED_F32_PLANS(X)

// alternative plans for A/B timing: X(n, variant, min CTAs per SM (0 = auto), tile row bytes (0 = auto), radices...)
#define B200FFT_ALT_PLANS(X)                                                              \
  X(1024, 1, 2, 0, 16, 8, 8) X(1024, 2, 0, 0, 8, 8, 4, 4) X(1024, 3, 0, 0, 4, 4, 8, 8)    \
  X(1024, 4, 0, 0, 8, 4, 4, 8) X(1024, 5, 1, 128, 4, 4, 8, 8) X(1024, 6, 3, 32, 4, 4, 8, 8) \
  X(1024, 7, 1, 128, 16, 8, 8) X(1536, 1, 0, 0, 4, 4, 8, 12) X(1536, 2, 0, 0, 8, 8, 2, 12)

// 2-CTA cluster plans of the strided pass for far strides (ClusterStridedK): X(n, radices of the n/2-point
// sub-transform each CTA runs after the cross stage).  The factor 3 stays in the last stage (fold).
#define B200FFT_CLUSTER_PLANS(X) \
  X(1024, 8, 8, 8) X(1536, 8, 8, 12) X(2048, 16, 8, 8) X(3072, 16, 8, 12)

// Fused z + y passes through L2 (fused_pair_kernel), compiled for the benchmark sizes: rows of 2*H reals
// next to columns of NY points.  X(H, NY, H-point row plan, NY-point column plan); the aliases repeat the
// radices of B200FFT_PLANS (template commas do not survive macro arguments).
namespace b200fft {
using FR256 = Plan<16, 16>;
using FR512 = Plan<8, 8, 8>;
using FR768 = Plan<8, 8, 12>;
using FC512 = Plan<8, 8, 8>;
using FC1024 = Plan<16, 8, 8>;
using FC1536 = Plan<16, 8, 12>;
}  // namespace b200fft
#define B200FFT_FUSED_PAIRS(X) X(256, 512, FR256, FC512) X(512, 1024, FR512, FC1024) X(768, 1536, FR768, FC1536)
