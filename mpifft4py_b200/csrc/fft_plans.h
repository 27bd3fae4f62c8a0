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
