// fused z + y passes through L2, single precision
#define REAL float
#define SUFFIX f32
#define B2_CAT_(a, b) a##b
#define B2_CAT(a, b) B2_CAT_(a, b)
#include "k_fused.inc"
