// fused z + y passes through L2, double precision
#define REAL double
#define SUFFIX f64
#define B2_CAT_(a, b) a##b
#define B2_CAT(a, b) B2_CAT_(a, b)
#include "k_fused.inc"
