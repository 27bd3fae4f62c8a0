// contiguous-row C2C kernels, double precision (sm_100a)
#define REAL double
#define SUFFIX f64
#define B2_CAT_(a, b) a##b
#define B2_CAT(a, b) B2_CAT_(a, b)
#include "k_rowc2c.inc"
