// Host stand-in for the CUDA driver API types used by b200fft.cu -- TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstdint>
typedef int CUresult;
typedef struct shim_stream* CUstream;
typedef unsigned long long CUdeviceptr;
typedef uint32_t cuuint32_t;
enum { CU_STREAM_WAIT_VALUE_GEQ = 0, CU_STREAM_WRITE_VALUE_DEFAULT = 0 };
