// Host stand-in for the CUDA runtime -- TEST INFRASTRUCTURE ONLY.
// Lets tests/emu build mpifft4py_b200/csrc/b200fft.cu (the C-ABI layer: plan objects, program execution,
// descriptor construction, timing records) with g++ so that its control flow runs in the `-m "not gpu"`
// suite: "device" memory is host memory, streams and events are inert (every launch executes at the call),
// kernel launches go to the CPU emulator (tests/emu/host_shim.cpp).  Never part of the product.
#pragma once
#include <chrono>
#include <cstdlib>
#include <cstring>

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorNotSupported = 801 };
typedef struct shim_stream* cudaStream_t;
struct shim_event { double t; long epoch; };
// One epoch per transform call of a rank (a thread): tests bump it through shim_next_epoch() before each call,
// and an event only counts as recorded within the call that waits for it.
inline long& shim_epoch() { static thread_local long e = 1; return e; }
typedef shim_event* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaIpcMemLazyEnablePeerAccess = 1 };
struct cudaIpcMemHandle_t { char reserved[64]; };

inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "shim: unsupported"; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::malloc(n ? n : 1); std::memset(*p, 0xff, n); return *p ? cudaSuccess : 2; }
template <class T> inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc(reinterpret_cast<void**>(p), n); }
inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = -1; return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = reinterpret_cast<cudaStream_t>(std::malloc(1)); return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { std::free(s); return cudaSuccess; }
// On the device a wait for an event that was never recorded is a silent no-op, i.e. a missing dependency:
// here it is an error, so that the host-shim tests catch it.
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t e, unsigned) { return (e && e->epoch == shim_epoch()) ? cudaSuccess : 1; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new shim_event{0, 0}; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) {
  if (!e) return 1;
  e->epoch = shim_epoch();
  e->t = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
  return cudaSuccess;
}
inline cudaError_t cudaEventSynchronize(cudaEvent_t e) { return e ? cudaSuccess : 1; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return cudaSuccess; }
// "IPC": the ranks of a host-shim test are threads of one process, so a handle is the pointer itself
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) { std::memset(h, 0, sizeof(*h)); std::memcpy(h->reserved, &p, sizeof(p)); return cudaSuccess; }
inline cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) { std::memcpy(p, h.reserved, sizeof(*p)); return *p ? cudaSuccess : cudaErrorNotSupported; }
inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
