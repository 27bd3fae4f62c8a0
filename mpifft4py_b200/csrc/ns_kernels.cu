// Elementwise kernels of the pseudo-spectral Navier-Stokes step (include/b200fft.h: b200fft_ns_*; bodies in
// ns_ops.cuh).  HBM-bound streaming passes: persistent grid-stride loops, one CTA wave of 148 SMs x 8, consecutive
// threads on consecutive points (the three components of a point sit n elements apart, each stream coalesced).
#include <cuda_runtime.h>

#include "../../include/b200fft.h"
#include "ns_ops.cuh"

namespace b200fft {
int ns_fail(int code, const char* msg);  // b200fft.cu: records the message for b200fft_last_error
}

namespace {
using namespace b200fft;

template <class real>
__global__ void __launch_bounds__(256) ns_curl_kernel(NsMesh<real> m, long long n, const cx<real>* u, cx<real>* c) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    ns_curl_point(m, n, i, u, c);
}
template <class real>
__global__ void __launch_bounds__(256) ns_cross_kernel(long long n, const real* a, const real* b, real* w) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    ns_cross_point(n, i, a, b, w);
}
template <class real>
__global__ void __launch_bounds__(256) ns_rhs_kernel(NsMesh<real> m, long long n, real nu, cx<real>* du, cx<real>* u, const cx<real>* u0,
                                                     cx<real>* u1, real a_dt, real b_dt, int last) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    ns_rhs_point(m, n, i, nu, du, u, u0, u1, a_dt, b_dt, last);
}

int grid_for(long long n) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long want = (n + 255) / 256, cap = (long long)sms * 8;
  return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

template <class real>
NsMesh<real> mesh_of(const b200fft_ns_mesh_t& m) {
  return NsMesh<real>{m.n0, m.n1, m.n2, (const real*)m.kx, (const real*)m.ky, (const real*)m.kz};
}

int check_mesh(const b200fft_ns_mesh_t* m) {
  if (!m || (m->precision != B200FFT_SINGLE && m->precision != B200FFT_DOUBLE) || m->n0 < 1 || m->n1 < 1 || m->n2 < 1 || !m->kx ||
      !m->ky || !m->kz)
    return ns_fail(B200FFT_ERR_ARG, "ns: bad mesh descriptor");
  return 0;
}

int launched() {
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : ns_fail(B200FFT_ERR_CUDA, cudaGetErrorString(e));
}
}  // namespace

extern "C" {

int b200fft_ns_curl(const b200fft_ns_mesh_t* m, const void* u_hat, void* curl_hat, void* stream) {
  if (int rc = check_mesh(m)) return rc;
  if (!u_hat || !curl_hat) return ns_fail(B200FFT_ERR_ARG, "ns_curl: null array");
  const long long n = m->n0 * m->n1 * m->n2;
  cudaStream_t st = (cudaStream_t)stream;
  if (m->precision == B200FFT_DOUBLE)
    ns_curl_kernel<double><<<grid_for(n), 256, 0, st>>>(mesh_of<double>(*m), n, (const cx<double>*)u_hat, (cx<double>*)curl_hat);
  else
    ns_curl_kernel<float><<<grid_for(n), 256, 0, st>>>(mesh_of<float>(*m), n, (const cx<float>*)u_hat, (cx<float>*)curl_hat);
  return launched();
}

int b200fft_ns_cross(int precision, long long npoints, const void* a, const void* b, void* out, void* stream) {
  if ((precision != B200FFT_SINGLE && precision != B200FFT_DOUBLE) || npoints < 1 || !a || !b || !out)
    return ns_fail(B200FFT_ERR_ARG, "ns_cross: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == B200FFT_DOUBLE)
    ns_cross_kernel<double><<<grid_for(npoints), 256, 0, st>>>(npoints, (const double*)a, (const double*)b, (double*)out);
  else
    ns_cross_kernel<float><<<grid_for(npoints), 256, 0, st>>>(npoints, (const float*)a, (const float*)b, (float*)out);
  return launched();
}

int b200fft_ns_rhs(const b200fft_ns_mesh_t* m, double nu, void* du, void* u_hat, const void* u_hat0, void* u_hat1, double a_dt,
                   double b_dt, int last, void* stream) {
  if (int rc = check_mesh(m)) return rc;
  if (!du || !u_hat || ((u_hat0 == nullptr) != (u_hat1 == nullptr))) return ns_fail(B200FFT_ERR_ARG, "ns_rhs: bad arguments");
  const long long n = m->n0 * m->n1 * m->n2;
  cudaStream_t st = (cudaStream_t)stream;
  if (m->precision == B200FFT_DOUBLE)
    ns_rhs_kernel<double><<<grid_for(n), 256, 0, st>>>(mesh_of<double>(*m), n, nu, (cx<double>*)du, (cx<double>*)u_hat,
                                                       (const cx<double>*)u_hat0, (cx<double>*)u_hat1, a_dt, b_dt, last);
  else
    ns_rhs_kernel<float><<<grid_for(n), 256, 0, st>>>(mesh_of<float>(*m), n, (float)nu, (cx<float>*)du, (cx<float>*)u_hat,
                                                      (const cx<float>*)u_hat0, (cx<float>*)u_hat1, (float)a_dt, (float)b_dt, last);
  return launched();
}

}  // extern "C"
