// Kernel launcher shared by the per-precision translation units.
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <mutex>

#include "fft_kernels.cuh"
#include "fft_plans.h"

namespace b200fft {

constexpr int SMEM_LIMIT = 227 * 1024;

// One-time set-up of a kernel instantiation on the current device (shared-memory opt-in, resident CTA count
// of persistent kernels), safe against concurrent first launches and against processes that drive more than
// one device.  `kernel`: the __global__ function; returns 0 or a cudaError_t; *resident = CTAs the device holds.
template <class Kernel>
int configure_kernel(Kernel kernel, int threads, int smem, bool want_resident, std::mutex& mu, std::map<int, unsigned long long>& done,
                     unsigned long long* resident) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  std::lock_guard<std::mutex> lk(mu);
  auto it = done.find(dev);
  if (it == done.end()) {
    if (smem > 48 * 1024) {
      e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) return (int)e;
    }
    unsigned long long res = 0;
    if (want_resident) {
      int sms = 0, occ = 0;
      e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem);
      if (e != cudaSuccess) return (int)e;
      res = (unsigned long long)sms * (unsigned long long)(occ < 1 ? 1 : occ);
    }
    it = done.emplace(dev, res).first;
  }
  *resident = it->second;
  return 0;
}

template <class K>
int launch_k(const typename K::Params& p, cudaStream_t st) {
  if (K::SMEM > SMEM_LIMIT) return -1;
  static std::mutex mu;
  static std::map<int, unsigned long long> done;
  unsigned long long resident = 0;  // CTAs the device holds at once (persistent kernels)
  if (int rc = configure_kernel(fft_kernel<K>, K::NT, K::SMEM, K::PIPE, mu, done, &resident)) return rc;
  unsigned long long nblk = K::blocks(p);
  if (nblk == 0) return 0;
  if (nblk > 2147483647ull) return -2;
  if (K::PIPE && nblk > resident) nblk = resident;
  fft_kernel<K><<<(unsigned)nblk, K::NT, K::SMEM, st>>>(p);
  return (int)cudaGetLastError();
}

}  // namespace b200fft
