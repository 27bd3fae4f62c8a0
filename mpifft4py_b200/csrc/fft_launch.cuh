// Kernel launcher shared by the per-precision translation units.
#pragma once
#include <cuda_runtime.h>
#include "fft_kernels.cuh"
#include "fft_plans.h"

namespace b200fft {

constexpr int SMEM_LIMIT = 227 * 1024;

template <class K, bool RB = false>
int launch_k(const typename K::Params& p, cudaStream_t st) {
  if (K::SMEM > SMEM_LIMIT) return -1;
  static bool configured = false;
  static unsigned long long resident = 0;  // CTAs the device holds at once (persistent kernels)
  if (!configured) {
    if (K::SMEM > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(fft_kernel<K, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM);
      if (e != cudaSuccess) return (int)e;
    }
    if (K::PIPE) {
      int dev = 0, sms = 0, occ = 0;
      cudaError_t e = cudaGetDevice(&dev);
      if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fft_kernel<K, RB>, K::NT, K::SMEM);
      if (e != cudaSuccess) return (int)e;
      resident = (unsigned long long)sms * (unsigned long long)(occ < 1 ? 1 : occ);
    }
    configured = true;
  }
  unsigned long long nblk = K::blocks(p);
  if (nblk == 0) return 0;
  if (nblk > 2147483647ull) return -2;
  if (K::PIPE && nblk > resident) nblk = resident;
  fft_kernel<K, RB><<<(unsigned)nblk, K::NT, K::SMEM, st>>>(p);
  return (int)cudaGetLastError();
}

// 2-CTA cluster kernels (ClusterStridedK): the cluster shape is a compile-time attribute of the kernel,
// the grid holds two CTAs per column tile
template <class K>
int launch_cluster_k(const typename K::Params& p, cudaStream_t st) {
  if (K::SMEM > SMEM_LIMIT) return -1;
  static bool configured = false;
  if (!configured) {
    if (K::SMEM > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(fft_cluster_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM);
      if (e != cudaSuccess) return (int)e;
    }
    configured = true;
  }
  const unsigned long long nblk = K::blocks(p);
  if (nblk == 0) return 0;
  if (nblk > 2147483647ull) return -2;
  fft_cluster_kernel<K><<<(unsigned)nblk, K::NT, K::SMEM, st>>>(p);
  return (int)cudaGetLastError();
}

}  // namespace b200fft
