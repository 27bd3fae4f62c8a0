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

template <class K, bool RB = false>
int launch_k(const typename K::Params& p, cudaStream_t st) {
  if (K::SMEM > SMEM_LIMIT) return -1;
  static std::mutex mu;
  static std::map<int, unsigned long long> done;
  unsigned long long resident = 0;  // CTAs the device holds at once (persistent kernels)
  if (int rc = configure_kernel(fft_kernel<K, RB>, K::NT, K::SMEM, K::PIPE, mu, done, &resident)) return rc;
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
  static std::mutex mu;
  static std::map<int, unsigned long long> done;
  unsigned long long unused = 0;
  if (int rc = configure_kernel(fft_cluster_kernel<K>, K::NT, K::SMEM, false, mu, done, &unused)) return rc;
  const unsigned long long nblk = K::blocks(p);
  if (nblk == 0) return 0;
  if (nblk > 2147483647ull) return -2;
  fft_cluster_kernel<K><<<(unsigned)nblk, K::NT, K::SMEM, st>>>(p);
  return (int)cudaGetLastError();
}

// Fused pair of passes (fused_pair_kernel) over `planes` planes in groups of `planes_per_group`.  `ctl`
// points at FUSE_CTL_WORDS words of device memory (queue head + per-group counters); -3 when there would
// be more groups than that or a group smaller than one block.

template <class KA, class KB>
int launch_fused_pair(const typename KA::Params& pa, const typename KB::Params& pb, long long planes, long long planes_per_group,
                      unsigned* ctl, cudaStream_t st) {
  using F = FusePair<KA, KB>;
  if (F::SMEM > SMEM_LIMIT) return -1;
  if (planes < 1 || planes_per_group < 1) return -3;
  const long long groups = (planes + planes_per_group - 1) / planes_per_group;
  FuseCtl c;
  c.a = KA::fuse_side(pa, planes, planes_per_group);
  c.b = KB::fuse_side(pb, planes, planes_per_group);
  if (groups + 1 > FUSE_CTL_WORDS || c.a.n == 0 || c.b.n == 0 || c.a.upg < c.a.upb || c.b.upg < c.b.upb ||
      (unsigned long long)c.a.n + c.b.n > 2147483647ull)
    return -3;
  static std::mutex mu;
  static std::map<int, unsigned long long> done;
  unsigned long long resident = 0;
  if (int rc = configure_kernel(fused_pair_kernel<KA, KB>, F::NT, F::SMEM, true, mu, done, &resident)) return rc;
  cudaError_t e = cudaMemsetAsync(ctl, 0, sizeof(unsigned) * (size_t)(1 + groups), st);
  if (e != cudaSuccess) return (int)e;
  c.ctr = ctl;
  c.done = ctl + 1;
  c.G = (unsigned)groups;
  // every CTA of the grid must be resident (a waiting block relies on the blocks before it making progress)
  const unsigned long long total = (unsigned long long)c.a.n + c.b.n;
  const unsigned long long grid = total < resident ? total : resident;
  fused_pair_kernel<KA, KB><<<(unsigned)grid, F::NT, F::SMEM, st>>>(pa, pb, c);
  return (int)cudaGetLastError();
}

}  // namespace b200fft
