// Flag words of the peer-mapped transports (b200fft.cu: run_exchange_p2p, credits): ONE single-warp kernel writes the
// sequence number into the flag word of every peer, ordered by the stream behind the data copies / the last reader.
// It replaces 2 stream operations per peer (a stream memory op into a staging word + a 4-byte DMA across NVLink):
// at 8 GPUs an exchange step queued 7 data copies and 14 flag operations, each with its own issue latency.
#include <cuda_runtime.h>

#include "fft_dispatch.h"

namespace b200fft {
namespace {
struct PeerWords {
  unsigned* w[16];
  unsigned value;
  int n;
};
__global__ void post_flags_kernel(PeerWords p) {
  if ((int)threadIdx.x < p.n) {
    __threadfence_system();  // (stream order already put the copies before this kernel; the fence orders this thread's view)
    *reinterpret_cast<volatile unsigned*>(p.w[threadIdx.x]) = p.value;
    __threadfence_system();
  }
}
}  // namespace

int launch_post_flags(unsigned* const* words, int n, unsigned value, cudaStream_t st) {
  if (n < 1) return 0;
  if (n > 16) return -1;
  PeerWords p;
  for (int i = 0; i < n; ++i) p.w[i] = words[i];
  p.value = value;
  p.n = n;
  post_flags_kernel<<<1, 32, 0, st>>>(p);
  return (int)cudaGetLastError();
}
}  // namespace b200fft
