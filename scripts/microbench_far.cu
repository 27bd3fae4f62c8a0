// Tile-movement microbenchmark for the strided FFT pass (no butterflies): a CTA moves a tile of n rows x W bytes
// from [B][n][pitch] into shared memory, touches it once, and writes it back in place -- the HBM / address
// translation side of `StridedK` alone.  Question it answers (VERDICT r01 task 2): is the
// far-stride penalty of the x pass (rows 8.4 MB apart, one 2 MB page per row) a property of per-thread
// asynchronous copies (LDGSTS), and does the TMA unit (cp.async.bulk.tensor, one box per 256 rows) pay it too?
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/microbench_far scripts/microbench_far.cu
//   /tmp/microbench_far            # prints a table: mode x W x buffers x stride -> GB/s (read + write bytes)
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                                     \
  do {                                                                                            \
    cudaError_t e_ = (x);                                                                         \
    if (e_ != cudaSuccess) {                                                                      \
      std::fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));     \
      std::exit(2);                                                                               \
    }                                                                                             \
  } while (0)

struct Geo {
  long long B;      // batch entries
  int n;            // rows of a tile
  long long pitch;  // bytes between rows
  long long J;      // bytes per row that belong to the array (multiple of W)
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------- per-thread asynchronous copies (today's kernel)
template <int W>
__global__ void k_cpasync(char* base, Geo g, long long tiles, int touch) {
  extern __shared__ __align__(1024) unsigned char sm[];
  constexpr int T = W / 16;  // 16-byte columns
  const int c = threadIdx.x % T, t0 = threadIdx.x / T, TC = blockDim.x / T;
  const long long tpr = g.J / W;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long b = tile / tpr, j = tile % tpr;
    char* p = base + b * g.n * g.pitch + j * W + c * 16;
    for (int i = t0; i < g.n; i += TC) {
      const unsigned d = smem_u32(sm + (size_t)i * W + c * 16);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(p + (long long)i * g.pitch) : "memory");
    }
    asm volatile("cp.async.wait_all;\n" ::: "memory");
    __syncthreads();
    for (int i = t0; i < g.n; i += TC) {
      double2 v = *reinterpret_cast<double2*>(sm + (size_t)i * W + c * 16);
      if (touch) v.x += 1.0;
      *reinterpret_cast<double2*>(p + (long long)i * g.pitch) = v;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------- TMA boxes, NBUF tile buffers per CTA
__device__ __forceinline__ void mbar_init(unsigned a, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(n)); }
__device__ __forceinline__ void mbar_expect(unsigned a, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned a, unsigned parity) {
  unsigned ok = 0;
  while (!ok) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tma_load3(unsigned dst, const CUtensorMap* m, int c0, int c1, int c2, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(dst),
               "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tma_store3(const CUtensorMap* m, unsigned src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];\n" ::"l"(m), "r"(src), "r"(c0), "r"(c1),
               "r"(c2)
               : "memory");
}

template <int W, int NBUF>
__global__ void k_tma(const __grid_constant__ CUtensorMap tmap, Geo g, long long tiles, int touch, int boxrows) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ __align__(8) unsigned long long bars[NBUF];
  const long long tpr = g.J / W;
  const unsigned tile_bytes = (unsigned)g.n * W;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NBUF; ++i) mbar_init(smem_u32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](long long tile, int buf) {
    const long long b = tile / tpr, j = tile % tpr;
    const unsigned bar = smem_u32(&bars[buf]);
    mbar_expect(bar, tile_bytes);
    for (int r = 0; r < g.n; r += boxrows)
      tma_load3(smem_u32(sm + (size_t)buf * tile_bytes + (size_t)r * W), &tmap, (int)(j * (W / 8)), r, (int)b, bar);
  };
  long long k = 0;
  if (threadIdx.x == 0 && blockIdx.x < tiles) issue(blockIdx.x, 0);
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++k) {
    const int buf = (int)(k % NBUF);
    const long long next = tile + gridDim.x;
    if (NBUF > 1 && threadIdx.x == 0 && next < tiles) {
      // the buffer the next tile lands in was stored from NBUF-1 iterations ago: its reads must be done
      asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(NBUF - 2) : "memory");
      issue(next, (int)((k + 1) % NBUF));
    }
    mbar_wait(smem_u32(&bars[buf]), (unsigned)((k / NBUF) & 1));
    unsigned char* s = sm + (size_t)buf * tile_bytes;
    if (touch) {
      for (unsigned o = threadIdx.x * 16; o < tile_bytes; o += blockDim.x * 16) {
        double2 v = *reinterpret_cast<double2*>(s + o);
        v.x += 1.0;
        *reinterpret_cast<double2*>(s + o) = v;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      const long long b = tile / tpr, j = tile % tpr;
      for (int r = 0; r < g.n; r += boxrows) tma_store3(&tmap, smem_u32(s + (size_t)r * W), (int)(j * (W / 8)), r, (int)b);
      asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
      if (NBUF == 1) {
        asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
        if (next < tiles) issue(next, 0);
      }
    }
    if (NBUF == 1) __syncthreads();
  }
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn get_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  return (EncodeFn)fn;
}

static float time_ms(cudaEvent_t e0, cudaEvent_t e1, int reps) {
  float ms;
  CK(cudaEventSynchronize(e1));
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / reps;
}

template <int W>
static void run_cpasync(char* base, Geo g, int ctas_per_sm, int threads, int sms, const char* label) {
  const long long tiles = g.B * (g.J / W);
  const size_t smem = (size_t)g.n * W;
  CK(cudaFuncSetAttribute(k_cpasync<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_cpasync<W>, threads, smem));
  if (occ < ctas_per_sm) ctas_per_sm = occ;
  if (ctas_per_sm < 1) return;
  const int grid = (int)((long long)sms * ctas_per_sm < tiles ? (long long)sms * ctas_per_sm : tiles);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  k_cpasync<W><<<grid, threads, smem>>>(base, g, tiles, 1);
  CK(cudaDeviceSynchronize());
  const int reps = 3;
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) k_cpasync<W><<<grid, threads, smem>>>(base, g, tiles, 1);
  CK(cudaEventRecord(e1));
  const float ms = time_ms(e0, e1, reps);
  const double bytes = 2.0 * (double)g.B * g.n * (double)g.J;
  std::printf("%-10s cpasync  W=%3d  bufs=1 ctas/SM=%d thr=%4d  %8.3f ms  %7.1f GB/s\n", label, W, ctas_per_sm, threads, ms, bytes / ms / 1e6);
  std::fflush(stdout);
}

template <int W, int NBUF>
static void run_tma(char* base, Geo g, int ctas_per_sm, int threads, int sms, const char* label, int boxrows) {
  static EncodeFn enc = get_encode();
  CUtensorMap m;
  cuuint64_t dims[3] = {(cuuint64_t)(g.J / 8), (cuuint64_t)g.n, (cuuint64_t)g.B};
  cuuint64_t strides[2] = {(cuuint64_t)g.pitch, (cuuint64_t)g.pitch * g.n};
  cuuint32_t box[3] = {(cuuint32_t)(W / 8), (cuuint32_t)boxrows, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    std::printf("%-10s tma W=%d: cuTensorMapEncodeTiled failed (%d)\n", label, W, (int)r);
    return;
  }
  const long long tiles = g.B * (g.J / W);
  const size_t smem = (size_t)g.n * W * NBUF;
  if (smem > 227 * 1024) return;
  CK(cudaFuncSetAttribute(k_tma<W, NBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_tma<W, NBUF>, threads, smem));
  if (occ < ctas_per_sm) ctas_per_sm = occ;
  if (ctas_per_sm < 1) return;
  const int grid = (int)((long long)sms * ctas_per_sm < tiles ? (long long)sms * ctas_per_sm : tiles);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  k_tma<W, NBUF><<<grid, threads, smem>>>(m, g, tiles, 1, boxrows);
  CK(cudaDeviceSynchronize());
  const int reps = 3;
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) k_tma<W, NBUF><<<grid, threads, smem>>>(m, g, tiles, 1, boxrows);
  CK(cudaEventRecord(e1));
  const float ms = time_ms(e0, e1, reps);
  const double bytes = 2.0 * (double)g.B * g.n * (double)g.J;
  std::printf("%-10s tma      W=%3d  bufs=%d ctas/SM=%d thr=%4d box=%3d  %8.3f ms  %7.1f GB/s\n", label, W, NBUF, ctas_per_sm, threads, boxrows, ms,
              bytes / ms / 1e6);
  std::fflush(stdout);
}

int main(int argc, char** argv) {
  int n = argc > 1 ? std::atoi(argv[1]) : 1024;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  // the real arrays: slab 1024^3 double, complex (N0, N1, Nf = 513): x rows are N1*Nf*16 B apart, y rows Nf*16 B
  const long long Nf = 513, N1 = 1024, N0 = 1024;
  const size_t total = (size_t)N0 * N1 * Nf * 16 + (1 << 20);
  char* base;
  CK(cudaMalloc(&base, total));
  CK(cudaMemset(base, 0, total));
  struct Case {
    const char* label;
    Geo g;
  };
  std::vector<Case> cases;
  // far: B = 1, rows 8.4 MB apart; only the first 512 columns of each 513 so that every W divides the row
  cases.push_back({"far8.4MB", Geo{1, n, N1 * Nf * 16, (N1 * Nf * 16 / 256) * 256}});
  // near: the y pass -- B = N0 planes, rows 8208 B apart (J = 512 of the 513 columns)
  cases.push_back({"near8KB", Geo{N0 * 1024 / n, n, Nf * 16, 512 * 16}});
  // middle: rows 512 KB apart
  cases.push_back({"mid512KB", Geo{16 * 1024 / n, n, 512 * 1024, 512 * 1024}});
  std::printf("SMs %d, n = %d rows per tile\n", sms, n);
  for (const Case& c : cases) {
    run_cpasync<64>(base, c.g, 3, 256, sms, c.label);
    run_cpasync<128>(base, c.g, 1, 512, sms, c.label);
    run_cpasync<128>(base, c.g, 1, 1024, sms, c.label);
    if (n <= 512) run_cpasync<128>(base, c.g, 3, 256, sms, c.label);
    run_tma<64, 1>(base, c.g, 3, 256, sms, c.label, 256);
    run_tma<128, 1>(base, c.g, 1, 512, sms, c.label, 256);
    run_tma<128, 1>(base, c.g, 1, 512, sms, c.label, 64);
    run_tma<64, 2>(base, c.g, 1, 512, sms, c.label, 256);
    run_tma<64, 3>(base, c.g, 1, 512, sms, c.label, 256);
    if (n <= 512) {
      run_tma<128, 1>(base, c.g, 3, 256, sms, c.label, 256);
      run_tma<128, 2>(base, c.g, 1, 512, sms, c.label, 256);
      run_tma<128, 3>(base, c.g, 1, 512, sms, c.label, 256);
      run_tma<256, 1>(base, c.g, 1, 512, sms, c.label, 256);
    }
  }
  return 0;
}
