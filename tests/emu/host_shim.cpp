// libb200fft built for the HOST -- TEST INFRASTRUCTURE ONLY (never loaded by the product).
// The real C-ABI layer (mpifft4py_b200/csrc/b200fft.cu, compiled as C++ against tests/emu/cuda_shim) with
// its kernel launchers bound to the CPU emulator below.  What runs is the product's own plan objects,
// program execution loop, descriptor construction and timing records; what
// is emulated is the kernels (same phase bodies, tests/emu/emu.cpp) and the CUDA runtime (inert).
#include "emu.cpp"

#include "../../mpifft4py_b200/csrc/fft_dispatch.h"

namespace b200fft {
template <class real>
static int h_strided(int n, const StridedParams<real>& p) {
  switch (n) {
#define X(nn, ...) \
  case nn:         \
    return emulate<StridedK<real, Plan<__VA_ARGS__>>>(p);
    B200FFT_PLANS(X)
#undef X
    default:
      return -1;
  }
}
template <class real>
static int h_rowc2c(int n, const StridedParams<real>& p) {
  switch (n) {
#define X(nn, ...) \
  case nn:         \
    return emulate<RowC2CK<real, Plan<__VA_ARGS__>>>(p);
    B200FFT_PLANS(X)
#undef X
    default:
      return -1;
  }
}
template <class real, bool FWD>
static int h_rows(int h, const RowParams<real>& p) {
  switch (h) {
#define X(nn, ...)                                                   \
  case nn:                                                           \
    if (FWD) return emulate<R2CK<real, Plan<__VA_ARGS__>>>(p);       \
    else if (nn < 256 || nn > 1536) return emulate<C2RK<real, Plan<__VA_ARGS__>>>(p); \
    else return emulate<C2RDK<real, Plan<__VA_ARGS__>>>(p);
    B200FFT_ROW_PLANS(X)
#undef X
    default:
      return -1;
  }
}
// (the flag kernel of the peer-mapped transports: plain stores; the waits of the stand-in poll the same words)
int launch_post_flags(unsigned* const* words, int n, unsigned value, cudaStream_t) {
  for (int i = 0; i < n; ++i) __atomic_store_n(words[i], value, __ATOMIC_RELEASE);
  return 0;
}
int launch_strided_f64(int n, const StridedParams<double>& p, cudaStream_t) { return h_strided<double>(n, p); }
int launch_strided_f32(int n, const StridedParams<float>& p, cudaStream_t) { return h_strided<float>(n, p); }
int launch_rowc2c_f64(int n, const StridedParams<double>& p, cudaStream_t) { return h_rowc2c<double>(n, p); }
int launch_rowc2c_f32(int n, const StridedParams<float>& p, cudaStream_t) { return h_rowc2c<float>(n, p); }
int launch_r2c_f64(int h, const RowParams<double>& p, cudaStream_t) { return h_rows<double, true>(h, p); }
int launch_r2c_f32(int h, const RowParams<float>& p, cudaStream_t) { return h_rows<float, true>(h, p); }
int launch_c2r_f64(int h, const RowParams<double>& p, cudaStream_t) { return h_rows<double, false>(h, p); }
int launch_c2r_f32(int h, const RowParams<float>& p, cudaStream_t) { return h_rows<float, false>(h, p); }
}  // namespace b200fft

// ---- stand-ins for libcuda's stream memory operations and for NCCL, reached through the library's own
// dlopen / dlsym calls (interposed below): the ranks of a test are threads of this process ------------
#include <dlfcn.h>
#include <sched.h>

#include <chrono>
#include <condition_variable>
#include <map>
#include <mutex>
#include <tuple>

#include <cuda.h>
#include <nccl.h>

namespace shim {
// everything executes at the call, so a wait is a blocking poll on the word the peer thread writes;
// bounded, so that a protocol mistake fails a test instead of hanging it
static CUresult wait_value32(CUstream, CUdeviceptr addr, cuuint32_t value, unsigned) {
  const auto t0 = std::chrono::steady_clock::now();
  while (*reinterpret_cast<volatile cuuint32_t*>(addr) < value) {
    sched_yield();
    if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(20)) return 999;
  }
  return 0;
}
static CUresult write_value32(CUstream, CUdeviceptr addr, cuuint32_t value, unsigned) {
  *reinterpret_cast<volatile cuuint32_t*>(addr) = value;
  return 0;
}

struct Comm { std::string id; int nranks, rank; };
struct Msg { const void* src; size_t bytes; bool taken; };
static std::mutex mu;
static std::condition_variable cv;
static std::map<std::tuple<std::string, int, int>, Msg> box;  // (communicator id, from, to) -> pending send
static thread_local std::vector<std::tuple<ncclComm_t, int, const void*, size_t>> sends;
static thread_local std::vector<std::tuple<ncclComm_t, int, void*, size_t>> recvs;
static int next_id = 0;

static ncclResult_t get_unique_id(ncclUniqueId* id) {
  std::lock_guard<std::mutex> lk(mu);
  std::memset(id, 0, sizeof(*id));
  std::snprintf(id->internal, sizeof(id->internal), "shim-comm-%d", next_id++);
  return 0;
}
static ncclResult_t comm_init_rank(ncclComm_t* c, int nranks, ncclUniqueId id, int rank) {
  *c = reinterpret_cast<ncclComm_t>(new Comm{std::string(id.internal), nranks, rank});
  return 0;
}
static ncclResult_t comm_destroy(ncclComm_t c) { delete reinterpret_cast<Comm*>(c); return 0; }
static ncclResult_t group_start() { sends.clear(); recvs.clear(); return 0; }
static ncclResult_t send(const void* p, size_t n, ncclDataType_t, int peer, ncclComm_t c, shim_stream*) { sends.emplace_back(c, peer, p, n); return 0; }
static ncclResult_t recv(void* p, size_t n, ncclDataType_t, int peer, ncclComm_t c, shim_stream*) { recvs.emplace_back(c, peer, p, n); return 0; }
static ncclResult_t group_end() {  // post the sends, take the matching receives, wait until the sends were taken
  std::unique_lock<std::mutex> lk(mu);
  const auto deadline = std::chrono::steady_clock::now() + std::chrono::seconds(20);
  for (auto& s : sends) {
    Comm* c = reinterpret_cast<Comm*>(std::get<0>(s));
    auto key = std::make_tuple(c->id, c->rank, std::get<1>(s));
    if (!cv.wait_until(lk, deadline, [&] { return box.find(key) == box.end(); })) return 5;  // previous message still pending
    box[key] = Msg{std::get<2>(s), std::get<3>(s), false};
  }
  cv.notify_all();
  for (auto& r : recvs) {
    Comm* c = reinterpret_cast<Comm*>(std::get<0>(r));
    auto key = std::make_tuple(c->id, std::get<1>(r), c->rank);
    if (!cv.wait_until(lk, deadline, [&] { auto it = box.find(key); return it != box.end() && !it->second.taken; })) return 6;
    Msg& m = box[key];
    if (m.bytes != std::get<3>(r)) return 7;  // send / receive counts must agree
    std::memcpy(std::get<2>(r), m.src, m.bytes);
    m.taken = true;
  }
  cv.notify_all();
  for (auto& s : sends) {
    Comm* c = reinterpret_cast<Comm*>(std::get<0>(s));
    auto key = std::make_tuple(c->id, c->rank, std::get<1>(s));
    if (!cv.wait_until(lk, deadline, [&] { return box[key].taken; })) return 8;
    box.erase(key);
  }
  cv.notify_all();
  return 0;
}
static const char* error_string(ncclResult_t r) { return r == 0 ? "ok" : "shim NCCL: unmatched or timed-out send/recv"; }

static int nccl_handle, cuda_handle;
static void* open(const char* name, int) {
  if (std::strstr(name, "libnccl")) return &nccl_handle;
  if (std::strstr(name, "libcuda")) return &cuda_handle;
  return nullptr;
}
static void* sym(void* h, const char* name) {
  const std::string n(name);
  if (h == &cuda_handle) {
    if (n == "cuStreamWaitValue32_v2") return reinterpret_cast<void*>(&wait_value32);
    if (n == "cuStreamWriteValue32_v2") return reinterpret_cast<void*>(&write_value32);
    return nullptr;
  }
  if (n == "ncclGetUniqueId") return reinterpret_cast<void*>(&get_unique_id);
  if (n == "ncclCommInitRank") return reinterpret_cast<void*>(&comm_init_rank);
  if (n == "ncclCommDestroy") return reinterpret_cast<void*>(&comm_destroy);
  if (n == "ncclGroupStart") return reinterpret_cast<void*>(&group_start);
  if (n == "ncclGroupEnd") return reinterpret_cast<void*>(&group_end);
  if (n == "ncclSend") return reinterpret_cast<void*>(&send);
  if (n == "ncclRecv") return reinterpret_cast<void*>(&recv);
  if (n == "ncclGetErrorString") return reinterpret_cast<void*>(&error_string);
  return nullptr;
}
}  // namespace shim
#define dlopen shim::open
#define dlsym shim::sym

#include "../../mpifft4py_b200/csrc/b200fft.cu"

// test hook: start a new epoch for this rank's events (see cuda_shim/cuda_runtime.h)
extern "C" void shim_next_epoch() { ++shim_epoch(); }

// ---- the Navier-Stokes pointwise operations (csrc/ns_ops.cuh) as plain host loops behind the product's entry points ----
#include "../../mpifft4py_b200/csrc/ns_ops.cuh"
namespace {
template <class real>
b200fft::NsMesh<real> shim_mesh(const b200fft_ns_mesh_t& m) {
  return b200fft::NsMesh<real>{m.n0, m.n1, m.n2, (const real*)m.kx, (const real*)m.ky, (const real*)m.kz};
}
}  // namespace
extern "C" {
int b200fft_ns_curl(const b200fft_ns_mesh_t* m, const void* u, void* c, void*) {
  using namespace b200fft;
  const long long n = m->n0 * m->n1 * m->n2;
  for (long long i = 0; i < n; ++i) {
    if (m->precision == B200FFT_DOUBLE) ns_curl_point(shim_mesh<double>(*m), n, i, (const cx<double>*)u, (cx<double>*)c);
    else ns_curl_point(shim_mesh<float>(*m), n, i, (const cx<float>*)u, (cx<float>*)c);
  }
  return 0;
}
int b200fft_ns_cross(int precision, long long n, const void* a, const void* b, void* w, void*) {
  using namespace b200fft;
  for (long long i = 0; i < n; ++i) {
    if (precision == B200FFT_DOUBLE) ns_cross_point(n, i, (const double*)a, (const double*)b, (double*)w);
    else ns_cross_point(n, i, (const float*)a, (const float*)b, (float*)w);
  }
  return 0;
}
int b200fft_ns_rhs(const b200fft_ns_mesh_t* m, double nu, void* du, void* u, const void* u0, void* u1, double a_dt, double b_dt, int last,
                   void*) {
  using namespace b200fft;
  const long long n = m->n0 * m->n1 * m->n2;
  for (long long i = 0; i < n; ++i) {
    if (m->precision == B200FFT_DOUBLE)
      ns_rhs_point(shim_mesh<double>(*m), n, i, nu, (cx<double>*)du, (cx<double>*)u, (const cx<double>*)u0, (cx<double>*)u1, a_dt, b_dt, last);
    else
      ns_rhs_point(shim_mesh<float>(*m), n, i, (float)nu, (cx<float>*)du, (cx<float>*)u, (const cx<float>*)u0, (cx<float>*)u1, (float)a_dt,
                   (float)b_dt, last);
  }
  return 0;
}
}
