// libb200fft.so -- C ABI (include/b200fft.h): fused FFT passes, NCCL communicators and the
// distributed slab / pencil / line R2C plans that replace mpiFFT4py's hot path.
//
// A plan is a short program of steps (row R2C/C2R pass, strided C2C pass, exchange) over five
// buffers: the caller's input and output and up to three device work buffers owned by the plan
// (the reference's work_arrays, mpibase.py:61-131).  Every copy the reference does between its
// FFT calls (pack, transpose, pad, truncate, mask, scale) is part of a pass's index map; the MPI
// collectives become one NCCL send/recv group per exchange, with each rank's own block written
// straight into the receive buffer by the producing FFT pass.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <chrono>
#include <thread>
#include <vector>

#include "../../include/b200fft.h"
#include "desc_convert.h"
#include "fft_dispatch.h"
#include "fft_plans.h"
#include "plan_program.h"

using namespace b200fft;

namespace b200fft {
bool plan_exists(int n) {
  switch (n) {
#define X(nn, ...) case nn:
    B200FFT_PLANS(X)
#undef X
    return true;
    default:
      return false;
  }
}
int ns_fail(int code, const char* msg) { return fail(code, "%s", msg); }  // ns_kernels.cu
}  // namespace b200fft

namespace {

#define g_err (b200fft::plan_err())
using b200fft::fail;

int cuda_fail(cudaError_t e, const char* what) {
  return fail(B200FFT_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

// ---- twiddle tables (device), one per (device, length, precision) ------------------------------
std::mutex g_tw_mu;
std::map<std::tuple<int, int, int>, void*> g_tw;

template <class real>
int get_tw(int len, const cx<real>** out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  std::lock_guard<std::mutex> lk(g_tw_mu);
  auto key = std::make_tuple(dev, len, (int)sizeof(real));
  auto it = g_tw.find(key);
  if (it == g_tw.end()) {
    std::vector<cx<real>> h = make_twiddles<real>(len);
    void* d = nullptr;
    e = cudaMalloc(&d, h.size() * sizeof(cx<real>));
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(twiddles)");
    e = cudaMemcpy(d, h.data(), h.size() * sizeof(cx<real>), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy(twiddles)");
    it = g_tw.emplace(key, d).first;
  }
  *out = reinterpret_cast<const cx<real>*>(it->second);
  return 0;
}

int map_launch_rc(int rc, const char* what, int n) {
  if (rc == 0) return 0;
  if (rc == -1) return fail(B200FFT_ERR_UNSUPPORTED, "%s: no kernel plan for length %d (supported: 2^k, 3*2^k that fit shared memory)", what, n);
  if (rc == -2) return fail(B200FFT_ERR_ARG, "%s: grid too large", what);
  return cuda_fail((cudaError_t)rc, what);
}

int exec_strided(const b200fft_strided_desc_t& d, cudaStream_t st) {
  if (const char* e = check_strided(d)) return fail(B200FFT_ERR_ARG, "strided pass: %s", e);
  if (d.B == 0 || d.J == 0) return 0;
  const bool rows = contiguous_rows(d);  // J == 1, unit stride: threads walk along the row instead
  if (d.precision == B200FFT_DOUBLE) {
    const cx<double>* tw;
    if (int rc = get_tw<double>(d.n, &tw)) return rc;
    auto p = convert_strided<double>(d, tw, 1);
    if (d.cross_n > 0) {
      if (int rc = get_tw<double>(d.cross_n, &p.tw2)) return rc;
      p.tw2_div = d.cross_div;
    }
    return map_launch_rc(rows ? launch_rowc2c_f64(d.n, p, st) : launch_strided_f64(d.n, p, st), "strided pass", d.n);
  }
  const cx<float>* tw;
  if (int rc = get_tw<float>(d.n, &tw)) return rc;
  auto p = convert_strided<float>(d, tw, 1);
  if (d.cross_n > 0) {
    if (int rc = get_tw<float>(d.cross_n, &p.tw2)) return rc;
    p.tw2_div = d.cross_div;
  }
  return map_launch_rc(rows ? launch_rowc2c_f32(d.n, p, st) : launch_strided_f32(d.n, p, st), "strided pass", d.n);
}

int exec_rows(const b200fft_rows_desc_t& d, bool fwd, cudaStream_t st) {
  if (const char* e = check_rows(d)) return fail(B200FFT_ERR_ARG, "row pass: %s", e);
  if (d.rows == 0) return 0;
  const int h = d.n / 2;
  int rc;
  if (d.precision == B200FFT_DOUBLE) {
    const cx<double>* tw;
    if (int r = get_tw<double>(d.n, &tw)) return r;
    auto p = convert_rows<double>(d, tw, 1, fwd);
    rc = fwd ? launch_r2c_f64(h, p, st) : launch_c2r_f64(h, p, st);
  } else {
    const cx<float>* tw;
    if (int r = get_tw<float>(d.n, &tw)) return r;
    auto p = convert_rows<float>(d, tw, 1, fwd);
    rc = fwd ? launch_r2c_f32(h, p, st) : launch_c2r_f32(h, p, st);
  }
  return map_launch_rc(rc, fwd ? "R2C pass" : "C2R pass", d.n);
}

// ---- NCCL through dlopen: the library loads (and does P == 1 work) without NCCL ---------------
struct NcclApi {
  void* h = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

int load_nccl() {
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if (g_nccl.h) return 0;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) return fail(B200FFT_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define L(sym)                                                                         \
  g_nccl.sym = reinterpret_cast<decltype(g_nccl.sym)>(dlsym(h, "nccl" #sym));          \
  if (!g_nccl.sym) return fail(B200FFT_ERR_NCCL, "libnccl lacks nccl" #sym);
  L(GetUniqueId) L(CommInitRank) L(CommDestroy) L(GroupStart) L(GroupEnd) L(Send) L(Recv) L(GetErrorString)
#undef L
  g_nccl.h = h;
  return 0;
}

int nccl_fail(ncclResult_t r, const char* what) {
  return fail(B200FFT_ERR_NCCL, "%s: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
}

// ---- driver API stream memory operations (copy-engine transport), through dlopen -----------------
struct CuApi {
  void* h = nullptr;
  CUresult (*WaitValue32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int) = nullptr;
  CUresult (*WriteValue32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int) = nullptr;
};
CuApi g_cu;
std::mutex g_cu_mu;

int load_cuda_driver() {
  std::lock_guard<std::mutex> lk(g_cu_mu);
  if (g_cu.h) return 0;
  void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return fail(B200FFT_ERR_CUDA, "cannot dlopen libcuda.so.1: %s", dlerror());
  auto sym = [&](const char* a, const char* b) {
    void* f = dlsym(h, a);
    return f ? f : dlsym(h, b);
  };
  g_cu.WaitValue32 = reinterpret_cast<decltype(g_cu.WaitValue32)>(sym("cuStreamWaitValue32_v2", "cuStreamWaitValue32"));
  g_cu.WriteValue32 = reinterpret_cast<decltype(g_cu.WriteValue32)>(sym("cuStreamWriteValue32_v2", "cuStreamWriteValue32"));
  if (!g_cu.WaitValue32 || !g_cu.WriteValue32) return fail(B200FFT_ERR_CUDA, "libcuda lacks cuStreamWaitValue32 / cuStreamWriteValue32");
  g_cu.h = h;
  return 0;
}

}  // namespace

constexpr int P2P_STAGE = 512;  // staging words for flag values (ring)

struct b200fft_comm {
  ncclComm_t comm;
  int nranks;
  int rank;
};

struct b200fft_plan {
  b200fft_plan_desc_t d;
  Program prog[2][3];  // [forward=0 / inverse=1][dealias]
  void* ws[NWORK] = {};
  size_t wbytes[NWORK] = {};
  int last_kernels = 0, last_exch = 0;
  int timing = 0;
  // copy-engine (P2P) transport: peers' work buffers and flag words mapped through CUDA IPC.
  // Blocks are pushed into the peer's receive buffer by cudaMemcpyAsync (DMA over NVLink, no SMs);
  // arrival and buffer-reuse credits are 32-bit sequence numbers written / awaited by stream
  // memory operations, so nothing blocks the host and no kernel spins.
  struct {
    bool connected = false;
    void* flags = nullptr;                   // uint32 arrived[MAXP] then credit[MAXP]
    void* peer_ws[B200FFT_MAXP][NPEERBUF] = {};
    void* peer_flags[B200FFT_MAXP] = {};
    cudaStream_t wait_stream = nullptr;
    std::vector<cudaEvent_t> send_ev;
    unsigned stage_slot = 0;                 // next staging word (flag values travel by 4-byte DMA)
    unsigned seq = 0;                        // exchange steps executed so far (identical on all ranks)
    unsigned calls = 0;                      // transforms with exchanges executed so far
  } p2p;
  cudaStream_t comm_stream = nullptr;   // exchanges of pipelined programs run here
  std::vector<cudaEvent_t> sched_ev;    // ordering events between the two streams
  std::vector<cudaEvent_t> ev;  // timing events, two per step (grown on demand)
  float last_fft_ms = -1.f, last_exch_ms = -1.f;
  std::vector<std::pair<int, int>> ev_marks;  // (event index start, is_exchange)
  std::vector<int> st_type, st_len, st_pass;
  std::vector<double> st_bytes;
};

namespace {

bool is_pow2(long long x) { return x > 0 && (x & (x - 1)) == 0; }

int validate_desc(const b200fft_plan_desc_t& d) {
  if (d.precision != B200FFT_SINGLE && d.precision != B200FFT_DOUBLE) return fail(B200FFT_ERR_ARG, "precision must be single or double");
  if (d.nranks < 1 || d.rank < 0 || d.rank >= d.nranks) return fail(B200FFT_ERR_ARG, "bad rank / nranks");
  const int dims = d.kind == B200FFT_LINE ? 2 : 3;
  for (int i = 0; i < dims; ++i)
    if (d.N[i] < 4 || d.N[i] % 2) return fail(B200FFT_ERR_ARG, "N[%d]=%lld: mesh sizes must be even and >= 4", i, d.N[i]);
  const int P = d.nranks;
  if (d.kind == B200FFT_SLAB || d.kind == B200FFT_SLAB_C2C) {
    if (!is_pow2(P) || P > d.N[0])  // slab.py:89-91
      return fail(B200FFT_ERR_RANKS, "Number of cpus must be a power of two <= N[0]");
    if (d.N[0] % P || d.N[1] % P) return fail(B200FFT_ERR_ARG, "N[0], N[1] must be divisible by the number of ranks");
    if (P > B200FFT_MAXP) return fail(B200FFT_ERR_RANKS, "at most %d ranks", B200FFT_MAXP);
  } else if (d.kind == B200FFT_LINE) {
    if (d.N[0] % P || d.N[1] % (2 * P)) return fail(B200FFT_ERR_ARG, "N[0], N[1]/2 must be divisible by the number of ranks");
    if (P > B200FFT_MAXP) return fail(B200FFT_ERR_RANKS, "at most %d ranks", B200FFT_MAXP);
  } else if (d.kind == B200FFT_PENCIL_X || d.kind == B200FFT_PENCIL_Y) {
    if (P < 2) return fail(B200FFT_ERR_ARG, "pencil decomposition needs more than one rank");  // pencil.py:176
    if (P % 2) return fail(B200FFT_ERR_RANKS, "Number of cpus must be even");                    // pencil.py:201-202
    if (d.P1 < 1 || d.P2 < 1 || d.P1 * d.P2 != P) return fail(B200FFT_ERR_ARG, "P1*P2 must equal the number of ranks");
    if ((d.P1 % 2) || (d.P2 % 2))  // pencil.py:204-205
      return fail(B200FFT_ERR_RANKS, "Number of cpus in each direction must be even power of 2");
    if (d.P1 > B200FFT_MAXP || d.P2 > B200FFT_MAXP) return fail(B200FFT_ERR_RANKS, "at most %d ranks per direction", B200FFT_MAXP);
    for (int i = 0; i < 3; ++i)
      if (d.N[i] % d.P1 || d.N[i] % d.P2) return fail(B200FFT_ERR_ARG, "N must be divisible by P1 and P2");
    const int zparts = d.kind == B200FFT_PENCIL_X ? d.P2 : d.P1;
    if ((d.N[2] / 2) % zparts) return fail(B200FFT_ERR_ARG, "N[2]/2 must be divisible by the z process count");
  } else {
    return fail(B200FFT_ERR_ARG, "unknown plan kind %d", d.kind);
  }
  if (P > 1) {
    if (d.transport == B200FFT_TRANSPORT_P2P || d.transport == B200FFT_TRANSPORT_STORE) {
      // no communicator: peers are reached through IPC-mapped buffers (b200fft_plan_p2p_connect)
    } else if (d.kind == B200FFT_SLAB || d.kind == B200FFT_SLAB_C2C || d.kind == B200FFT_LINE) {
      if (!d.comm) return fail(B200FFT_ERR_ARG, "multi-rank plan needs a communicator");
      if (d.comm->nranks != P || d.comm->rank != d.rank) return fail(B200FFT_ERR_ARG, "communicator does not match nranks / rank");
    } else {
      if (!d.comm0 || !d.comm1) return fail(B200FFT_ERR_ARG, "pencil plan needs comm0 and comm1");
      if (d.comm0->nranks != d.P1 || d.comm0->rank != d.rank % d.P1) return fail(B200FFT_ERR_ARG, "comm0 must hold the P1 ranks with equal rank / P1");
      if (d.comm1->nranks != d.P2 || d.comm1->rank != d.rank / d.P1) return fail(B200FFT_ERR_ARG, "comm1 must hold the P2 ranks with equal rank %% P1");
    }
  }
  return 0;
}

// lengths used by a program must have kernels
int check_lengths(const Program& pg) {
  for (const Step& s : pg.steps) {
    if (s.type == ST_STRIDED && !plan_exists(s.n))
      return fail(B200FFT_ERR_UNSUPPORTED, "no kernel plan for complex length %d", s.n);
    if ((s.type == ST_R2C || s.type == ST_C2R) && (s.n % 2 || !plan_exists(s.n / 2)))
      return fail(B200FFT_ERR_UNSUPPORTED, "no kernel plan for real length %d", s.n);
  }
  return 0;
}

int ensure_program(b200fft_plan* pl, int inverse, int dealias) {
  Program& pg = pl->prog[inverse][dealias];
  if (pg.built) {
    if (pg.error) return fail(pg.error, "%s", pg.errmsg.c_str());
    return 0;
  }
  pg.built = true;
  if (dealias == B200FFT_DEALIAS_3_2) {
    const int dims = pl->d.kind == B200FFT_LINE ? 2 : 3;
    bool ok = std::fabs(pl->d.padsize - 1.5) < 1e-12;
    for (int i = 0; i < dims; ++i) ok = ok && (pl->d.N[i] % 2 == 0);
    if (!ok) {
      pg.error = fail(B200FFT_ERR_UNSUPPORTED, "3/2-rule kernels exist for padsize == 1.5 only (got %g)", pl->d.padsize);
      pg.errmsg = g_err;
      return pg.error;
    }
  }
  int rc = build_program(pl->d, inverse, dealias, pg);
  if (!rc) rc = check_lengths(pg);
  if (rc) {
    pg.error = rc;
    pg.errmsg = g_err;
    pg.steps.clear();
    return rc;
  }
  // grow the work buffers
  const size_t csz = pl->d.precision == B200FFT_DOUBLE ? 16 : 8;
  for (int w = 0; w < NWORK; ++w) {
    const size_t need = (size_t)pg.need[BUF_W0 + w] * csz;
    if (need > pl->wbytes[w]) {
      // (every early return below leaves the program unbuilt, so that a later call cannot run it with
      // undersized work buffers)
      if (pl->p2p.connected) {
        pg.built = false;
        return fail(B200FFT_ERR_NOMEM, "P2P plans size their buffers at connect time (need %zu > %zu)", need, pl->wbytes[w]);
      }
      if (pl->ws[w]) {
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          pg.built = false;
          return cuda_fail(e, "cudaDeviceSynchronize");
        }
        cudaFree(pl->ws[w]);
        pl->ws[w] = nullptr;
        pl->wbytes[w] = 0;
      }
      cudaError_t e = cudaMalloc(&pl->ws[w], need);
      if (e != cudaSuccess) {
        pg.built = false;
        return fail(B200FFT_ERR_NOMEM, "cudaMalloc(%zu bytes of work space): %s", need, cudaGetErrorString(e));
      }
      pl->wbytes[w] = need;
    }
  }
  return 0;
}

void* resolve(const b200fft_plan* pl, const Ref& r, const void* in, void* out, size_t esz) {
  char* base;
  switch (r.buf) {
    case BUF_IN: base = (char*)const_cast<void*>(in); break;
    case BUF_OUT: base = (char*)out; break;
    default:  // r.peer >= 0: that rank's work buffer through its IPC mapping (fused transport)
      base = (char*)((r.peer >= 0 && r.peer != pl->d.rank) ? pl->p2p.peer_ws[r.peer][r.buf - BUF_W0] : pl->ws[r.buf - BUF_W0]);
      break;
  }
  return base + (size_t)r.off * esz;
}

void fill_side(const b200fft_plan* pl, const SideT& s, b200fft_side_t& o, const void* in, void* out, size_t csz) {
  std::memset(&o, 0, sizeof(o));
  for (int q = 0; q < s.nchunk; ++q) {
    o.base[q] = resolve(pl, s.base[q], in, out, csz);
    o.sb[q] = s.sb[q];
    o.si[q] = s.si[q];
  }
  o.chunk = s.chunk;
  o.nchunk = s.nchunk;
  o.nphys = s.nphys;
}

int run_exchange(b200fft_plan* pl, const Step& s, const void* in, void* out, size_t csz, cudaStream_t st) {
  b200fft_comm_t c = s.comm == 0 ? pl->d.comm : (s.comm == 1 ? pl->d.comm0 : pl->d.comm1);
  if (!c) return fail(B200FFT_ERR_ARG, "exchange without communicator");
  ncclResult_t r = g_nccl.GroupStart();
  if (r != ncclSuccess) return nccl_fail(r, "ncclGroupStart");
  for (int q = 0; q < s.npeers; ++q) {
    if (q == s.me) continue;  // own block was written in place by the producing pass
    r = g_nccl.Send(resolve(pl, s.send[q], in, out, csz), (size_t)s.scnt[q] * csz, ncclChar, q, c->comm, st);
    if (r != ncclSuccess) return nccl_fail(r, "ncclSend");
    r = g_nccl.Recv(resolve(pl, s.recv[q], in, out, csz), (size_t)s.rcnt[q] * csz, ncclChar, q, c->comm, st);
    if (r != ncclSuccess) return nccl_fail(r, "ncclRecv");
  }
  r = g_nccl.GroupEnd();
  if (r != ncclSuccess) return nccl_fail(r, "ncclGroupEnd");
  return 0;
}

int cu_fail(CUresult r, const char* what) { return fail(B200FFT_ERR_CUDA, "%s: CUresult %d", what, (int)r); }

// Publish `value` in a peer's flag word, ordered after everything queued on `st` so far: a stream
// memory operation writes it into a local staging word and a 4-byte DMA copy carries it across
// NVLink (stream order = the data copies before it have completed).
int post_flag(b200fft_plan* pl, cudaStream_t st, void* peer_word, unsigned value) {
  auto& pp = pl->p2p;
  unsigned* stage = reinterpret_cast<unsigned*>(pp.flags) + 2 * B200FFT_MAXP + (pp.stage_slot++ % P2P_STAGE);
  if (CUresult r = g_cu.WriteValue32((CUstream)st, (CUdeviceptr)stage, value, CU_STREAM_WRITE_VALUE_DEFAULT))
    return cu_fail(r, "cuStreamWriteValue32");
  cudaError_t e = cudaMemcpyAsync(peer_word, stage, sizeof(unsigned), cudaMemcpyDeviceToDevice, st);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync(flag)");
  return 0;
}

// The same for several peers at once: one single-warp kernel (p2p_kernels.cu) instead of two stream operations per
// peer.  B200FFT_FLAG_DMA=1 keeps the per-peer DMA form (A/B runs; also what machines without kernel access to
// peer memory would need).
int post_flags(b200fft_plan* pl, cudaStream_t st, unsigned* const* words, int n, unsigned value) {
  static const bool dma = [] { const char* e = std::getenv("B200FFT_FLAG_DMA"); return e && e[0] == '1'; }();
  if (dma) {
    for (int i = 0; i < n; ++i)
      if (int rc = post_flag(pl, st, words[i], value)) return rc;
    return 0;
  }
  const int rc = b200fft::launch_post_flags(words, n, value, st);
  if (rc) return fail(B200FFT_ERR_CUDA, "flag kernel launch failed (%d)", rc);
  return 0;
}

// Block `st` until every peer has handed this rank's blocks of the previous transform back (its last
// reader of received data has run): only then may this rank write into the peers' buffers again.
int wait_credits(b200fft_plan* pl, cudaStream_t st) {
  auto& pp = pl->p2p;
  if (pp.calls == 0) return 0;
  unsigned* fl = reinterpret_cast<unsigned*>(pp.flags);
  for (int q = 0; q < pl->d.nranks; ++q)
    if (q != pl->d.rank)
      if (CUresult r = g_cu.WaitValue32((CUstream)st, (CUdeviceptr)(fl + B200FFT_MAXP + q), pp.calls, CU_STREAM_WAIT_VALUE_GEQ))
        return cu_fail(r, "cuStreamWaitValue32(credit)");
  return 0;
}

// Copy-engine exchange of one step: push every peer's block, publish the sequence number, then
// (on the wait stream) wait for every peer's block to land here and record the step's event.
int run_exchange_p2p(b200fft_plan* pl, const Step& s, const void* in, void* out, size_t csz, cudaStream_t s1) {
  auto& pp = pl->p2p;
  if (!pp.connected) return fail(B200FFT_ERR_ARG, "P2P plan used before b200fft_plan_p2p_connect");
  const unsigned seq = ++pp.seq;
  if (s.first_exch && !s.fused)  // peers must have finished reading what the previous transform sent them
    if (int rc = wait_credits(pl, s1)) return rc;
  // fused transport: the FFT pass this step waited for has stored the blocks already
  // peers are members q of the step's (sub)communicator; buffers and flags are indexed by world rank
  const int me_w = pl->d.rank;
  // (one copy stream per peer -- the pushes of a step issued side by side -- was measured at 8 GPUs and lost:
  // 6.65 vs 5.47 ms, profiles/r02_multi_8; the pushes stay in order on the communication stream)
  for (int k = 1; k < s.npeers && !s.fused; ++k) {  // staggered peer order: no two ranks target the same GPU at once
    const int q = (s.me + k) % s.npeers;
    const int w = world_rank(pl->d, s.comm, me_w, q);
    char* dst = (char*)pp.peer_ws[w][s.rpeer[q].buf - BUF_W0] + (size_t)s.rpeer[q].off * csz;
    cudaError_t e = cudaMemcpyAsync(dst, resolve(pl, s.send[q], in, out, csz), (size_t)s.scnt[q] * csz, cudaMemcpyDeviceToDevice, s1);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync(peer)");
  }
  unsigned* words[B200FFT_MAXP];
  int nw = 0;
  for (int q = 0; q < s.npeers; ++q)
    if (q != s.me) words[nw++] = (unsigned*)pp.peer_flags[world_rank(pl->d, s.comm, me_w, q)] + me_w;
  return post_flags(pl, s1, words, nw, seq);
}

// wait-stream half of a P2P exchange step: own sends done + all peers' blocks arrived -> rec_ev
int finish_exchange_p2p(b200fft_plan* pl, const Step& s, cudaStream_t s1, int idx) {
  auto& pp = pl->p2p;
  while ((int)pp.send_ev.size() <= idx) {
    cudaEvent_t e;
    if (cudaError_t rc = cudaEventCreateWithFlags(&e, cudaEventDisableTiming)) return cuda_fail(rc, "cudaEventCreate");
    pp.send_ev.push_back(e);
  }
  cudaEventRecord(pp.send_ev[(size_t)idx], s1);
  cudaStreamWaitEvent(pp.wait_stream, pp.send_ev[(size_t)idx], 0);
  unsigned* fl = reinterpret_cast<unsigned*>(pp.flags);
  for (int q = 0; q < s.npeers; ++q)
    if (q != s.me)
      if (CUresult r = g_cu.WaitValue32((CUstream)pp.wait_stream, (CUdeviceptr)(fl + world_rank(pl->d, s.comm, pl->d.rank, q)), pp.seq,
                                        CU_STREAM_WAIT_VALUE_GEQ))
        return cu_fail(r, "cuStreamWaitValue32(arrived)");
  if (s.rec_ev >= 0) {
    cudaError_t e = cudaEventRecord(pl->sched_ev[(size_t)s.rec_ev], pp.wait_stream);
    if (e != cudaSuccess) return cuda_fail(e, "cudaEventRecord");
  }
  return 0;
}

b200fft_strided_desc_t strided_desc(const b200fft_plan* pl, const Step& s, const void* in, void* out, size_t csz) {
  b200fft_strided_desc_t d;
  std::memset(&d, 0, sizeof(d));
  d.precision = pl->d.precision;
  d.n = s.n;
  d.B = s.B;
  d.J = s.J;
  d.inverse = s.inverse;
  d.fold_mode = s.fold;
  d.scale = s.scale;
  fill_side(pl, s.in, d.in, in, out, csz);
  fill_side(pl, s.out, d.out, in, out, csz);
  d.mask = s.mask;
  d.cross_n = s.cross_n;
  d.cross_div = s.cross_div;
  return d;
}

b200fft_rows_desc_t rows_desc(const b200fft_plan* pl, const Step& s, const void* in, void* out, size_t csz) {
  b200fft_rows_desc_t d;
  std::memset(&d, 0, sizeof(d));
  d.precision = pl->d.precision;
  d.n = s.n;
  d.rows = s.rows;
  d.nk = s.nk;
  d.scale = s.scale;
  d.real_base = resolve(pl, s.real, in, out, csz / 2);
  d.rpitch = s.rpitch;
  fill_side(pl, s.cside, d.cside, in, out, csz);
  d.rm_period = s.rm_period;
  d.rm_block = s.rm_block;
  d.rm_planes = s.rm_planes;
  return d;
}

double step_bytes(const Step& s, size_t csz) {  // algorithmic bytes: operand read once, result written once
  if (s.type == ST_STRIDED && s.cross_n > 0) return 0.0;  // first launch of a four-step pass: the pass's algorithmic bytes are counted once, on the second
  if (s.type == ST_STRIDED) return (double)s.B * s.J * ((double)s.in.nphys + (double)s.out.nphys) * (double)csz;
  if (s.type == ST_R2C || s.type == ST_C2R) return (double)s.rows * ((double)s.n * (csz / 2) + (double)s.nk * csz);
  double b = 0;
  for (int q = 0; q < s.npeers; ++q)
    if (q != s.me) b += (double)s.scnt[q] * csz;
  return b;
}

int run_program(b200fft_plan* pl, int inverse, int dealias, const void* in, void* out, cudaStream_t st) {
  if (dealias < 0 || dealias > 2) return fail(B200FFT_ERR_ARG, "dealias must be None, '3/2-rule' or '2/3-rule'");
  if (!inverse && dealias == B200FFT_DEALIAS_2_3) dealias = B200FFT_DEALIAS_NONE;  // forward 2/3 == plain (slab.py:389)
  if (!in || !out) return fail(B200FFT_ERR_ARG, "null data pointer");
  if (int rc = ensure_program(pl, inverse, dealias)) return rc;
  const Program& pg = pl->prog[inverse][dealias];
  const size_t csz = pl->d.precision == B200FFT_DOUBLE ? 16 : 8;
  pl->last_kernels = 0;
  pl->last_exch = 0;
  pl->ev_marks.clear();
  pl->st_type.clear();
  pl->st_len.clear();
  pl->st_pass.clear();
  pl->st_bytes.clear();
  int evi = 0;
  if (pl->timing)
    while (pl->ev.size() < 2 * pg.steps.size()) {
      cudaEvent_t e;
      if (cudaError_t rc = cudaEventCreate(&e)) return cuda_fail(rc, "cudaEventCreate");
      pl->ev.push_back(e);
    }
  // scheduling resources of pipelined programs: a high-priority communication stream (its few
  // NCCL CTAs must win SM slots against the FFT grids) and one event per cross-stream edge
  if (pg.nevents > 0) {
    if (!pl->comm_stream) {
      int lo = 0, hi = 0;
      cudaDeviceGetStreamPriorityRange(&lo, &hi);
      cudaError_t e = cudaStreamCreateWithPriority(&pl->comm_stream, cudaStreamNonBlocking, hi);
      if (e != cudaSuccess) return cuda_fail(e, "cudaStreamCreateWithPriority");
    }
    while ((int)pl->sched_ev.size() < pg.nevents) {
      cudaEvent_t e;
      cudaError_t rc = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
      if (rc != cudaSuccess) return cuda_fail(rc, "cudaEventCreate");
      pl->sched_ev.push_back(e);
    }
  }
  if (pg.fork_ev >= 0) {  // the program's first step runs on the second stream: order it after the caller's work
    cudaError_t e = cudaEventRecord(pl->sched_ev[(size_t)pg.fork_ev], st);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(pl->comm_stream, pl->sched_ev[(size_t)pg.fork_ev], 0);
    if (e != cudaSuccess) return cuda_fail(e, "fork event");
  }
  bool use_p2p = false;
  for (const Step& s : pg.steps) use_p2p = use_p2p || (s.type == ST_EXCH);
  use_p2p = use_p2p && (pl->d.transport == B200FFT_TRANSPORT_P2P || pl->d.transport == B200FFT_TRANSPORT_STORE);
  if (use_p2p && !pl->p2p.connected) return fail(B200FFT_ERR_ARG, "P2P plan used before b200fft_plan_p2p_connect");
  int nexch = 0;
  cudaStream_t caller = st;
  for (size_t si = 0; si < pg.steps.size(); ++si) {
    const Step& s = pg.steps[si];
    st = (s.stream == 1 && pl->comm_stream) ? pl->comm_stream : caller;
    if (s.wait_ev >= 0) {
      cudaError_t e = cudaStreamWaitEvent(st, pl->sched_ev[(size_t)s.wait_ev], 0);
      if (e != cudaSuccess) return cuda_fail(e, "cudaStreamWaitEvent");
    }
    if (use_p2p && s.wait_credits)  // this pass stores into the peers' receive buffers
      if (int rc = wait_credits(pl, st)) return rc;
    if (pl->timing) cudaEventRecord(pl->ev[(size_t)evi], st);
    int rc = 0;
    if (s.type == ST_STRIDED) {
      rc = exec_strided(strided_desc(pl, s, in, out, csz), st);
      pl->last_kernels++;
    } else if (s.type == ST_R2C || s.type == ST_C2R) {
      rc = exec_rows(rows_desc(pl, s, in, out, csz), s.type == ST_R2C, st);
      pl->last_kernels++;
    } else if (use_p2p) {
      rc = run_exchange_p2p(pl, s, in, out, csz, st);
      pl->last_exch++;
    } else {
      rc = run_exchange(pl, s, in, out, csz, st);
      pl->last_exch++;
    }
    if (rc) return rc;
    pl->st_type.push_back((int)s.type);
    pl->st_len.push_back(s.type == ST_EXCH ? s.npeers : s.n);
    pl->st_pass.push_back(s.pass);
    pl->st_bytes.push_back(step_bytes(s, csz));
    if (pl->timing) {
      cudaEventRecord(pl->ev[(size_t)evi + 1], st);
      pl->ev_marks.emplace_back(evi, s.type == ST_EXCH ? 1 : 0);
      evi += 2;
    }
    if (s.type == ST_EXCH && use_p2p) {
      if (int rc2 = finish_exchange_p2p(pl, s, st, nexch++)) return rc2;
    } else if (s.rec_ev >= 0) {
      cudaError_t e = cudaEventRecord(pl->sched_ev[(size_t)s.rec_ev], st);
      if (e != cudaSuccess) return cuda_fail(e, "cudaEventRecord");
    }
    const bool last_reader = s.last_reader != 0;
    if (use_p2p && last_reader) {  // hand the receive buffers back to the peers
      unsigned* words[B200FFT_MAXP];
      int nw = 0;
      for (int q = 0; q < pl->d.nranks; ++q)
        if (q != pl->d.rank) words[nw++] = (unsigned*)pl->p2p.peer_flags[q] + B200FFT_MAXP + pl->d.rank;
      if (int rc2 = post_flags(pl, st, words, nw, pl->p2p.calls + 1)) return rc2;
    }
  }
  if (use_p2p) pl->p2p.calls++;
  return 0;
}

}  // namespace

// ================================================================================================
// extern "C"
// ================================================================================================
extern "C" {

int b200fft_version(void) { return 100; }

const char* b200fft_last_error(void) { return g_err.c_str(); }

int b200fft_supported_length(int n) { return plan_exists(n) ? 1 : 0; }

int b200fft_copy(void* dst, const void* src, size_t bytes, void* stream) {
  if (bytes == 0) return 0;
  if (!dst || !src) return fail(B200FFT_ERR_ARG, "null pointer in copy");
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync");
  return 0;
}

int b200fft_stream_sync(void* stream) {
  cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "cudaStreamSynchronize");
  return 0;
}

int b200fft_exec_strided(const b200fft_strided_desc_t* d, void* stream) {
  if (!d) return fail(B200FFT_ERR_ARG, "null descriptor");
  return exec_strided(*d, (cudaStream_t)stream);
}

int b200fft_exec_r2c(const b200fft_rows_desc_t* d, void* stream) {
  if (!d) return fail(B200FFT_ERR_ARG, "null descriptor");
  return exec_rows(*d, true, (cudaStream_t)stream);
}

int b200fft_exec_c2r(const b200fft_rows_desc_t* d, void* stream) {
  if (!d) return fail(B200FFT_ERR_ARG, "null descriptor");
  return exec_rows(*d, false, (cudaStream_t)stream);
}

int b200fft_comm_unique_id(void* id128) {
  if (!id128) return fail(B200FFT_ERR_ARG, "null id buffer");
  if (int rc = load_nccl()) return rc;
  ncclUniqueId id;
  static_assert(sizeof(ncclUniqueId) == 128, "unique id is 128 bytes");
  ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r != ncclSuccess) return nccl_fail(r, "ncclGetUniqueId");
  std::memcpy(id128, &id, 128);
  return 0;
}

int b200fft_comm_create(b200fft_comm_t* comm, int nranks, int rank, const void* id128) {
  if (!comm || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(B200FFT_ERR_ARG, "bad communicator arguments");
  if (int rc = load_nccl()) return rc;
  ncclUniqueId id;
  std::memcpy(&id, id128, 128);
  ncclComm_t c;
  ncclResult_t r = g_nccl.CommInitRank(&c, nranks, id, rank);
  if (r != ncclSuccess) return nccl_fail(r, "ncclCommInitRank");
  b200fft_comm* out = new b200fft_comm;
  out->comm = c;
  out->nranks = nranks;
  out->rank = rank;
  *comm = out;
  return 0;
}

int b200fft_comm_destroy(b200fft_comm_t comm) {
  if (!comm) return 0;
  if (g_nccl.CommDestroy) g_nccl.CommDestroy(comm->comm);
  delete comm;
  return 0;
}

int b200fft_plan_create(b200fft_plan_t* plan, const b200fft_plan_desc_t* d) {
  if (!plan || !d) return fail(B200FFT_ERR_ARG, "null argument");
  if (int rc = validate_desc(*d)) return rc;
  const bool peer_mapped = d->transport == B200FFT_TRANSPORT_P2P || d->transport == B200FFT_TRANSPORT_STORE;
  if (d->transport != B200FFT_TRANSPORT_NCCL && !peer_mapped)
    return fail(B200FFT_ERR_ARG, "unknown transport %d", d->transport);
  if (d->pipeline != B200FFT_PIPELINE_AUTO && d->pipeline != B200FFT_PIPELINE_X && d->pipeline != B200FFT_PIPELINE_KZ)
    return fail(B200FFT_ERR_ARG, "unknown pipeline %d", d->pipeline);
  if (peer_mapped && d->nranks > B200FFT_MAXP)
    return fail(B200FFT_ERR_RANKS, "peer-mapped transports address at most %d ranks", B200FFT_MAXP);
  if (d->nranks > 1 && d->transport == B200FFT_TRANSPORT_NCCL)
    if (int rc = load_nccl()) return rc;
  if (d->nranks > 1 && peer_mapped)
    if (int rc = load_cuda_driver()) return rc;
  b200fft_plan* pl = new b200fft_plan;
  pl->d = *d;
  if (pl->d.kind == B200FFT_LINE) pl->d.N[2] = 0;
  // the plain programs are validated eagerly so that unsupported sizes fail at construction
  Program probe;
  int rc = build_program(pl->d, 0, B200FFT_DEALIAS_NONE, probe);
  if (!rc) rc = check_lengths(probe);
  if (rc) {
    delete pl;
    return rc;
  }
  *plan = pl;
  return 0;
}

int b200fft_plan_destroy(b200fft_plan_t plan) {
  if (!plan) return 0;
  cudaDeviceSynchronize();
  if (plan->p2p.connected) {
    // Peer-mapped plans: a slower peer's last write into THIS rank's exported memory is the credit word of the
    // last transform (posted by its last reader, after every push and arrival flag).  Freeing the flags or the
    // work buffers before it landed would let the peer's DMA hit freed memory, so wait (bounded: a peer that
    // died never posts) until every peer's credit shows the last call, then drop the imported mappings BEFORE
    // the exported allocations go.
    if (plan->p2p.calls > 0 && plan->p2p.flags) {
      unsigned host[2 * B200FFT_MAXP];
      for (int spin = 0; spin < 20000; ++spin) {
        if (cudaMemcpy(host, plan->p2p.flags, sizeof(host), cudaMemcpyDeviceToHost) != cudaSuccess) break;
        bool all = true;
        for (int q = 0; q < plan->d.nranks; ++q)
          if (q != plan->d.rank && (int)(host[B200FFT_MAXP + q] - plan->p2p.calls) < 0) all = false;
        if (all) break;
        std::this_thread::sleep_for(std::chrono::microseconds(100));
      }
    }
    for (int q = 0; q < plan->d.nranks; ++q) {
      if (q == plan->d.rank) continue;
      for (int w = 0; w < NPEERBUF; ++w)
        if (plan->p2p.peer_ws[q][w]) cudaIpcCloseMemHandle(plan->p2p.peer_ws[q][w]);
      if (plan->p2p.peer_flags[q]) cudaIpcCloseMemHandle(plan->p2p.peer_flags[q]);
    }
  }
  for (int w = 0; w < NWORK; ++w)
    if (plan->ws[w]) cudaFree(plan->ws[w]);
  for (cudaEvent_t e : plan->ev) cudaEventDestroy(e);
  for (cudaEvent_t e : plan->sched_ev) cudaEventDestroy(e);
  for (cudaEvent_t e : plan->p2p.send_ev) cudaEventDestroy(e);
  if (plan->p2p.flags) cudaFree(plan->p2p.flags);
  if (plan->p2p.wait_stream) cudaStreamDestroy(plan->p2p.wait_stream);
  if (plan->comm_stream) cudaStreamDestroy(plan->comm_stream);
  delete plan;
  return 0;
}

size_t b200fft_plan_workspace_bytes(b200fft_plan_t plan) {
  if (!plan) return 0;
  size_t total = 0;
  for (int w = 0; w < NWORK; ++w) total += plan->wbytes[w];
  return total;
}

int b200fft_exec_forward(b200fft_plan_t plan, const void* u, void* fu, int dealias, void* stream) {
  if (!plan) return fail(B200FFT_ERR_ARG, "null plan");
  return run_program(plan, 0, dealias, u, fu, (cudaStream_t)stream);
}

int b200fft_exec_inverse(b200fft_plan_t plan, const void* fu, void* u, int dealias, void* stream) {
  if (!plan) return fail(B200FFT_ERR_ARG, "null plan");
  return run_program(plan, 1, dealias, fu, u, (cudaStream_t)stream);
}

int b200fft_plan_p2p_handles(b200fft_plan_t plan, void* handles256) {
  if (!plan || !handles256) return fail(B200FFT_ERR_ARG, "null argument");
  if ((plan->d.transport != B200FFT_TRANSPORT_P2P && plan->d.transport != B200FFT_TRANSPORT_STORE) || plan->d.nranks < 2)
    return fail(B200FFT_ERR_ARG, "not a multi-rank P2P plan");
  // size the work buffers for every program now: their addresses are what the peers map
  size_t need[NWORK];
  for (int w = 0; w < NWORK; ++w) need[w] = 256;
  const size_t csz = plan->d.precision == B200FFT_DOUBLE ? 16 : 8;
  for (int inv = 0; inv < 2; ++inv)
    for (int de = 0; de < 3; ++de) {
      if (de == B200FFT_DEALIAS_3_2 && std::fabs(plan->d.padsize - 1.5) > 1e-12) continue;
      Program pg;
      if (build_program(plan->d, inv, de, pg) || check_lengths(pg)) continue;
      for (int w = 0; w < NWORK; ++w) need[w] = std::max(need[w], (size_t)pg.need[BUF_W0 + w] * csz);
    }
  for (int w = 0; w < NWORK; ++w) {
    if (plan->ws[w]) cudaFree(plan->ws[w]);
    cudaError_t e = cudaMalloc(&plan->ws[w], need[w]);
    if (e != cudaSuccess) return fail(B200FFT_ERR_NOMEM, "cudaMalloc(%zu bytes of work space): %s", need[w], cudaGetErrorString(e));
    plan->wbytes[w] = need[w];
  }
  cudaError_t e = cudaMalloc(&plan->p2p.flags, (2 * B200FFT_MAXP + P2P_STAGE) * sizeof(unsigned));
  if (e == cudaSuccess) e = cudaMemset(plan->p2p.flags, 0, (2 * B200FFT_MAXP + P2P_STAGE) * sizeof(unsigned));
  if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(flags)");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handles are 64 bytes");
  cudaIpcMemHandle_t* h = reinterpret_cast<cudaIpcMemHandle_t*>(handles256);
  for (int w = 0; w < NPEERBUF; ++w)  // (W3 is a local send buffer: peers never address it)
    if ((e = cudaIpcGetMemHandle(&h[w], plan->ws[w])) != cudaSuccess) return cuda_fail(e, "cudaIpcGetMemHandle");
  if ((e = cudaIpcGetMemHandle(&h[3], plan->p2p.flags)) != cudaSuccess) return cuda_fail(e, "cudaIpcGetMemHandle(flags)");
  return 0;
}

int b200fft_plan_p2p_connect(b200fft_plan_t plan, const void* all_handles) {
  if (!plan || !all_handles) return fail(B200FFT_ERR_ARG, "null argument");
  if (!plan->p2p.flags) return fail(B200FFT_ERR_ARG, "call b200fft_plan_p2p_handles first");
  const cudaIpcMemHandle_t* h = reinterpret_cast<const cudaIpcMemHandle_t*>(all_handles);
  for (int q = 0; q < plan->d.nranks; ++q) {
    if (q == plan->d.rank) continue;
    for (int w = 0; w < NPEERBUF; ++w) {
      cudaError_t e = cudaIpcOpenMemHandle(&plan->p2p.peer_ws[q][w], h[4 * q + w], cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) return cuda_fail(e, "cudaIpcOpenMemHandle (is NVLink / P2P available between the GPUs?)");
    }
    cudaError_t e = cudaIpcOpenMemHandle(&plan->p2p.peer_flags[q], h[4 * q + 3], cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return cuda_fail(e, "cudaIpcOpenMemHandle(flags)");
  }
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  cudaError_t e = cudaStreamCreateWithPriority(&plan->p2p.wait_stream, cudaStreamNonBlocking, hi);
  if (e != cudaSuccess) return cuda_fail(e, "cudaStreamCreateWithPriority");
  plan->p2p.connected = true;
  return 0;
}

int b200fft_plan_last_launches(b200fft_plan_t plan, int* kernels, int* exchanges) {
  if (!plan) return fail(B200FFT_ERR_ARG, "null plan");
  if (kernels) *kernels = plan->last_kernels;
  if (exchanges) *exchanges = plan->last_exch;
  return 0;
}

int b200fft_plan_set_timing(b200fft_plan_t plan, int on) {
  if (!plan) return fail(B200FFT_ERR_ARG, "null plan");
  plan->timing = on ? 1 : 0;
  return 0;
}

int b200fft_plan_last_phase_ms(b200fft_plan_t plan, float* fft_ms, float* exchange_ms) {
  if (!plan) return fail(B200FFT_ERR_ARG, "null plan");
  float f = -1.f, x = -1.f;
  if (plan->timing && !plan->ev_marks.empty()) {
    f = 0.f;
    x = 0.f;
    for (auto& m : plan->ev_marks) {
      cudaEventSynchronize(plan->ev[m.first + 1]);
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, plan->ev[m.first], plan->ev[m.first + 1]) == cudaSuccess) (m.second ? x : f) += ms;
    }
  }
  if (fft_ms) *fft_ms = f;
  if (exchange_ms) *exchange_ms = x;
  return 0;
}

int b200fft_plan_last_steps(b200fft_plan_t plan, int max, int* n, int* type, float* ms, double* bytes, int* len, int* pass) {
  if (!plan || !n) return fail(B200FFT_ERR_ARG, "null argument");
  const int cnt = (int)plan->st_type.size();
  *n = cnt;
  for (int i = 0; i < cnt && i < max; ++i) {
    if (type) type[i] = plan->st_type[i];
    if (bytes) bytes[i] = plan->st_bytes[i];
    if (len) len[i] = plan->st_len[i];
    if (pass) pass[i] = plan->st_pass[i];
    if (ms) {
      ms[i] = -1.f;
      if (plan->timing && i < (int)plan->ev_marks.size()) {
        const int e0 = plan->ev_marks[i].first;
        cudaEventSynchronize(plan->ev[e0 + 1]);
        float t = 0.f;
        if (cudaEventElapsedTime(&t, plan->ev[e0], plan->ev[e0 + 1]) == cudaSuccess) ms[i] = t;
      }
    }
  }
  return 0;
}

}  // extern "C"
