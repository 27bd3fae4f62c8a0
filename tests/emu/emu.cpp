// CPU emulator of the CUDA FFT kernels -- TEST INFRASTRUCTURE ONLY (never loaded by the product).
//
// Compiles mpifft4py_b200/csrc/fft_kernels.cuh with g++ (no CUDA) and executes the very same
// phase bodies the __global__ kernels run: all threads of a block are stepped through phase s
// before any thread starts phase s+1, which is exactly what __syncthreads() enforces on the GPU.
// This checks radix plans, digit-reversal, swizzles and every fused index map (pad, truncate +
// fold, peer chunks, masks) in the `-m "not gpu"` suite; the `-m gpu` suite checks the real thing.
#include <cstdio>
#include <map>
#include <vector>

#include "../../mpifft4py_b200/csrc/desc_convert.h"
#include "../../mpifft4py_b200/csrc/fft_plans.h"
#include "../../mpifft4py_b200/csrc/plan_program.h"

using namespace b200fft;

template <class K, int s>
static void emu_phases(const typename K::Params& p, std::vector<unsigned char>& sm, int bx, int by) {
  for (int tid = 0; tid < K::NT; ++tid) K::template phase<s>(p, sm.data(), tid, bx, by);
  if constexpr (s + 1 < K::NPHASE) emu_phases<K, s + 1>(p, sm, bx, by);
}

template <class K>
static int emulate(const typename K::Params& p) {
  const unsigned long long nblk = K::blocks(p);
  std::vector<unsigned char> sm((size_t)(K::SMEM > 0 ? K::SMEM : 16) + 256);
  for (unsigned long long b = 0; b < nblk; ++b) {
    int bx, by;
    K::decode(p, (unsigned)b, bx, by);
    // poison shared memory so that reads of unwritten slots show up
    for (auto& c : sm) c = 0x7f;
    emu_phases<K, 0>(p, sm, bx, by);
  }
  return 0;
}

// 2-CTA cluster kernels: both CTAs of a cluster are stepped through phase s (each seeing the other's
// shared memory as `peer`) before either starts phase s+1 -- at least as strict as the cluster barriers
// around the cross stage and the CTA barriers elsewhere.
template <class K, int s>
static void emu_cluster_phases(const typename K::Params& p, std::vector<unsigned char>* sm, int bx, int by) {
  for (int rank = 0; rank < 2; ++rank)
    for (int tid = 0; tid < K::NT; ++tid) K::template phase<s>(p, sm[rank].data(), sm[rank ^ 1].data(), tid, rank, bx, by);
  if constexpr (s + 1 < K::NPHASE) emu_cluster_phases<K, s + 1>(p, sm, bx, by);
}

template <class K>
static int emulate_cluster(const typename K::Params& p) {
  const unsigned long long nblk = K::blocks(p);
  std::vector<unsigned char> sm[2];
  for (auto& v : sm) v.resize((size_t)K::SMEM + 256);
  for (unsigned long long b = 0; b < nblk; b += 2) {
    int bx, by, bx1, by1;
    K::decode(p, (unsigned)b, bx, by);
    K::decode(p, (unsigned)b + 1, bx1, by1);
    if (bx != bx1 || by != by1) return 95;  // both CTAs of a cluster work on the same tile
    for (auto& v : sm)
      for (auto& c : v) c = 0x7f;
    emu_cluster_phases<K, 0>(p, sm, bx, by);
  }
  return 0;
}

// Fused pair of passes (fused_pair_kernel): the queue is walked in index order by ONE worker, which is a
// legal schedule of the persistent kernel; on the way it checks what the device relies on -- the decode
// visits every block of both grids exactly once, and every block of pass B finds the pass-A counters of
// the groups it reads complete (i.e. all blocks it would wait for have a smaller queue index).
template <class KA, class KB>
static int emulate_fused_pair(const typename KA::Params& pa, const typename KB::Params& pb, long long planes, long long ppg) {
  if (planes < 1 || ppg < 1) return -3;
  FuseCtl c;
  c.ctr = nullptr;
  c.done = nullptr;
  c.G = (unsigned)((planes + ppg - 1) / ppg);
  c.a = KA::fuse_side(pa, planes, ppg);
  c.b = KB::fuse_side(pb, planes, ppg);
  if (c.G + 1 > 4097 || c.a.n == 0 || c.b.n == 0 || c.a.upg < c.a.upb || c.b.upg < c.b.upb) return -3;
  std::vector<unsigned> done(c.G, 0);
  std::vector<char> seenA(c.a.n, 0), seenB(c.b.n, 0);
  const size_t smem = (size_t)(KA::SMEM1 > KB::SMEM1 ? KA::SMEM1 : KB::SMEM1) + 256;
  std::vector<unsigned char> sm(smem);
  for (unsigned i = 0; i < c.a.n + c.b.n; ++i) {
    bool isB;
    unsigned blk, g1, g2, u1, u2;
    fuse_decode(c, i, isB, blk);
    for (auto& x : sm) x = 0x7f;
    int bx, by;
    if (!isB) {
      if (blk >= c.a.n || seenA[blk]++) return 97;
      KA::decode(pa, blk, bx, by);
      emu_phases<KA, 0>(pa, sm, bx, by);
      c.a.groups(blk, g1, g2, u1, u2);
      if (g2 >= c.G || g2 > g1 + 1) return 94;
      done[g1] += u1;
      if (g2 != g1) done[g2] += u2;
    } else {
      if (blk >= c.b.n || seenB[blk]++) return 98;
      c.b.groups(blk, g1, g2, u1, u2);
      if (g2 >= c.G) return 95;
      for (unsigned g = g1; g <= g2; ++g)
        if (done[g] != c.a.need(g)) return 96;  // the device would wait for a block that is queued later: deadlock
      KB::decode(pb, blk, bx, by);
      emu_phases<KB, 0>(pb, sm, bx, by);
    }
  }
  for (char x : seenA) if (!x) return 99;
  for (char x : seenB) if (!x) return 99;
  for (unsigned g = 0; g < c.G; ++g) if (done[g] != c.a.need(g)) return 93;
  return 0;
}

template <class real>
static int fused_zy(const b200fft_rows_desc_t& r, const b200fft_strided_desc_t& c, int inverse_order, int ppg);

static int g_emu_variant = 0;
static int g_emu_fused_runs = 0;  // fused launches executed by emu_plan_run (tests check the path was taken)

template <class real>
static const cx<real>* table(int len) {
  static std::map<int, std::vector<cx<real>>> cache;
  auto it = cache.find(len);
  if (it == cache.end()) it = cache.emplace(len, make_twiddles<real>(len)).first;
  return it->second.data();
}

template <class real>
static int strided(const b200fft_strided_desc_t& d) {
  auto p = convert_strided<real>(d, table<real>(d.n), 1);
  if (contiguous_rows(d)) {
    switch (d.n) {
#define X(n, ...) \
  case n:         \
    return emulate<RowC2CK<real, Plan<__VA_ARGS__>>>(p);
      B200FFT_PLANS(X)
#undef X
      default:
        return -1;
    }
  }
  if (g_emu_variant == 21) {
#define X(nn, ...) \
  if (d.n == nn) return emulate_cluster<ClusterStridedK<real, Plan<__VA_ARGS__>>>(p);
    B200FFT_CLUSTER_PLANS(X)
#undef X
  }
  if (g_emu_variant == 23) {  // 64-byte rows
#define X(nn, ...) \
  if (d.n == nn) return emulate_cluster<ClusterStridedK<real, Plan<__VA_ARGS__>, 64>>(p);
    B200FFT_CLUSTER_PLANS(X)
#undef X
  }
  if (g_emu_variant == 35) {  // first stage fed straight from memory (plans with two or more stages)
    switch (d.n) {
#define X(n, ...)                                                                            \
  case n:                                                                                    \
    if constexpr (Plan<__VA_ARGS__>::S >= 2 && StridedCfg<real, Plan<__VA_ARGS__>>::TAB)     \
      return emulate<StridedDK<real, Plan<__VA_ARGS__>>>(p);                                 \
    break;
      B200FFT_PLANS(X)
#undef X
      default:
        break;
    }
  }
  if (d.in.jc > 0 || d.out.jc > 0) {  // blocked column layouts: the JS form of the kernel (every length here)
    switch (d.n) {
#define X(n, ...) \
  case n:         \
    return emulate<StridedK<real, Plan<__VA_ARGS__>, 0, 0, false, true>>(p);
      B200FFT_PLANS(X)
#undef X
      default:
        return -1;
    }
  }
  switch (d.n) {
#define X(n, ...) \
  case n:         \
    return emulate<StridedK<real, Plan<__VA_ARGS__>>>(p);
    B200FFT_PLANS(X)
#undef X
    default:
      return -1;
  }
}

template <class real, bool FWD>
static int rows(const b200fft_rows_desc_t& d) {
  auto p = convert_rows<real>(d, table<real>(d.n), 1, FWD);
  switch (d.n / 2) {
#define X(n, ...)                                                     \
  case n:                                                             \
    if (FWD && g_emu_variant == 33) {                                 \
      if constexpr (Plan<__VA_ARGS__>::S >= 2) return emulate<R2CPK<real, Plan<__VA_ARGS__>>>(p); \
    }                                                                 \
    if (FWD) return emulate<R2CK<real, Plan<__VA_ARGS__>>>(p);        \
    else if (g_emu_variant == 31) return emulate<C2RDK<real, Plan<__VA_ARGS__>>>(p); \
    else if (g_emu_variant == 34) {                                    \
      using PP = Plan<__VA_ARGS__>;                                    \
      if constexpr (PP::S >= 2 && PP::template R<0> <= 8 && (PP::template M<0> % 2) == 0) return emulate<C2RPK<real, PP>>(p); \
      else return emulate<C2RK<real, PP>>(p);                          \
    }                                                                  \
    else return emulate<C2RK<real, Plan<__VA_ARGS__>>>(p);
    B200FFT_ROW_PLANS(X)
#undef X
    default:
      return -1;
  }
}

template <class real>
static int fused_zy(const b200fft_rows_desc_t& r, const b200fft_strided_desc_t& c, int inverse_order, int ppg) {
  if (contiguous_rows(c) || c.B < 1 || r.rows % c.B || c.in.jc > 0 || c.out.jc > 0) return -1;
  auto pr = convert_rows<real>(r, table<real>(r.n), 1, !inverse_order);
  auto pc = convert_strided<real>(c, table<real>(c.n), 1);
#define X(h, ny, PR, PC)                                                                                  \
  if (r.n / 2 == h && c.n == ny)                                                                          \
    return inverse_order ? emulate_fused_pair<StridedK<real, PC>, C2RK<real, PR>>(pc, pr, c.B, ppg)       \
                         : emulate_fused_pair<R2CK<real, PR>, StridedK<real, PC>>(pr, pc, c.B, ppg);
  B200FFT_FUSED_PAIRS(X)
#undef X
  return -1;
}

extern "C" {
// -1: no fused kernel for the size pair, -3: too many / too small groups (callers fall back)
int emu_exec_fused_zy(const b200fft_rows_desc_t* r, const b200fft_strided_desc_t* c, int inverse_order, int ppg) {
  if (const char* e = check_rows(*r)) { std::fprintf(stderr, "emu: %s\n", e); return 1; }
  if (const char* e = check_strided(*c)) { std::fprintf(stderr, "emu: %s\n", e); return 1; }
  return r->precision == B200FFT_DOUBLE ? fused_zy<double>(*r, *c, inverse_order, ppg) : fused_zy<float>(*r, *c, inverse_order, ppg);
}
int emu_fused_runs() { return g_emu_fused_runs; }
// Queue of the fused pair kernel without running any FFT: rows * planes rows in blocks of `rpc` rows next to
// `tiles` column tiles per plane, groups of `ppg` planes, either order.  Checks that the queue visits every
// block of both passes once and that no block of the second pass precedes a block it waits for.
int emu_check_fuse_queue(long long rows_per_plane, int rpc, long long planes, long long ppg, int tiles, int rows_first) {
  FuseCtl c;
  c.ctr = c.done = nullptr;
  c.G = (unsigned)((planes + ppg - 1) / ppg);
  const long long rows = rows_per_plane * planes;
  const FuseSide r{(unsigned)((rows + rpc - 1) / rpc), (unsigned)rpc, (unsigned)(ppg * rows_per_plane), (unsigned)rows};
  const FuseSide t{(unsigned)(planes * tiles), 1u, (unsigned)(ppg * tiles), (unsigned)(planes * tiles)};
  c.a = rows_first ? r : t;
  c.b = rows_first ? t : r;
  if (c.a.upg < c.a.upb || c.b.upg < c.b.upb) return -3;
  std::vector<unsigned> done(c.G, 0);
  std::vector<char> seenA(c.a.n, 0), seenB(c.b.n, 0);
  for (unsigned i = 0; i < c.a.n + c.b.n; ++i) {
    bool isB;
    unsigned blk, g1, g2, u1, u2;
    fuse_decode(c, i, isB, blk);
    if (!isB) {
      if (blk >= c.a.n || seenA[blk]++) return 97;
      c.a.groups(blk, g1, g2, u1, u2);
      if (g2 >= c.G || g2 > g1 + 1) return 94;
      done[g1] += u1;
      if (g2 != g1) done[g2] += u2;
    } else {
      if (blk >= c.b.n || seenB[blk]++) return 98;
      c.b.groups(blk, g1, g2, u1, u2);
      if (g2 >= c.G) return 95;
      for (unsigned g = g1; g <= g2; ++g)
        if (done[g] != c.a.need(g)) return 96;
    }
  }
  for (char x : seenA) if (!x) return 99;
  for (char x : seenB) if (!x) return 99;
  return 0;
}
int emu_set_variant(int v) {  // 21: lengths with a cluster plan run the 2-CTA cluster kernel
  const int old = g_emu_variant;
  g_emu_variant = v;
  return old;
}
int emu_exec_strided(const b200fft_strided_desc_t* d) {
  if (const char* e = check_strided(*d)) { std::fprintf(stderr, "emu: %s\n", e); return 1; }
  return d->precision == B200FFT_DOUBLE ? strided<double>(*d) : strided<float>(*d);
}
int emu_exec_r2c(const b200fft_rows_desc_t* d) {
  if (const char* e = check_rows(*d)) { std::fprintf(stderr, "emu: %s\n", e); return 1; }
  return d->precision == B200FFT_DOUBLE ? rows<double, true>(*d) : rows<float, true>(*d);
}
int emu_exec_c2r(const b200fft_rows_desc_t* d) {
  if (const char* e = check_rows(*d)) { std::fprintf(stderr, "emu: %s\n", e); return 1; }
  return d->precision == B200FFT_DOUBLE ? rows<double, false>(*d) : rows<float, false>(*d);
}
// Run one distributed transform for ALL ranks in lockstep: the same plan programs as
// libb200fft.so (plan_program.h), kernels emulated on the CPU, exchanges done by memcpy.
int emu_plan_run(const b200fft_plan_desc_t* d0, int inverse, int dealias, void** ins, void** outs) {
  const int P = d0->nranks;
  if (!inverse && dealias == B200FFT_DEALIAS_2_3) dealias = B200FFT_DEALIAS_NONE;
  const size_t csz = d0->precision == B200FFT_DOUBLE ? 16 : 8;
  std::vector<Program> pg((size_t)P);
  std::vector<std::vector<unsigned char>> ws((size_t)P * NWORK);
  for (int r = 0; r < P; ++r) {
    b200fft_plan_desc_t d = *d0;
    d.rank = r;
    if (int rc = build_program(d, inverse, dealias, pg[r])) {
      std::fprintf(stderr, "emu plan: %s\n", plan_err().c_str());
      return rc;
    }
    for (int w = 0; w < NWORK; ++w) ws[(size_t)r * NWORK + w].assign((size_t)pg[r].need[BUF_W0 + w] * csz + 64, 0xff);  // NaN poison
  }
  auto resolve = [&](int r, const Ref& ref, size_t esz) -> void* {
    char* base;
    if (ref.buf == BUF_IN) base = (char*)ins[r];
    else if (ref.buf == BUF_OUT) base = (char*)outs[r];
    else base = (char*)ws[(size_t)(ref.peer >= 0 ? ref.peer : r) * NWORK + (ref.buf - BUF_W0)].data();  // peer: fused transport
    return base + (size_t)ref.off * esz;
  };
  auto side = [&](int r, const SideT& s) {
    b200fft_side_t o;
    std::memset(&o, 0, sizeof(o));
    for (int q = 0; q < s.nchunk; ++q) { o.base[q] = resolve(r, s.base[q], csz); o.sb[q] = s.sb[q]; o.si[q] = s.si[q]; }
    o.chunk = s.chunk; o.nchunk = s.nchunk; o.nphys = s.nphys; o.jc = s.jc; o.sj = s.sj;
    return o;
  };
  const size_t nsteps = pg[0].steps.size();
  for (int r = 1; r < P; ++r) if (pg[r].steps.size() != nsteps) return 90;
  auto strided_of = [&](int r, const Step& s) {
    b200fft_strided_desc_t d;
    std::memset(&d, 0, sizeof(d));
    d.precision = d0->precision; d.n = s.n; d.B = s.B; d.J = s.J; d.inverse = s.inverse;
    d.fold_mode = s.fold; d.scale = s.scale; d.in = side(r, s.in); d.out = side(r, s.out); d.mask = s.mask;
    return d;
  };
  auto rows_of = [&](int r, const Step& s) {
    b200fft_rows_desc_t d;
    std::memset(&d, 0, sizeof(d));
    d.precision = d0->precision; d.n = s.n; d.rows = s.rows; d.nk = s.nk; d.scale = s.scale;
    d.real_base = resolve(r, s.real, csz / 2); d.rpitch = s.rpitch; d.cside = side(r, s.cside);
    return d;
  };
  for (size_t si = 0; si < nsteps; ++si) {
    if (pg[0].steps[si].fuse_planes > 0 && si + 1 < nsteps) {  // as b200fft.cu: fused launch, else two passes
      bool all = true;
      for (int r = 0; r < P && all; ++r) {  // (the same sizes on every rank: all fuse or none does)
        const Step& s = pg[r].steps[si];
        const Step& t = pg[r].steps[si + 1];
        const Step& rs = (s.type == ST_STRIDED) ? t : s;
        const Step& cs = (s.type == ST_STRIDED) ? s : t;
        if (cs.type != ST_STRIDED || (rs.type != ST_R2C && rs.type != ST_C2R)) { all = false; break; }
        auto rd = rows_of(r, rs);
        auto cd = strided_of(r, cs);
        const int rc = emu_exec_fused_zy(&rd, &cd, s.type == ST_STRIDED, s.fuse_planes);
        if (rc == -1 || rc == -3) {
          if (r != 0) return 92;
          all = false;
        } else if (rc) {
          return rc;
        } else {
          ++g_emu_fused_runs;
        }
      }
      if (all) { ++si; continue; }
    }
    for (int r = 0; r < P; ++r) {
      const Step& s = pg[r].steps[si];
      int rc = 0;
      if (s.type == ST_STRIDED) {
        b200fft_strided_desc_t d;
        std::memset(&d, 0, sizeof(d));
        d.precision = d0->precision; d.n = s.n; d.B = s.B; d.J = s.J; d.inverse = s.inverse;
        d.fold_mode = s.fold; d.scale = s.scale; d.in = side(r, s.in); d.out = side(r, s.out); d.mask = s.mask;
        rc = emu_exec_strided(&d);
      } else if (s.type == ST_R2C || s.type == ST_C2R) {
        b200fft_rows_desc_t d;
        std::memset(&d, 0, sizeof(d));
        d.precision = d0->precision; d.n = s.n; d.rows = s.rows; d.nk = s.nk; d.scale = s.scale;
        d.real_base = resolve(r, s.real, csz / 2); d.rpitch = s.rpitch; d.cside = side(r, s.cside);
        rc = s.type == ST_R2C ? emu_exec_r2c(&d) : emu_exec_c2r(&d);
      } else {
        for (int q = 0; q < s.npeers; ++q) {
          if (q == s.me) continue;
          const int w = world_rank(*d0, s.comm, r, q);  // pencil.py:192-195
          const Step& t = pg[w].steps[si];
          if (t.type != ST_EXCH || t.rcnt[s.me] != s.scnt[q]) return 91;
          if (s.fused) continue;  // the producing pass stored straight into the peer's buffer
          std::memcpy(resolve(w, t.recv[s.me], csz), resolve(r, s.send[q], csz), (size_t)s.scnt[q] * csz);
        }
      }
      if (rc) return rc;
    }
  }
  return 0;
}

// Copy-engine transport invariants of every program of a plan: rank r pushes its block for peer q to
// rpeer[q]; that must be exactly where q's own program expects the block from r (recv[r]), with equal
// counts, inside q's buffer; exactly one first_exch and one last_reader per program with exchanges;
// all ranks agree on the number of exchange steps (the sequence numbers advance in lockstep).
int emu_check_p2p(const b200fft_plan_desc_t* d0, int inverse, int dealias, int* nexch_out) {
  const int P = d0->nranks;
  std::vector<Program> pg((size_t)P);
  for (int r = 0; r < P; ++r) {
    b200fft_plan_desc_t d = *d0;
    d.rank = r;
    if (int rc = build_program(d, inverse, dealias, pg[r])) return rc;
  }
  const size_t nsteps = pg[0].steps.size();
  int nexch = 0;
  for (int r = 0; r < P; ++r) {
    if (pg[r].steps.size() != nsteps) return 100;
    int first = 0, last = 0, ex = 0, credits = 0, peer_stores = 0;
    for (size_t si = 0; si < nsteps; ++si) {
      const Step& s = pg[r].steps[si];
      last += s.last_reader;
      if (s.type != ST_EXCH) {
        // fused transport: a pass may store into peer q's buffer only where q's exchange step expects
        // this rank's block, and only after the credits were awaited
        credits += s.wait_credits;
        const SideT& o = s.type == ST_STRIDED ? s.out : s.cside;
        const SideT& in = s.in;
        for (int q = 0; q < in.nchunk && s.type == ST_STRIDED; ++q)
          if (in.base[q].peer >= 0) return 110;  // no remote loads
        for (int q = 0; q < o.nchunk; ++q) {
          if (o.base[q].peer < 0) continue;
          if (d0->transport != B200FFT_TRANSPORT_STORE || s.type == ST_C2R) return 111;
          if (!credits) return 112;
          ++peer_stores;
          // the next exchange step of this program is the one that announces these stores
          size_t sx = si + 1;
          while (sx < nsteps && pg[r].steps[sx].type != ST_EXCH) ++sx;
          if (sx == nsteps) return 113;
          const Step& x = pg[r].steps[sx];
          const int w = world_rank(*d0, x.comm, r, q);  // chunk q belongs to member q of that exchange's communicator
          if (o.base[q].peer != w || w == r || o.nchunk != x.npeers) return 111;
          const Step& t = pg[w].steps[sx];
          if (!x.fused || t.type != ST_EXCH || !t.fused || t.comm != x.comm) return 114;
          // it must lie inside the block peer w receives from this rank (member x.me of the communicator);
          // all of it, when the pass is not one of several L2 groups of the chunk
          const long long rows_q = (q == o.nchunk - 1) ? o.nphys - (long long)q * o.chunk : o.chunk;
          const long long nb = s.type == ST_STRIDED ? s.B : s.rows, nj = s.type == ST_STRIDED ? s.J : 1;
          const long long ext = (nb - 1) * o.sb[q] + (rows_q - 1) * o.si[q] + nj;
          if (o.base[q].buf != t.recv[x.me].buf || o.base[q].off < t.recv[x.me].off ||
              o.base[q].off + ext > t.recv[x.me].off + t.rcnt[x.me])
            return 115;
          if (d0->l2_planes <= 0 && (o.base[q].off != t.recv[x.me].off || ext != t.rcnt[x.me])) return 120;
          if (o.base[q].buf < BUF_W0 || o.base[q].buf > BUF_W2 || t.recv[x.me].off + t.rcnt[x.me] > pg[w].need[t.recv[x.me].buf]) return 116;
        }
        continue;
      }
      if ((d0->transport == B200FFT_TRANSPORT_STORE) != (s.fused != 0)) return 117;
      ++ex;
      first += s.first_exch;
      if (s.rec_ev < 0) return 101;  // arrival is awaited on another stream: the consumer needs the event
      for (int q = 0; q < s.npeers; ++q) {
        if (q == s.me) continue;
        const int w = world_rank(*d0, s.comm, r, q);
        const Step& t = pg[w].steps[si];
        if (t.type != ST_EXCH || t.comm != s.comm || world_rank(*d0, t.comm, w, s.me) != r) return 102;
        if (s.rpeer[q].buf != t.recv[s.me].buf || s.rpeer[q].off != t.recv[s.me].off) return 103;
        if (s.scnt[q] != t.rcnt[s.me]) return 104;
        if (s.rpeer[q].buf < BUF_W0 || s.rpeer[q].buf > BUF_W2) return 105;  // peers may only write the mapped plan buffers
        if (s.rpeer[q].off + s.scnt[q] > pg[w].need[s.rpeer[q].buf]) return 106;
      }
    }
    if (ex > 0 && (first != 1 || last != 1)) return 107;
    if (d0->transport == B200FFT_TRANSPORT_STORE && ex > 0 && (credits != 1 || peer_stores == 0)) return 118;
    if (d0->transport != B200FFT_TRANSPORT_STORE && (credits || peer_stores)) return 119;
    if (r == 0) nexch = ex;
    else if (ex != nexch) return 108;
  }
  if (nexch_out) *nexch_out = nexch;
  return 0;
}

}  // extern "C"

// Schedule check of every rank's program: the device runs steps on two streams ordered by events
// (Step::stream / wait_ev / rec_ev), the emulator in program order -- so a missing dependency would
// pass every emulator run and race on the GPU.  This derives the happens-before relation the device
// actually enforces (stream order + event edges; a wait on an event that is recorded LATER in program
// order is a no-op in CUDA and counts as an error) and requires it between any two steps of a rank
// that touch overlapping parts of the same local buffer with at least one write.  Regions are
// bounding intervals (conservative).  Peer stores of the fused transport are ordered by the flag /
// credit protocol, not by this rank's streams, and are left out.
namespace {
struct Region { int buf; long long lo, hi; bool write; };

void side_regions(const SideT& s, long long nb, long long nj, bool write, long long unit, std::vector<Region>& out) {
  for (int q = 0; q < s.nchunk; ++q) {
    if (s.base[q].peer >= 0) continue;
    const long long rows_q = (q == s.nchunk - 1) ? s.nphys - (long long)q * s.chunk : s.chunk;
    if (s.jc > 0) {  // blocked columns: one interval per block (the blocks of one pass are sj apart)
      for (long long c = 0; c * s.jc < nj; ++c) {
        const long long w = (nj - c * s.jc < s.jc) ? nj - c * s.jc : s.jc;
        const long long lo = s.base[q].off + c * s.sj, ext = (nb - 1) * s.sb[q] + (rows_q - 1) * s.si[q] + w;
        out.push_back(Region{s.base[q].buf, lo * unit, (lo + ext) * unit, write});
      }
      continue;
    }
    const long long ext = (nb - 1) * s.sb[q] + (rows_q - 1) * s.si[q] + nj;
    out.push_back(Region{s.base[q].buf, s.base[q].off * unit, (s.base[q].off + ext) * unit, write});
  }
}

std::vector<Region> step_regions(const Step& s) {
  std::vector<Region> r;
  if (s.type == ST_STRIDED) {
    side_regions(s.in, s.B, s.J, false, 2, r);
    side_regions(s.out, s.B, s.J, true, 2, r);
  } else if (s.type == ST_R2C || s.type == ST_C2R) {  // real side in real units, complex side in 2 real units
    r.push_back(Region{s.real.buf, s.real.off, s.real.off + (s.rows - 1) * s.rpitch + s.n, s.type == ST_C2R});
    side_regions(s.cside, s.rows, 1, s.type == ST_R2C, 2, r);
  } else {
    for (int q = 0; q < s.npeers; ++q) {
      if (q == s.me) continue;
      if (!s.fused) r.push_back(Region{s.send[q].buf, s.send[q].off * 2, (s.send[q].off + s.scnt[q]) * 2, false});
      r.push_back(Region{s.recv[q].buf, s.recv[q].off * 2, (s.recv[q].off + s.rcnt[q]) * 2, true});
    }
  }
  return r;
}
}  // namespace

extern "C" int emu_check_schedule(const b200fft_plan_desc_t* d0, int inverse, int dealias) {
  if (!inverse && dealias == B200FFT_DEALIAS_2_3) dealias = B200FFT_DEALIAS_NONE;
  for (int rank = 0; rank < d0->nranks; ++rank) {
    b200fft_plan_desc_t d = *d0;
    d.rank = rank;
    Program pg;
    if (int rc = build_program(d, inverse, dealias, pg)) return rc;
    const int n = (int)pg.steps.size();
    std::vector<std::vector<char>> hb((size_t)n, std::vector<char>((size_t)n, 0));  // hb[i][j]: i happens before j
    int last_on[2] = {-1, -1};
    for (int j = 0; j < n; ++j) {
      const Step& s = pg.steps[(size_t)j];
      const int st = s.stream == 1 ? 1 : 0;
      std::vector<int> preds;
      if (last_on[st] >= 0) preds.push_back(last_on[st]);
      for (int ev : {s.wait_ev, s.wait_ev2}) {
        if (ev < 0) continue;
        int rec = -1;
        for (int i = 0; i < j; ++i)
          if (pg.steps[(size_t)i].rec_ev == ev) rec = i;
        if (ev == pg.fork_ev) continue;  // recorded at program start
        if (rec < 0) {
          std::fprintf(stderr, "schedule: rank %d step %d waits for event %d that no earlier step records\n", rank, j, ev);
          return 201;
        }
        preds.push_back(rec);
      }
      // a step on the second stream with no predecessor at all would run unordered against the caller's
      // earlier work: only allowed behind the fork event
      if (st == 1 && preds.empty() && !(pg.fork_ev >= 0 && s.wait_ev == pg.fork_ev)) {
        std::fprintf(stderr, "schedule: rank %d step %d starts the second stream without an event\n", rank, j);
        return 202;
      }
      for (int i : preds) {
        hb[(size_t)i][(size_t)j] = 1;
        for (int k = 0; k < n; ++k)
          if (hb[(size_t)k][(size_t)i]) hb[(size_t)k][(size_t)j] = 1;
      }
      last_on[st] = j;
    }
    // the caller's stream must end up behind everything: its last step (or a later one) follows every step
    for (int i = 0; i < n; ++i)
      if (i != last_on[0] && !hb[(size_t)i][(size_t)last_on[0]]) {
        std::fprintf(stderr, "schedule: rank %d step %d is not joined into the caller's stream\n", rank, i);
        return 203;
      }
    std::vector<std::vector<Region>> regs((size_t)n);
    for (int i = 0; i < n; ++i) regs[(size_t)i] = step_regions(pg.steps[(size_t)i]);
    for (int i = 0; i < n; ++i)
      for (int j = i + 1; j < n; ++j) {
        if (hb[(size_t)i][(size_t)j]) continue;
        for (const Region& a : regs[(size_t)i])
          for (const Region& c : regs[(size_t)j])
            if (a.buf == c.buf && (a.write || c.write) && a.lo < c.hi && c.lo < a.hi) {
              std::fprintf(stderr, "schedule: rank %d steps %d and %d touch buffer %d [%lld,%lld) / [%lld,%lld) unordered\n", rank, i, j,
                           a.buf, a.lo, a.hi, c.lo, c.hi);
              return 204;
            }
      }
  }
  return 0;
}

extern "C" {
// kernel launch geometry, for DESIGN.md / tests
int emu_strided_config(int precision, int n, int* T, int* TC, int* smem) {
  switch (n) {
#define X(nn, ...)                                                                        \
  case nn:                                                                                \
    if (precision == B200FFT_DOUBLE) {                                                    \
      using C = StridedCfg<double, Plan<__VA_ARGS__>>;                                    \
      *T = C::T; *TC = C::TC; *smem = C::SMEM;                                            \
    } else {                                                                              \
      using C = StridedCfg<float, Plan<__VA_ARGS__>>;                                     \
      *T = C::T; *TC = C::TC; *smem = C::SMEM;                                            \
    }                                                                                     \
    return 0;
    B200FFT_PLANS(X)
#undef X
    default:
      return -1;
  }
}
}
