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

static int g_emu_c2r_staging = 0;  // tests: 1 = the staging C2R kernel (C2RK) for every length

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
  if (d.cross_n > 0) {
    p.tw2 = table<real>(d.cross_n);
    p.tw2_div = d.cross_div;
  }
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

// the same kernel choice as csrc/k_rows.inc: four-CTA register budget for the 3/2-rule R2C lengths (no effect on the
// emulated arithmetic), the register-staged C2R kernel for half-lengths 256 ... 1536 and the staging kernel elsewhere
template <class real, bool FWD>
static int rows(const b200fft_rows_desc_t& d) {
  auto p = convert_rows<real>(d, table<real>(d.n), 1, FWD);
  switch (d.n / 2) {
#define X(n, ...)                                                                                            \
  case n:                                                                                                    \
    if (FWD) return emulate<R2CK<real, Plan<__VA_ARGS__>>>(p);                                               \
    else if (g_emu_c2r_staging || n < 256 || n > 1536) return emulate<C2RK<real, Plan<__VA_ARGS__>>>(p);     \
    else return emulate<C2RDK<real, Plan<__VA_ARGS__>>>(p);
    B200FFT_ROW_PLANS(X)
#undef X
    default:
      return -1;
  }
}

extern "C" {
int emu_set_c2r_staging(int v) {
  const int old = g_emu_c2r_staging;
  g_emu_c2r_staging = v;
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
// Step list of rank 0's program for tests: per step (type: 0 strided / 1 R2C / 2 C2R / 3 exchange, n, B or rows, J,
// row stride of the load side, of the store side, chunks of the load side, of the store side).  Returns the step
// count (and fills `out` when it has room for `cap` steps), negative on a build error.
int emu_plan_steps(const b200fft_plan_desc_t* d, int inverse, int dealias, long long* out, int cap) {
  Program pg;
  if (int rc = build_program(*d, inverse, dealias, pg)) return -rc;
  const int n = (int)pg.steps.size();
  if (out && cap >= n) {
    for (int i = 0; i < n; ++i) {
      const Step& s = pg.steps[(size_t)i];
      long long* o = out + 8 * i;
      const bool rows = s.type == ST_R2C || s.type == ST_C2R;
      o[0] = (long long)s.type;
      o[1] = s.n;
      o[2] = rows ? s.rows : s.B;
      o[3] = rows ? s.nk : s.J;
      o[4] = rows ? s.cside.sb[0] : s.in.si[0];
      o[5] = rows ? s.cside.sb[0] : s.out.si[0];
      o[6] = rows ? s.cside.nchunk : s.in.nchunk;
      o[7] = rows ? s.cside.nchunk : s.out.nchunk;
    }
  }
  return n;
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
    o.chunk = s.chunk; o.nchunk = s.nchunk; o.nphys = s.nphys;
    return o;
  };
  const size_t nsteps = pg[0].steps.size();
  for (int r = 1; r < P; ++r) if (pg[r].steps.size() != nsteps) return 90;
  for (size_t si = 0; si < nsteps; ++si) {
    for (int r = 0; r < P; ++r) {
      const Step& s = pg[r].steps[si];
      int rc = 0;
      if (s.type == ST_STRIDED) {
        b200fft_strided_desc_t d;
        std::memset(&d, 0, sizeof(d));
        d.precision = d0->precision; d.n = s.n; d.B = s.B; d.J = s.J; d.inverse = s.inverse;
        d.fold_mode = s.fold; d.scale = s.scale; d.in = side(r, s.in); d.out = side(r, s.out); d.mask = s.mask;
        d.cross_n = s.cross_n; d.cross_div = s.cross_div;
        rc = emu_exec_strided(&d);
      } else if (s.type == ST_R2C || s.type == ST_C2R) {
        b200fft_rows_desc_t d;
        std::memset(&d, 0, sizeof(d));
        d.precision = d0->precision; d.n = s.n; d.rows = s.rows; d.nk = s.nk; d.scale = s.scale;
        d.real_base = resolve(r, s.real, csz / 2); d.rpitch = s.rpitch; d.cside = side(r, s.cside);
        d.rm_period = s.rm_period; d.rm_block = s.rm_block; d.rm_planes = s.rm_planes;
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
          // it must be exactly the block peer w receives from this rank (member x.me of the communicator)
          const long long rows_q = (q == o.nchunk - 1) ? o.nphys - (long long)q * o.chunk : o.chunk;
          const long long nb = s.type == ST_STRIDED ? s.B : s.rows, nj = s.type == ST_STRIDED ? s.J : 1;
          const long long ext = (nb - 1) * o.sb[q] + (rows_q - 1) * o.si[q] + nj;
          if (o.base[q].buf != t.recv[x.me].buf || o.base[q].off < t.recv[x.me].off ||
              o.base[q].off + ext > t.recv[x.me].off + t.rcnt[x.me])
            return 115;
          if (o.base[q].off != t.recv[x.me].off || ext != t.rcnt[x.me]) return 120;
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
