// Plan programs: the step lists (row pass / strided pass / exchange) of the distributed slab,
// pencil and line transforms.  Host-only and CUDA-free so that libb200fft.so and the CPU emulator
// (tests/emu) build the very same programs.
#pragma once
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <deque>
#include <string>
#include <vector>

#include "../../include/b200fft.h"

namespace b200fft {
constexpr int PMAXP = B200FFT_MAXP;

inline std::string& plan_err() {
  static thread_local std::string e;
  return e;
}
inline int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  plan_err() = buf;
  return code;
}
}  // namespace b200fft

namespace b200fft {
// ================================================================================================
// plan programs
// ================================================================================================
// W0..W2 may be written by peers (peer-mapped transports export exactly these three); W3 is a local
// send buffer of the pipelined pencil programs
enum Buf { BUF_IN = 0, BUF_OUT = 1, BUF_W0 = 2, BUF_W1 = 3, BUF_W2 = 4, BUF_W3 = 5, NBUF = 6 };
constexpr int NWORK = NBUF - BUF_W0;  // work buffers of a plan
constexpr int NPEERBUF = 3;           // ... of which the first three are mapped by the peers

struct Ref {
  int buf = BUF_IN;
  long long off = 0;
  int peer = -1;  // >= 0: the buffer of that rank (fused transport: stores land in the peer's memory)
};

struct SideT {
  Ref base[PMAXP];
  long long sb[PMAXP] = {0};
  long long si[PMAXP] = {0};
  int chunk = 1, nchunk = 1, nphys = 1;
};

enum StepType { ST_STRIDED, ST_R2C, ST_C2R, ST_EXCH };

struct Step {
  StepType type = ST_STRIDED;
  // strided
  int n = 0;
  long long B = 1;
  int J = 1;
  int inverse = 0, fold = 0;
  double scale = 1.0;
  SideT in, out;
  b200fft_mask_t mask;
  int cross_n = 0, cross_div = 1;  // four-step cross twiddle (b200fft_strided_desc_t)
  // rows
  long long rows = 0;
  int nk = 0;
  Ref real;
  long long rpitch = 0;
  SideT cside;
  long long rm_period = 0, rm_block = 0, rm_planes = 0;  // complex-side row permutation (b200fft_rows_desc_t)
  // scheduling: stream 0 = the caller's stream (FFT passes), 1 = the plan's communication stream;
  // wait_ev: event the step's stream waits for before the step (-1 none); rec_ev: event recorded
  // after it.  Lets the exchange of one chunk overlap the FFT passes of the next.
  int stream = 0, wait_ev = -1, wait_ev2 = -1, rec_ev = -1;
  int pass = 0;  // logical pass of the transform this step belongs to (chunks of one pass share it)
  // exchange
  int comm = 0;  // 0: world, 1: comm0, 2: comm1
  int npeers = 0, me = 0;
  Ref send[PMAXP], recv[PMAXP];
  long long scnt[PMAXP] = {0}, rcnt[PMAXP] = {0};
  // copy-engine transport: where this rank's block lands in peer q's receive buffer (the peer's
  // recv[me]), whether the step is the first exchange of the program (waits for the peers' credits)
  // and whether a pass is the last reader of received data (returns the credits)
  Ref rpeer[PMAXP];
  int first_exch = 0, last_reader = 0;
  // fused (peer-store) transport: the producing FFT pass has already stored the blocks into the
  // peers' receive buffers, so the exchange step moves no data and only publishes / awaits the
  // sequence flags (`fused`); the first pass of a program that stores into peer memory must hold the
  // peers' credits first (`wait_credits`)
  int fused = 0, wait_credits = 0;
};

struct Program {
  // a deque: the builders keep references to steps while they append further ones (push_back on a
  // deque leaves references to existing elements valid; a vector's reallocation does not)
  std::deque<Step> steps;
  long long need[NBUF] = {0, 0, 0, 0, 0, 0};  // complex elements needed in W0..W3
  int nevents = 0;
  int fork_ev = -1;  // >= 0: recorded on the caller's stream when the program starts and awaited by the
                     // second stream (programs whose first step runs there)
  bool built = false;
  int error = 0;
  std::string errmsg;
};

inline b200fft_mask_t mask_off() {
  b200fft_mask_t m;
  std::memset(&m, 0, sizeof(m));
  m.jdiv = 1;
  m.i_lo = m.b_lo = m.jq_lo = m.jr_lo = 1;
  m.i_hi = m.b_hi = m.jq_hi = m.jr_hi = 0;
  return m;
}

// band of indices zeroed by the 2/3-rule along an axis of (unpadded) size N:
// keep iff |k| < kmax, kmax = 2/3*(N//2+1)  (slab.py:191-197)
inline void band(long long N, bool half, int& lo, int& hi) {
  const double kmax = 2. / 3. * (double)(N / 2 + 1);
  const int l = (int)std::ceil(kmax);
  lo = l;
  hi = half ? 0x3fffffff : (int)N - l;
}

inline SideT nat(int buf, long long off, long long sb, long long si, int nphys) {
  SideT s;
  s.base[0].buf = buf;
  s.base[0].off = off;
  s.sb[0] = sb;
  s.si[0] = si;
  s.chunk = nphys;
  s.nchunk = 1;
  s.nphys = nphys;
  return s;
}

struct Builder {
  Program& pg;
  int next_pass = 0, fixed = -1;  // pass id given to new steps: `fixed` when >= 0, else a running index
  explicit Builder(Program& p) : pg(p) {}
  int pass_id() { return fixed >= 0 ? fixed : next_pass++; }
  void use(int buf, long long end) {
    if (buf >= BUF_W0 && end > pg.need[buf]) pg.need[buf] = end;
  }
  Step& strided(int n, long long B, long long J, int inverse, const SideT& in, const SideT& out, int fold = 0,
                double scale = 1.0) {
    Step s;
    s.type = ST_STRIDED;
    s.n = n;
    s.B = B;
    s.J = (int)J;
    s.inverse = inverse;
    s.fold = fold;
    s.scale = scale;
    s.in = in;
    s.out = out;
    s.mask = mask_off();
    s.pass = pass_id();
    pg.steps.push_back(s);
    return pg.steps.back();
  }
  Step& rows(bool fwd, long long rows, int n, int nk, int realbuf, const SideT& cside, double scale = 1.0) {
    Step s;
    s.type = fwd ? ST_R2C : ST_C2R;
    s.rows = rows;
    s.n = n;
    s.nk = nk;
    s.real.buf = realbuf;
    s.real.off = 0;
    s.rpitch = n;
    s.cside = cside;
    s.scale = scale;
    s.mask = mask_off();
    s.pass = pass_id();
    pg.steps.push_back(s);
    return pg.steps.back();
  }
  Step& exch(int comm, int npeers, int me) {
    Step s;
    s.type = ST_EXCH;
    s.comm = comm;
    s.npeers = npeers;
    s.me = me;
    s.mask = mask_off();
    s.pass = pass_id();
    pg.steps.push_back(s);
    return pg.steps.back();
  }
};

inline int ipad(double p, long long x) { return (int)(p * (double)x); }

// Pipeline depth of a chunked exchange: the local extent `ext` is cut into C equal chunks; chunk c's
// exchange runs on the communication stream while the FFT passes of chunk c+1 run.  Auto: the
// largest C in {8, 4, 2} that divides ext and keeps every per-peer message above a floor.
//  * NCCL exchanges are kernels that compete with the FFT grids for SMs -- measured at 1024^3 double
//    on 4 GPUs: 10.6 ms (1 chunk), 10.0 (2), 11.1 (4), 11.7 (8) -- so NCCL plans stop at 2 chunks
//    (floor 4 MB: below that launch latency, not NVLink bandwidth, sets the exchange time).
//  * Copy-engine (P2P) exchanges leave the SMs alone but pay ~25 us per queued copy + flag
//    (2 GPUs, 8 chunks: +0.19 ms over 1 chunk; 8 GPUs, 8 chunks of 16.8 MB x 7 peers: 2.58 ms
//    against 1.07 ms of pure transfer).  A 48 MB floor keeps that overhead under ~30% of the
//    transfer time: 8 chunks at P = 2 and 4, 2 chunks at P = 8 for 1024^3 double.
inline int pick_chunks(int requested, long long ext, long long peer_msg_bytes, bool p2p, bool store = false) {
  if (store && requested <= 0) return 1;  // the producing pass is the transfer: nothing to pipeline by default
  if (requested > 0) {
    int c = requested;
    while (c > 1 && ext % c) --c;
    return c < 1 ? 1 : c;
  }
  const int max_auto = p2p ? 8 : 2;
  const long long floor_bytes = p2p ? (48ll << 20) : (4ll << 20);
  for (int c : {8, 4, 2})
    if (c <= max_auto && ext % c == 0 && peer_msg_bytes / c >= floor_bytes) return c;
  return 1;
}

// Pipeline depth of pencil and line plans when the caller leaves it open (d.chunks == 0): two chunks with the
// copy-engine transport once a rank exchanges 256 MB or more per step, else one.  8 GPUs, profiles/r02_multi_8:
// pencil X 1024^3 double (1.07 GB per rank) 6.76 / 5.86 / 6.44 ms at 1 / 2 / 4 chunks, pencil Y 2048^3 single
// (4.3 GB) 32.0 / 27.0 / 27.1, pencil X 512^3 double (134 MB) 1.19 / 1.20 / 1.33, line 16384^2 single (134 MB)
// 1.78 / 1.89 / 2.20; NCCL exchanges lose with any chunking there (7.68 / 8.70 / 9.25 ms).
inline int grid_chunks(const b200fft_plan_desc_t& d, long long local_complex_elems) {
  if (d.transport == B200FFT_TRANSPORT_STORE) return 1;
  if (d.chunks > 0) return d.chunks;
  const long long csz = d.precision == B200FFT_DOUBLE ? 16 : 8;
  return (d.transport == B200FFT_TRANSPORT_P2P && local_complex_elems * csz >= (256ll << 20)) ? 2 : 1;
}

// kz pipeline: number of kz ranges (each at least 8 entries wide so that tiles stay full)
inline int kz_chunks(int requested, long long Nf) {
  long long c = requested > 0 ? requested : 4;
  if (c > Nf / 8) c = Nf / 8;
  return c < 1 ? 1 : (int)c;
}

// world rank of peer q of communicator `comm` (0: world, 1: comm0 = ranks with equal rank / P1,
// 2: comm1 = ranks with equal rank % P1; pencil.py:184-195) as seen from world rank `me`
inline int world_rank(const b200fft_plan_desc_t& d, int comm, int me, int q) {
  if (comm == 0) return q;
  if (comm == 1) return (me / d.P1) * d.P1 + q;
  return q * d.P1 + (me % d.P1);
}

// Peer-mapped transports of the pencil / line programs (the slab programs set these fields as they are
// built).  Copy engines: first_exch / last_reader bracket the buffers' hand-over between transforms.
// Fused stores: the pass right before an exchange step writes the per-peer blocks into its send buffer;
// point those store bases at the place the block would be copied to -- `rpeer` in the peer's memory --
// and the exchange step is left with the flags only.
inline int finish_peer_mapped(const b200fft_plan_desc_t& d, Program& pg) {
  const bool store = d.transport == B200FFT_TRANSPORT_STORE;
  bool first = true, credits = false;
  for (size_t i = 0; i < pg.steps.size(); ++i) {
    Step& x = pg.steps[i];
    if (x.type != ST_EXCH) continue;
    x.first_exch = first;
    first = false;
    // arrival is awaited on the plan's wait stream: the consuming pass needs an event to follow it
    if (x.rec_ev < 0 && i + 1 < pg.steps.size() && pg.steps[i + 1].wait_ev < 0) {
      x.rec_ev = pg.nevents++;
      pg.steps[i + 1].wait_ev = x.rec_ev;
    }
    if (!store) continue;
    if (i == 0) return fail(B200FFT_ERR_ARG, "exchange without a producing pass");
    Step& y = pg.steps[i - 1];
    SideT& o = (y.type == ST_STRIDED) ? y.out : y.cside;
    if (y.type == ST_EXCH || y.type == ST_C2R || o.nchunk != x.npeers) return fail(B200FFT_ERR_ARG, "producing pass does not match its exchange");
    for (int q = 0; q < x.npeers; ++q) {
      if (q == x.me) continue;
      if (o.base[q].buf != x.send[q].buf || o.base[q].off != x.send[q].off) return fail(B200FFT_ERR_ARG, "producing pass does not fill the send blocks");
      o.base[q] = x.rpeer[q];
      o.base[q].peer = world_rank(d, x.comm, d.rank, q);
    }
    x.fused = 1;
    if (!credits) y.wait_credits = 1;
    credits = true;
  }
  if (!first) pg.steps.back().last_reader = 1;  // every reader of received data has run by then
  return 0;
}

inline int build_slab(const b200fft_plan_desc_t& d, int inverse, int dealias, Program& pg) {
  Builder b(pg);
  const bool padded = dealias == B200FFT_DEALIAS_3_2;
  const bool masked = inverse && dealias == B200FFT_DEALIAS_2_3;
  const double p = padded ? d.padsize : 1.0;
  const int P = d.nranks, me = d.rank;
  const long long N0 = d.N[0], N1 = d.N[1], N2 = d.N[2];
  // slab.C2C (slab.py:538-825) reuses every R2C shape with Nf = N[2] (slab.py:565-567); its z pass is
  // a contiguous-row C2C, its truncations fold in y and z at P > 1 (copy_from_padded, :816-823) and
  // keep mode -N/2 in x (:796-797) and everywhere at P == 1 (the `ks` gather, :735-738)
  const bool c2c = d.kind == B200FFT_SLAB_C2C;
  const long long Np0 = N0 / P, Np1 = N1 / P, Nf = c2c ? N2 : N2 / 2 + 1;
  const int pN0 = ipad(p, N0), pNp0 = ipad(p, Np0), pN1 = ipad(p, N1), pN2 = ipad(p, N2);
  if (padded && P > 1 && P > N0 / 2)  // slab.py:311,446
    return fail(B200FFT_ERR_ARG, "Number of processors cannot be larger than N[0]//2 for 3/2-rule");
  if (padded && (long long)pNp0 * P != pN0)  // int(padsize * N0 / P) must tile the padded mesh (real_shape_padded, slab.py:104-107)
    return fail(B200FFT_ERR_ARG, "3/2-rule: padsize * N[0] / ranks must be an integer");
  const double p3 = p * p * p;
  const long long blk = (long long)pNp0 * Np1 * Nf;
  const long long csz = d.precision == B200FFT_DOUBLE ? 16 : 8;
  // fused transport: the y (forward) / x (inverse) pass stores each peer's block into that peer's
  // receive buffer -- exactly where the copy-engine transport's DMA would put it (`rpeer`)
  const bool store = d.transport == B200FFT_TRANSPORT_STORE && P > 1;
  const bool p2p = (d.transport == B200FFT_TRANSPORT_P2P || store) && P > 1;  // peers write into plan-owned buffers only
  // Pipeline direction.  AUTO: kz ranges for plain / 2/3-rule R2C transforms over the copy engines once a per-peer
  // message reaches 96 MB, in 4 chunks while a chunk's message stays >= 64 MB, else 2.  1024^3 double, round trip in
  // ms, x planes (depth of pick_chunks) against kz ranges (profiles/r02_flags_2, r02_sweep_4, r02_final_8):
  // 2 GPUs 13.07 / 12.88 (4), 4 GPUs 8.69 / 8.01 (4) / 8.45 (2), 8 GPUs 5.16 / 5.02 (2) / 5.67 (4); with the
  // 3/2-rule the x-plane pipeline wins (4 GPUs 15.06 / 15.79, 8 GPUs 8.28 / 9.00) and keeps the default there.
  const bool auto_kz = d.pipeline == B200FFT_PIPELINE_AUTO && P > 1 && !padded && !c2c && d.transport == B200FFT_TRANSPORT_P2P &&
                       blk * csz >= (96ll << 20);
  const bool use_kz = d.pipeline == B200FFT_PIPELINE_KZ || auto_kz;
  const int kz_req = (auto_kz && d.chunks <= 0) ? ((blk * csz) / 4 >= (64ll << 20) ? 4 : 2) : d.chunks;
  const int yfold = !padded ? 0 : (c2c && P == 1) ? 2 : 1;
  const int xfold = !padded ? 0 : c2c ? 2 : 1;
  const int zfold = !padded ? 0 : (P == 1) ? 2 : 1;
  // z pass over `rows` rows starting at row `row0` of the caller's array; the spectrum side is
  // [rows][Nf] at `coff` of buffer `cbuf`
  auto zfwd = [&](long long rows, long long row0, int cbuf, long long coff) -> Step& {
    if (c2c)
      return b.strided(pN2, rows, 1, 0, nat(BUF_IN, row0 * pN2, pN2, 1, pN2), nat(cbuf, coff, Nf, 1, (int)Nf), zfold);
    Step& z = b.rows(true, rows, pN2, (int)Nf, BUF_IN, nat(cbuf, coff, Nf, 1, (int)Nf));
    z.real.off = row0 * pN2;
    return z;
  };
  auto zinv = [&](long long rows, long long row0, int cbuf, long long coff, double scale) -> Step& {
    if (c2c)
      return b.strided(pN2, rows, 1, 1, nat(cbuf, coff, Nf, 1, (int)Nf), nat(BUF_OUT, row0 * pN2, pN2, 1, pN2), 0, scale);
    Step& z = b.rows(false, rows, pN2, (int)Nf, BUF_OUT, nat(cbuf, coff, Nf, 1, (int)Nf), scale);
    z.real.off = row0 * pN2;
    return z;
  };
  // (`zy_groups(x0, xn, emit)` calls emit(first plane, plane count) for the z and y passes over local x planes
  // [x0, x0 + xn): one group.  Cutting them into L2-sized groups -- z(g), y(g), z(g+1), ... -- was measured in
  // round 2 and lost in every form, profiles/r02_single/ab_single.txt: the small launches cost more than the
  // L2 hits save.)
  auto zy_groups = [&](long long x0, long long xn, auto&& emit) { emit(x0, xn); };
  // Single-rank R2C plans: "y-blocked" intermediate.  In the natural layout [x][y][kz] the x pass works on rows
  // N1*Nf elements apart (8.4 MB at 1024^3 double): one 2 MB page per row segment on BOTH of its sides, 0.57 of
  // the HBM figure where the near-stride y pass reaches 0.76 (profiles/r01c, r02_single).  Nothing forces that
  // layout between the passes.  The z pass writes (forward) / reads (inverse) the intermediate as
  //     W0 = [y block c][x][y in block][kz],   NC = at most 16 blocks of YB = pN1 / NC rows  (row map of the row kernels)
  // the x pass runs IN PLACE on it with B = NC, J = YB*Nf -- rows YB*Nf elements apart (525 KB at 1024^3) --
  // and the y pass, now last (forward) / first (inverse), gathers its column from the NC blocks (the per-chunk
  // bases every exchange layout already uses) and has the caller's array [x][y][kz] on its other side, where
  // rows along y are Nf elements apart.  No pass has a far side; the order of the x and y passes is free
  // because the truncations / pads of the 3/2-rule act per axis.  One work buffer instead of two for the 3/2-rule.
  const bool yblock = P == 1 && !c2c && d.layout != B200FFT_LAYOUT_NATURAL;
  if (yblock) {
    int NC = 1;
    while (NC < PMAXP && NC * 2 <= PMAXP && pN1 % (NC * 2) == 0) NC *= 2;
    const long long YB = pN1 / NC;
    const long long blkc = (long long)pN0 * YB * Nf;  // one y block: [x][y in block][kz]
    b.use(BUF_W0, NC * blkc);
    SideT yside;  // a y column of the intermediate for batch entry b = x: chunk c at c*blkc + b*YB*Nf + (y - c*YB)*Nf
    yside.chunk = (int)YB;
    yside.nchunk = NC;
    yside.nphys = pN1;
    for (int c = 0; c < NC; ++c) {
      yside.base[c].buf = BUF_W0;
      yside.base[c].off = c * blkc;
      yside.sb[c] = YB * Nf;
      yside.si[c] = Nf;
    }
    auto rowmap = [&](Step& z) {
      z.rm_period = pN1;
      z.rm_block = YB;
      z.rm_planes = pN0;
    };
    if (!inverse) {  // slab.py:366-387
      b.fixed = 0;
      Step& z = b.rows(true, (long long)pN0 * pN1, pN2, (int)Nf, BUF_IN, nat(BUF_W0, 0, Nf, 1, (int)Nf));
      rowmap(z);
      b.fixed = 1;
      b.strided(pN0, NC, YB * Nf, 0, nat(BUF_W0, 0, blkc, YB * Nf, pN0), nat(BUF_W0, 0, blkc, YB * Nf, (int)N0), xfold);
      b.fixed = 2;
      b.strided(pN1, N0, Nf, 0, yside, nat(BUF_OUT, 0, N1 * Nf, Nf, (int)N1), yfold, padded ? 1.0 / p3 : 1.0);
    } else {  // slab.py:247-268
      const double scale = (padded ? p3 : 1.0) / ((double)pN0 * (double)pN1 * (double)pN2);
      b.fixed = 0;
      Step& sy = b.strided(pN1, N0, Nf, 1, nat(BUF_IN, 0, N1 * Nf, Nf, (int)N1), yside);
      if (masked) {  // rows = ky, batch entries = kx, columns = kz
        sy.mask.on = 1;
        sy.mask.jdiv = 0x3fffffff;
        band(N1, false, sy.mask.i_lo, sy.mask.i_hi);
        band(N0, false, sy.mask.b_lo, sy.mask.b_hi);
        band(N2, true, sy.mask.jr_lo, sy.mask.jr_hi);
      }
      b.fixed = 1;
      b.strided(pN0, NC, YB * Nf, 1, nat(BUF_W0, 0, blkc, YB * Nf, (int)N0), nat(BUF_W0, 0, blkc, YB * Nf, pN0));
      b.fixed = 2;
      Step& z = b.rows(false, (long long)pN0 * pN1, pN2, (int)Nf, BUF_OUT, nat(BUF_W0, 0, Nf, 1, (int)Nf), scale);
      rowmap(z);
    }
    return 0;
  }
  if (!inverse) {
    if (P == 1) {
      if (!padded) {  // slab.py:366-370
        zy_groups(0, N0, [&](long long g0, long long gn) {
          b.fixed = 0;
          zfwd(gn * N1, g0 * N1, BUF_OUT, g0 * N1 * Nf);
          b.fixed = 1;
          b.strided((int)N1, gn, Nf, 0, nat(BUF_OUT, g0 * N1 * Nf, N1 * Nf, Nf, (int)N1), nat(BUF_OUT, g0 * N1 * Nf, N1 * Nf, Nf, (int)N1));
        });
        b.fixed = 2;
        b.strided((int)N0, 1, N1 * Nf, 0, nat(BUF_OUT, 0, 0, N1 * Nf, (int)N0), nat(BUF_OUT, 0, 0, N1 * Nf, (int)N0));
      } else {  // slab.py:371-387
        b.use(BUF_W0, (long long)pN0 * pN1 * Nf);
        b.use(BUF_W1, (long long)pN0 * N1 * Nf);
        zy_groups(0, pN0, [&](long long g0, long long gn) {
          b.fixed = 0;
          zfwd(gn * pN1, g0 * pN1, BUF_W0, g0 * pN1 * Nf);
          b.fixed = 1;
          b.strided(pN1, gn, Nf, 0, nat(BUF_W0, g0 * pN1 * Nf, pN1 * Nf, Nf, pN1), nat(BUF_W1, g0 * N1 * Nf, N1 * Nf, Nf, (int)N1), yfold);
        });
        b.fixed = 2;
        b.strided(pN0, 1, N1 * Nf, 0, nat(BUF_W1, 0, 0, N1 * Nf, pN0), nat(BUF_OUT, 0, 0, N1 * Nf, (int)N0), xfold, 1.0 / p3);
      }
    } else if (use_kz) {  // slab.py:389-483, three-stage pipeline
      // one z pass; then per kz range c:  y(c) -> exchange(c) -> x(c).  Send and receive buffers are
      // chunk-major -- [c][peer q][x][y][kz in c] -- so every (chunk, peer) message is contiguous.
      // exchange(c) (second stream) overlaps y(c+1..) before it and x(..c-1) after it; with the fused
      // transport the y passes ARE the transfer (NVLink-bound) and run on the second stream beside the
      // HBM-bound x passes.
      const int recvbuf = BUF_W2;
      const int C = kz_chunks(kz_req, Nf);
      const long long kc = Nf / C;
      b.use(BUF_W0, (long long)pNp0 * pN1 * Nf);
      if (!store) b.use(BUF_W1, P * blk);
      b.use(recvbuf, P * blk);
      b.fixed = 0;
      Step& z = zfwd((long long)pNp0 * pN1, 0, BUF_W0, 0);
      int z_ev = -1;
      if (store) z_ev = z.rec_ev = pg.nevents++;
      std::vector<int> xev((size_t)C);
      for (int c = 0; c < C; ++c) {
        const long long k0 = c * kc, kcc = (c == C - 1) ? Nf - k0 : kc;
        const long long coff = (long long)P * pNp0 * Np1 * k0, blkc = (long long)pNp0 * Np1 * kcc;
        SideT o;
        o.chunk = (int)Np1;
        o.nchunk = P;
        o.nphys = (int)N1;
        for (int q = 0; q < P; ++q) {
          o.base[q].buf = (q == me) ? recvbuf : BUF_W1;
          o.base[q].off = coff + q * blkc;
          if (store && q != me) {
            o.base[q].buf = recvbuf;
            o.base[q].off = coff + me * blkc;
            o.base[q].peer = q;
          }
          o.sb[q] = Np1 * kcc;
          o.si[q] = kcc;
        }
        b.fixed = 1;
        Step& y = b.strided(pN1, pNp0, kcc, 0, nat(BUF_W0, k0, pN1 * Nf, Nf, pN1), o, yfold);
        y.wait_credits = store && c == 0;
        if (store) {
          y.stream = 1;
          if (c == 0) y.wait_ev = z_ev;
        }
        y.rec_ev = pg.nevents++;
        b.fixed = 2;
        Step& x = b.exch(0, P, me);
        x.stream = 1;
        x.wait_ev = y.rec_ev;
        x.rec_ev = pg.nevents++;
        xev[(size_t)c] = x.rec_ev;
        x.first_exch = (c == 0);
        x.fused = store;
        for (int q = 0; q < P; ++q) {
          x.send[q].buf = BUF_W1; x.send[q].off = coff + q * blkc; x.scnt[q] = blkc;
          x.recv[q].buf = recvbuf; x.recv[q].off = coff + q * blkc; x.rcnt[q] = blkc;
          x.rpeer[q].buf = recvbuf; x.rpeer[q].off = coff + me * blkc;
        }
      }
      for (int c = 0; c < C; ++c) {
        const long long k0 = c * kc, kcc = (c == C - 1) ? Nf - k0 : kc;
        const long long coff = (long long)P * pNp0 * Np1 * k0, blkc = (long long)pNp0 * Np1 * kcc;
        SideT g;
        g.chunk = pNp0;
        g.nchunk = P;
        g.nphys = pN0;
        for (int q = 0; q < P; ++q) {
          g.base[q].buf = recvbuf;
          g.base[q].off = coff + q * blkc;
          g.sb[q] = kcc;
          g.si[q] = Np1 * kcc;
        }
        b.fixed = 3;
        Step& fx = b.strided(pN0, Np1, kcc, 0, g, nat(BUF_OUT, k0, Nf, Np1 * Nf, (int)N0), xfold, padded ? 1.0 / p3 : 1.0);
        fx.wait_ev = xev[(size_t)c];
        fx.last_reader = (c == C - 1);
      }
    } else {  // slab.py:389-483
      // z and y passes of chunk c (a range of local x planes) run while chunk c-1 is exchanged
      const int recvbuf = (padded || p2p) ? BUF_W2 : BUF_OUT;
      const int C = pick_chunks(d.chunks, pNp0, blk * csz, p2p, store);
      const long long xc = pNp0 / C;
      b.use(BUF_W0, (long long)pNp0 * pN1 * Nf);
      if (!store) b.use(BUF_W1, P * blk);  // send buffer
      b.use(recvbuf, P * blk);
      for (int c = 0; c < C; ++c) {
        const long long x0 = c * xc;
        int y_ev = -1;
        zy_groups(x0, xc, [&](long long g0, long long gn) {
          b.fixed = 0;
          const size_t zi = pg.steps.size();
          zfwd(gn * pN1, g0 * pN1, BUF_W0, g0 * pN1 * Nf);
          SideT o;
          o.chunk = (int)Np1;
          o.nchunk = P;
          o.nphys = (int)N1;
          for (int q = 0; q < P; ++q) {
            o.base[q].buf = (q == me) ? recvbuf : BUF_W1;
            o.base[q].off = q * blk + g0 * Np1 * Nf;
            if (store && q != me) {  // block `me` of peer q's receive buffer
              o.base[q].buf = recvbuf;
              o.base[q].off = me * blk + g0 * Np1 * Nf;
              o.base[q].peer = q;
            }
            o.sb[q] = Np1 * Nf;
            o.si[q] = Nf;
          }
          b.fixed = 1;
          Step& y = b.strided(pN1, gn, Nf, 0, nat(BUF_W0, g0 * pN1 * Nf, pN1 * Nf, Nf, pN1), o, yfold);
          y.wait_credits = store && c == 0 && g0 == x0;
          if (g0 + gn == x0 + xc) y_ev = y.rec_ev = pg.nevents++;  // the exchange follows the chunk's last group
        });
        b.fixed = 2;
        Step& x = b.exch(0, P, me);
        x.stream = 1;
        x.wait_ev = y_ev;
        x.rec_ev = pg.nevents++;
        x.first_exch = (c == 0);
        x.fused = store;
        for (int q = 0; q < P; ++q) {
          x.send[q].buf = BUF_W1; x.send[q].off = q * blk + x0 * Np1 * Nf; x.scnt[q] = xc * Np1 * Nf;
          x.recv[q].buf = recvbuf; x.recv[q].off = q * blk + x0 * Np1 * Nf; x.rcnt[q] = xc * Np1 * Nf;
          x.rpeer[q].buf = recvbuf; x.rpeer[q].off = me * blk + x0 * Np1 * Nf;
        }
      }
      const int last_ev = pg.nevents - 1;  // exchanges run in order on one stream
      b.fixed = 3;
      Step& fx = b.strided(pN0, 1, Np1 * Nf, 0, nat(recvbuf, 0, 0, Np1 * Nf, pN0), nat(BUF_OUT, 0, 0, Np1 * Nf, (int)N0),
                           xfold, padded ? 1.0 / p3 : 1.0);
      fx.wait_ev = last_ev;
      fx.last_reader = 1;
    }
  } else {
    const double scale = (padded ? p3 : 1.0) / ((double)pN0 * (double)pN1 * (double)pN2);
    if (P == 1) {  // slab.py:247-268
      Step& sx = b.strided(pN0, 1, N1 * Nf, 1, nat(BUF_IN, 0, 0, N1 * Nf, (int)N0), nat(BUF_W0, 0, 0, N1 * Nf, pN0));
      b.use(BUF_W0, (long long)pN0 * N1 * Nf);
      if (masked) {
        sx.mask.on = 1;
        sx.mask.jdiv = (int)Nf;
        band(N0, false, sx.mask.i_lo, sx.mask.i_hi);
        band(N1, false, sx.mask.jq_lo, sx.mask.jq_hi);
        band(N2, !c2c, sx.mask.jr_lo, sx.mask.jr_hi);
      }
      if (!padded) {
        zy_groups(0, N0, [&](long long g0, long long gn) {
          b.fixed = 1;
          b.strided((int)N1, gn, Nf, 1, nat(BUF_W0, g0 * N1 * Nf, N1 * Nf, Nf, (int)N1), nat(BUF_W0, g0 * N1 * Nf, N1 * Nf, Nf, (int)N1));
          b.fixed = 2;
          zinv(gn * N1, g0 * N1, BUF_W0, g0 * N1 * Nf, scale);
        });
      } else {
        b.use(BUF_W1, (long long)pN0 * pN1 * Nf);
        zy_groups(0, pN0, [&](long long g0, long long gn) {
          b.fixed = 1;
          b.strided(pN1, gn, Nf, 1, nat(BUF_W0, g0 * N1 * Nf, N1 * Nf, Nf, (int)N1), nat(BUF_W1, g0 * pN1 * Nf, pN1 * Nf, Nf, pN1));
          b.fixed = 2;
          zinv(gn * pN1, g0 * pN1, BUF_W1, g0 * pN1 * Nf, scale);
        });
      }
    } else if (use_kz) {  // slab.py:270-345, three-stage pipeline
      // per kz range c:  x(c) -> exchange(c) -> y(c);  then one z pass (mirror of the forward program)
      const int C = kz_chunks(kz_req, Nf);
      const long long kc = Nf / C;
      const int ybuf = BUF_W2;
      if (!store) b.use(BUF_W0, P * blk);
      b.use(BUF_W1, P * blk);
      b.use(ybuf, (long long)pNp0 * pN1 * Nf);
      if (store) pg.fork_ev = pg.nevents++;  // the x passes (the transfer) run on the second stream
      std::vector<int> xev((size_t)C);
      for (int c = 0; c < C; ++c) {
        const long long k0 = c * kc, kcc = (c == C - 1) ? Nf - k0 : kc;
        const long long coff = (long long)P * pNp0 * Np1 * k0, blkc = (long long)pNp0 * Np1 * kcc;
        SideT o;
        o.chunk = pNp0;
        o.nchunk = P;
        o.nphys = pN0;
        for (int q = 0; q < P; ++q) {
          o.base[q].buf = (q == me) ? BUF_W1 : BUF_W0;
          o.base[q].off = coff + q * blkc;
          if (store && q != me) {
            o.base[q].buf = BUF_W1;
            o.base[q].off = coff + me * blkc;
            o.base[q].peer = q;
          }
          o.sb[q] = kcc;
          o.si[q] = Np1 * kcc;
        }
        b.fixed = 0;
        Step& sx = b.strided(pN0, Np1, kcc, 1, nat(BUF_IN, k0, Nf, Np1 * Nf, (int)N0), o);
        if (masked) {  // batch index = local ky, column index = kz - k0
          sx.mask.on = 1;
          sx.mask.jdiv = 0x3fffffff;
          band(N0, false, sx.mask.i_lo, sx.mask.i_hi);
          band(N1, false, sx.mask.b_lo, sx.mask.b_hi);
          sx.mask.b_off = (int)(me * Np1);
          band(N2, !c2c, sx.mask.jr_lo, sx.mask.jr_hi);
          sx.mask.jr_off = (int)k0;
        }
        sx.wait_credits = store && c == 0;
        if (store) {
          sx.stream = 1;
          if (c == 0) sx.wait_ev = pg.fork_ev;
        }
        sx.rec_ev = pg.nevents++;
        b.fixed = 1;
        Step& x = b.exch(0, P, me);
        x.stream = 1;
        x.wait_ev = sx.rec_ev;
        x.rec_ev = pg.nevents++;
        xev[(size_t)c] = x.rec_ev;
        x.first_exch = (c == 0);
        x.fused = store;
        for (int q = 0; q < P; ++q) {
          x.send[q].buf = BUF_W0; x.send[q].off = coff + q * blkc; x.scnt[q] = blkc;
          x.recv[q].buf = BUF_W1; x.recv[q].off = coff + q * blkc; x.rcnt[q] = blkc;
          x.rpeer[q].buf = BUF_W1; x.rpeer[q].off = coff + me * blkc;
        }
      }
      for (int c = 0; c < C; ++c) {
        const long long k0 = c * kc, kcc = (c == C - 1) ? Nf - k0 : kc;
        const long long coff = (long long)P * pNp0 * Np1 * k0, blkc = (long long)pNp0 * Np1 * kcc;
        SideT g;
        g.chunk = (int)Np1;
        g.nchunk = P;
        g.nphys = (int)N1;
        for (int q = 0; q < P; ++q) {
          g.base[q].buf = BUF_W1;
          g.base[q].off = coff + q * blkc;
          g.sb[q] = Np1 * kcc;
          g.si[q] = kcc;
        }
        b.fixed = 2;
        Step& y = b.strided(pN1, pNp0, kcc, 1, g, nat(ybuf, k0, pN1 * Nf, Nf, pN1));
        y.wait_ev = xev[(size_t)c];
      }
      b.fixed = 3;
      zinv((long long)pNp0 * pN1, 0, ybuf, 0, scale).last_reader = 1;
    } else {  // slab.py:270-345
      // x pass, then per chunk of local x planes: exchange (communication stream) -> y and z
      // passes; the exchange of chunk c+1 overlaps the passes of chunk c
      SideT o;
      o.chunk = pNp0;
      o.nchunk = P;
      o.nphys = pN0;
      for (int q = 0; q < P; ++q) {
        o.base[q].buf = (q == me) ? BUF_W1 : BUF_W0;
        o.base[q].off = q * blk;
        if (store && q != me) {  // block `me` of peer q's receive buffer
          o.base[q].buf = BUF_W1;
          o.base[q].off = me * blk;
          o.base[q].peer = q;
        }
        o.sb[q] = 0;
        o.si[q] = Np1 * Nf;
      }
      Step& sx = b.strided(pN0, 1, Np1 * Nf, 1, nat(BUF_IN, 0, 0, Np1 * Nf, (int)N0), o);
      sx.wait_credits = store;
      if (masked) {
        sx.mask.on = 1;
        sx.mask.jdiv = (int)Nf;
        band(N0, false, sx.mask.i_lo, sx.mask.i_hi);
        band(N1, false, sx.mask.jq_lo, sx.mask.jq_hi);
        sx.mask.jq_off = (int)(me * Np1);
        band(N2, !c2c, sx.mask.jr_lo, sx.mask.jr_hi);
      }
      sx.rec_ev = pg.nevents++;
      const int x_ev = sx.rec_ev;
      if (!store) b.use(BUF_W0, P * blk);  // send buffer
      b.use(BUF_W1, P * blk);
      // the y pass of chunk c writes W0 rows that the sends of later chunks still read when the
      // padded planes are larger than the send blocks: give its output a buffer of its own then
      const int ybuf = BUF_W2;
      b.use(ybuf, (long long)pNp0 * pN1 * Nf);
      // (fused transport: the x pass has moved everything, one flag step covers all planes)
      const int C = store ? 1 : pick_chunks(d.chunks, pNp0, blk * csz, p2p);
      const long long xc = pNp0 / C;
      std::vector<int> xev((size_t)C);
      for (int c = 0; c < C; ++c) {  // all exchanges are queued first: they only depend on the x pass
        const long long x0 = c * xc;
        b.fixed = 1;
        Step& x = b.exch(0, P, me);
        x.stream = 1;
        x.wait_ev = (c == 0) ? x_ev : -1;
        x.rec_ev = pg.nevents++;
        xev[(size_t)c] = x.rec_ev;
        x.first_exch = (c == 0);
        x.fused = store;
        for (int q = 0; q < P; ++q) {
          x.send[q].buf = BUF_W0; x.send[q].off = q * blk + x0 * Np1 * Nf; x.scnt[q] = xc * Np1 * Nf;
          x.recv[q].buf = BUF_W1; x.recv[q].off = q * blk + x0 * Np1 * Nf; x.rcnt[q] = xc * Np1 * Nf;
          x.rpeer[q].buf = BUF_W1; x.rpeer[q].off = me * blk + x0 * Np1 * Nf;
        }
      }
      for (int c = 0; c < C; ++c) {
        const long long x0 = c * xc;
        zy_groups(x0, xc, [&](long long g0, long long gn) {
          SideT g;
          g.chunk = (int)Np1;
          g.nchunk = P;
          g.nphys = (int)N1;
          for (int q = 0; q < P; ++q) {
            g.base[q].buf = BUF_W1;
            g.base[q].off = q * blk + g0 * Np1 * Nf;
            g.sb[q] = Np1 * Nf;
            g.si[q] = Nf;
          }
          b.fixed = 2;
          Step& y = b.strided(pN1, gn, Nf, 1, g, nat(ybuf, g0 * pN1 * Nf, pN1 * Nf, Nf, pN1));
          if (g0 == x0) y.wait_ev = xev[(size_t)c];
          b.fixed = 3;
          Step& z = zinv(gn * pN1, g0 * pN1, ybuf, g0 * pN1 * Nf, scale);
          // credits go back after the last z pass, not the last y pass: W2 (this program's y output)
          // is the forward program's receive buffer, which the peers fill as soon as they hold credits
          z.last_reader = (c == C - 1) && (g0 + gn == x0 + xc);
        });
      }
    }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// pencil.R2CX / R2CY programs (pencil.py:386-883, 1001-1477)
// ------------------------------------------------------------------------------------------------
inline int build_pencil(const b200fft_plan_desc_t& d, int inverse, int dealias, Program& pg) {
  Builder b(pg);
  const bool padded = dealias == B200FFT_DEALIAS_3_2;
  const bool masked = inverse && dealias == B200FFT_DEALIAS_2_3;
  const double p = padded ? d.padsize : 1.0;
  const int P = d.nranks, me = d.rank;
  const long long N0 = d.N[0], N1 = d.N[1], N2 = d.N[2];
  const bool peer_mapped = (d.transport == B200FFT_TRANSPORT_P2P || d.transport == B200FFT_TRANSPORT_STORE) && P > 1;
  const bool alignX = d.kind == B200FFT_PENCIL_X;
  const int P1 = d.P1, P2 = d.P2;
  const int c0 = me % P1, c1 = me / P1;
  const long long Nf = N2 / 2 + 1;
  const long long a = N0 / P1, bq = N1 / P2;      // real block: (a, bq, N2)
  const int pa = ipad(p, a), pbq = ipad(p, bq);
  const int pN0 = ipad(p, N0), pN1 = ipad(p, N1), pN2 = ipad(p, N2);
  const int zparts = alignX ? P2 : P1, zme = alignX ? c1 : c0;
  const long long C = Nf / zparts;
  long long zc[PMAXP], zoff[PMAXP + 1];
  zoff[0] = 0;
  for (int q = 0; q < zparts; ++q) {
    zc[q] = C + ((q == zparts - 1 && !d.drop_nyquist) ? Nf % zparts : 0);
    zoff[q + 1] = zoff[q] + zc[q];
  }
  const long long kzl = zc[zme];
  const int nk = (int)zoff[zparts];  // Nf, or Nf-1 for the AlltoallN layout
  const double p3 = p * p * p;
  const long long rowsz = (long long)pa * pbq;  // z rows per rank
  if (padded && ((long long)pa * P1 != pN0 || (long long)pbq * P2 != pN1 || (long long)ipad(p, N1 / P1) * P1 != pN1 ||
                 (long long)ipad(p, N0 / P2) * P2 != pN0))  // the padded blocks must tile the padded mesh (pencil.py:273-275)
    return fail(B200FFT_ERR_ARG, "3/2-rule: padsize * N / P1 and padsize * N / P2 must be integers");
  
  const double iscale = (padded ? p3 : 1.0) / ((double)pN0 * (double)pN1 * (double)pN2);

  // z pass stores / loads: kz cut into zparts chunks, one per peer of the z communicator
  auto zside = [&](int selfbuf, long long selfoff, int otherbuf, bool recv_layout) {
    SideT s;
    s.chunk = (int)C;
    s.nchunk = zparts;
    s.nphys = nk;
    for (int q = 0; q < zparts; ++q) {
      if (recv_layout) {  // blocks from q: [rowsz][zc[q]] at rowsz*zoff[q]
        s.base[q].buf = otherbuf;
        s.base[q].off = rowsz * zoff[q];
      } else {
        s.base[q].buf = (q == zme) ? selfbuf : otherbuf;
        s.base[q].off = (q == zme) ? selfoff : rowsz * zoff[q];
      }
      s.sb[q] = zc[q];
      s.si[q] = 1;
    }
    return s;
  };

  if (alignX) {
    const long long y1 = N1 / P1;
    const long long blk2 = (long long)pa * y1 * kzl;     // comm0 exchange block
    const long long blk1 = rowsz * kzl;                  // comm1 block [pa][pbq][kzl]
    // Pipelined programs (d.chunks > 1; NCCL and copy-engine transports): the local x planes are cut into
    // CH chunks and both exchanges run per chunk on the second stream -- forward z(c) | e1(c) | y(c) |
    // e2(c) then one x pass, inverse one x pass then e2(c) | y(c) | e1(c) | z(c) -- so each exchange
    // overlaps the FFT passes of the neighbouring chunks.  The send buffer of the second exchange of a
    // direction must not alias the first one's any more (W3 where W2 is taken).
    int CH = grid_chunks(d, N0 * N1 * (N2 / 2 + 1) / P);
    while (CH > 1 && pa % CH) --CH;
    const long long xc = pa / CH;
    // kz-chunked side of the z pass restricted to local planes [x0, x0 + xc)
    auto zside_x = [&](SideT sd, long long x0) {
      for (int q = 0; q < zparts; ++q) sd.base[q].off += x0 * pbq * zc[q];
      return sd;
    };
    if (!inverse && CH > 1) {  // pencil.py:1312-1337 (+ padded :1440-1475), pipelined
      const int recv2 = (padded || peer_mapped) ? BUF_W2 : BUF_OUT;
      const int send2 = (recv2 == BUF_OUT) ? BUF_W2 : BUF_W3;
      b.use(BUF_W0, rowsz * nk);
      b.use(BUF_W1, P2 * blk1);
      b.use(send2, P1 * blk2);
      b.use(recv2, P1 * blk2);
      std::vector<int> e1ev((size_t)CH);
      for (int c = 0; c < CH; ++c) {
        const long long x0 = c * xc;
        b.fixed = 0;
        Step& z = b.rows(true, xc * pbq, pN2, nk, BUF_IN, zside_x(zside(BUF_W1, zme * blk1, BUF_W0, false), x0));
        z.real.off = x0 * pbq * pN2;
        z.rec_ev = pg.nevents++;
        const int zev = z.rec_ev;
        b.fixed = 1;
        Step& x1 = b.exch(2, P2, c1);
        x1.stream = 1;
        x1.wait_ev = zev;
        e1ev[(size_t)c] = x1.rec_ev = pg.nevents++;
        for (int q = 0; q < P2; ++q) {
          x1.send[q].buf = BUF_W0; x1.send[q].off = rowsz * zoff[q] + x0 * pbq * zc[q]; x1.scnt[q] = xc * pbq * zc[q];
          x1.recv[q].buf = BUF_W1; x1.recv[q].off = q * blk1 + x0 * pbq * kzl; x1.rcnt[q] = xc * pbq * kzl;
          x1.rpeer[q].buf = BUF_W1; x1.rpeer[q].off = c1 * rowsz * zc[q] + x0 * pbq * zc[q];
        }
      }
      int last_ev = -1;
      for (int c = 0; c < CH; ++c) {
        const long long x0 = c * xc;
        SideT g;
        g.chunk = pbq; g.nchunk = P2; g.nphys = pN1;
        for (int q = 0; q < P2; ++q) { g.base[q].buf = BUF_W1; g.base[q].off = q * blk1 + x0 * pbq * kzl; g.sb[q] = pbq * kzl; g.si[q] = kzl; }
        SideT o;
        o.chunk = (int)y1; o.nchunk = P1; o.nphys = (int)N1;
        for (int q = 0; q < P1; ++q) {
          o.base[q].buf = (q == c0) ? recv2 : send2; o.base[q].off = q * blk2 + x0 * y1 * kzl; o.sb[q] = y1 * kzl; o.si[q] = kzl;
        }
        b.fixed = 2;
        Step& y = b.strided(pN1, xc, kzl, 0, g, o, padded ? 1 : 0);
        y.wait_ev = e1ev[(size_t)c];
        y.rec_ev = pg.nevents++;
        const int yev = y.rec_ev;
        b.fixed = 3;
        Step& x2 = b.exch(1, P1, c0);
        x2.stream = 1;
        x2.wait_ev = yev;
        last_ev = x2.rec_ev = pg.nevents++;
        for (int q = 0; q < P1; ++q) {
          x2.send[q].buf = send2; x2.send[q].off = q * blk2 + x0 * y1 * kzl; x2.scnt[q] = xc * y1 * kzl;
          x2.recv[q].buf = recv2; x2.recv[q].off = q * blk2 + x0 * y1 * kzl; x2.rcnt[q] = xc * y1 * kzl;
          x2.rpeer[q].buf = recv2; x2.rpeer[q].off = c0 * blk2 + x0 * y1 * kzl;
        }
      }
      b.fixed = 4;
      Step& fx = b.strided(pN0, 1, y1 * kzl, 0, nat(recv2, 0, 0, y1 * kzl, pN0), nat(BUF_OUT, 0, 0, y1 * kzl, (int)N0),
                           padded ? 1 : 0, padded ? 1.0 / p3 : 1.0);
      fx.wait_ev = last_ev;  // exchanges complete in order on the second stream
    } else if (inverse && CH > 1) {  // pencil.py:1082-1105 (+ padded :1196-1223), pipelined
      SideT o;
      o.chunk = pa; o.nchunk = P1; o.nphys = pN0;
      for (int q = 0; q < P1; ++q) {
        o.base[q].buf = (q == c0) ? BUF_W1 : BUF_W0; o.base[q].off = q * blk2; o.sb[q] = 0; o.si[q] = y1 * kzl;
      }
      b.fixed = 0;
      Step& sx = b.strided(pN0, 1, y1 * kzl, 1, nat(BUF_IN, 0, 0, y1 * kzl, (int)N0), o);
      if (masked) {
        sx.mask.on = 1;
        sx.mask.jdiv = (int)kzl;
        band(N0, false, sx.mask.i_lo, sx.mask.i_hi);
        band(N1, false, sx.mask.jq_lo, sx.mask.jq_hi);
        sx.mask.jq_off = (int)(c0 * y1);
        band(N2, true, sx.mask.jr_lo, sx.mask.jr_hi);
        sx.mask.jr_off = (int)(c1 * C);
      }
      const int xev = sx.rec_ev = pg.nevents++;
      b.use(BUF_W0, P1 * blk2);
      b.use(BUF_W1, P1 * blk2);
      b.use(BUF_W3, P2 * blk1);      // send buffer of the second exchange
      b.use(BUF_W2, rowsz * nk);
      std::vector<int> e2ev((size_t)CH), e1ev((size_t)CH);
      for (int c = 0; c < CH; ++c) {  // all first exchanges are queued at once: they only depend on the x pass
        const long long x0 = c * xc;
        b.fixed = 1;
        Step& x2 = b.exch(1, P1, c0);
        x2.stream = 1;
        x2.wait_ev = (c == 0) ? xev : -1;
        e2ev[(size_t)c] = x2.rec_ev = pg.nevents++;
        for (int q = 0; q < P1; ++q) {
          x2.send[q].buf = BUF_W0; x2.send[q].off = q * blk2 + x0 * y1 * kzl; x2.scnt[q] = xc * y1 * kzl;
          x2.recv[q].buf = BUF_W1; x2.recv[q].off = q * blk2 + x0 * y1 * kzl; x2.rcnt[q] = xc * y1 * kzl;
          x2.rpeer[q].buf = BUF_W1; x2.rpeer[q].off = c0 * blk2 + x0 * y1 * kzl;
        }
      }
      for (int c = 0; c < CH; ++c) {
        const long long x0 = c * xc;
        SideT g;
        g.chunk = (int)y1; g.nchunk = P1; g.nphys = (int)N1;
        for (int q = 0; q < P1; ++q) { g.base[q].buf = BUF_W1; g.base[q].off = q * blk2 + x0 * y1 * kzl; g.sb[q] = y1 * kzl; g.si[q] = kzl; }
        SideT o2;
        o2.chunk = pbq; o2.nchunk = P2; o2.nphys = pN1;
        for (int q = 0; q < P2; ++q) {
          o2.base[q].buf = (q == c1) ? BUF_W2 : BUF_W3;
          o2.base[q].off = ((q == c1) ? rowsz * zoff[c1] : q * blk1) + x0 * pbq * kzl;
          o2.sb[q] = pbq * kzl; o2.si[q] = kzl;
        }
        b.fixed = 2;
        Step& y = b.strided(pN1, xc, kzl, 1, g, o2);
        y.wait_ev = e2ev[(size_t)c];
        const int yev = y.rec_ev = pg.nevents++;
        b.fixed = 3;
        Step& x1 = b.exch(2, P2, c1);
        x1.stream = 1;
        x1.wait_ev = yev;
        e1ev[(size_t)c] = x1.rec_ev = pg.nevents++;
        for (int q = 0; q < P2; ++q) {
          x1.send[q].buf = BUF_W3; x1.send[q].off = q * blk1 + x0 * pbq * kzl; x1.scnt[q] = xc * pbq * kzl;
          x1.recv[q].buf = BUF_W2; x1.recv[q].off = rowsz * zoff[q] + x0 * pbq * zc[q]; x1.rcnt[q] = xc * pbq * zc[q];
          x1.rpeer[q].buf = BUF_W2; x1.rpeer[q].off = rowsz * zoff[c1] + x0 * pbq * kzl;
        }
      }
      for (int c = 0; c < CH; ++c) {
        const long long x0 = c * xc;
        b.fixed = 4;
        Step& z = b.rows(false, xc * pbq, pN2, nk, BUF_OUT, zside_x(zside(0, 0, BUF_W2, true), x0), iscale);
        z.real.off = x0 * pbq * pN2;
        z.wait_ev = e1ev[(size_t)c];
      }
    } else if (!inverse) {  // pencil.py:1312-1337 (+ padded :1440-1475)
      const int recv2 = (padded || peer_mapped) ? BUF_W2 : BUF_OUT;  // peers write plan-owned buffers only
      b.rows(true, rowsz, pN2, nk, BUF_IN, zside(BUF_W1, zme * blk1, BUF_W0, false));
      b.use(BUF_W0, rowsz * nk);
      b.use(BUF_W1, P2 * blk1);
      Step& x1 = b.exch(2, P2, c1);
      for (int q = 0; q < P2; ++q) {
        x1.send[q].buf = BUF_W0; x1.send[q].off = rowsz * zoff[q]; x1.scnt[q] = rowsz * zc[q];
        x1.recv[q].buf = BUF_W1; x1.recv[q].off = q * blk1; x1.rcnt[q] = blk1;
        x1.rpeer[q].buf = BUF_W1; x1.rpeer[q].off = c1 * rowsz * zc[q];
      }
      SideT g;  // gather y from the P2 peers
      g.chunk = pbq; g.nchunk = P2; g.nphys = pN1;
      for (int q = 0; q < P2; ++q) { g.base[q].buf = BUF_W1; g.base[q].off = q * blk1; g.sb[q] = pbq * kzl; g.si[q] = kzl; }
      SideT o;  // split y over the P1 peers
      o.chunk = (int)y1; o.nchunk = P1; o.nphys = (int)N1;
      for (int q = 0; q < P1; ++q) {
        o.base[q].buf = (q == c0) ? recv2 : BUF_W0; o.base[q].off = q * blk2; o.sb[q] = y1 * kzl; o.si[q] = kzl;
      }
      b.strided(pN1, pa, kzl, 0, g, o, padded ? 1 : 0);
      b.use(BUF_W0, P1 * blk2);
      b.use(recv2, P1 * blk2);
      Step& x2 = b.exch(1, P1, c0);
      for (int q = 0; q < P1; ++q) {
        x2.send[q].buf = BUF_W0; x2.send[q].off = q * blk2; x2.scnt[q] = blk2;
        x2.recv[q].buf = recv2; x2.recv[q].off = q * blk2; x2.rcnt[q] = blk2;
        x2.rpeer[q].buf = recv2; x2.rpeer[q].off = c0 * blk2;
      }
      b.strided(pN0, 1, y1 * kzl, 0, nat(recv2, 0, 0, y1 * kzl, pN0), nat(BUF_OUT, 0, 0, y1 * kzl, (int)N0),
                padded ? 1 : 0, padded ? 1.0 / p3 : 1.0);
    } else {  // pencil.py:1082-1105 (+ padded :1196-1223)
      SideT o;
      o.chunk = pa; o.nchunk = P1; o.nphys = pN0;
      for (int q = 0; q < P1; ++q) {
        o.base[q].buf = (q == c0) ? BUF_W1 : BUF_W0; o.base[q].off = q * blk2; o.sb[q] = 0; o.si[q] = y1 * kzl;
      }
      Step& sx = b.strided(pN0, 1, y1 * kzl, 1, nat(BUF_IN, 0, 0, y1 * kzl, (int)N0), o);
      if (masked) {
        sx.mask.on = 1;
        sx.mask.jdiv = (int)kzl;
        band(N0, false, sx.mask.i_lo, sx.mask.i_hi);
        band(N1, false, sx.mask.jq_lo, sx.mask.jq_hi);
        sx.mask.jq_off = (int)(c0 * y1);
        band(N2, true, sx.mask.jr_lo, sx.mask.jr_hi);
        sx.mask.jr_off = (int)(c1 * C);
      }
      b.use(BUF_W0, P1 * blk2);
      b.use(BUF_W1, P1 * blk2);
      Step& x2 = b.exch(1, P1, c0);
      for (int q = 0; q < P1; ++q) {
        x2.send[q].buf = BUF_W0; x2.send[q].off = q * blk2; x2.scnt[q] = blk2;
        x2.recv[q].buf = BUF_W1; x2.recv[q].off = q * blk2; x2.rcnt[q] = blk2;
        x2.rpeer[q].buf = BUF_W1; x2.rpeer[q].off = c0 * blk2;
      }
      SideT g;
      g.chunk = (int)y1; g.nchunk = P1; g.nphys = (int)N1;
      for (int q = 0; q < P1; ++q) { g.base[q].buf = BUF_W1; g.base[q].off = q * blk2; g.sb[q] = y1 * kzl; g.si[q] = kzl; }
      SideT o2;
      o2.chunk = pbq; o2.nchunk = P2; o2.nphys = pN1;
      for (int q = 0; q < P2; ++q) {
        o2.base[q].buf = (q == c1) ? BUF_W2 : BUF_W0;
        o2.base[q].off = (q == c1) ? rowsz * zoff[c1] : q * blk1;
        o2.sb[q] = pbq * kzl; o2.si[q] = kzl;
      }
      b.strided(pN1, pa, kzl, 1, g, o2);
      b.use(BUF_W0, P2 * blk1);
      b.use(BUF_W2, rowsz * nk);
      Step& x1 = b.exch(2, P2, c1);
      for (int q = 0; q < P2; ++q) {
        x1.send[q].buf = BUF_W0; x1.send[q].off = q * blk1; x1.scnt[q] = blk1;
        x1.recv[q].buf = BUF_W2; x1.recv[q].off = rowsz * zoff[q]; x1.rcnt[q] = rowsz * zc[q];
        x1.rpeer[q].buf = BUF_W2; x1.rpeer[q].off = rowsz * zoff[c1];
      }
      b.rows(false, rowsz, pN2, nk, BUF_OUT, zside(0, 0, BUF_W2, true), iscale);
    }
  } else {  // alignment Y
    const long long x2l = N0 / P2;                        // final local x extent
    const long long blk = x2l * pbq * kzl;                // comm1 exchange block
    const long long blk1 = rowsz * kzl;                   // comm0 block [pa][pbq][kzl]
    // Pipelined programs (d.chunks > 1; NCCL and copy-engine transports).  No local axis survives both
    // exchanges here except kz, so every rank's kz range is cut into CH sub-ranges and all exchange
    // buffers become sub-range-major -- [c][peer][...][kz in c] -- which keeps every (sub-range, peer)
    // message contiguous.  forward: z | ea(c) | x(c) | eb(c) | y(c);  inverse: y(c) | eb(c) | x(c) | ea(c) | z:
    // a three-stage pipeline on both sides of the x pass.  The z pass addresses the P1*CH kz chunks of
    // its complex side directly (hence P1*CH <= 16).
    int CH = grid_chunks(d, N0 * N1 * (N2 / 2 + 1) / P);
    while (CH > 1 && (C % CH || (long long)P1 * CH > PMAXP)) --CH;
    const long long kzc = C / CH;
    auto kq = [&](int c, int q) { return kzc + ((c == CH - 1) ? zc[q] - C : 0); };   // width of sub-range c on rank q
    auto zoffc = [&](int c, int q) { return rowsz * ((long long)c * P1 * kzc + (long long)q * kzc); };  // [c][q][rowsz][kq]
    if (CH > 1) {
      std::vector<long long> k0((size_t)CH), kcc((size_t)CH);
      for (int c = 0; c < CH; ++c) {
        k0[(size_t)c] = c * kzc;
        kcc[(size_t)c] = kq(c, c0);
      }
      auto roff = [&](int c, int q) { return (long long)P1 * rowsz * k0[(size_t)c] + (long long)q * rowsz * kcc[(size_t)c]; };
      auto off2 = [&](int c, int q) { return (long long)P2 * x2l * pbq * k0[(size_t)c] + (long long)q * x2l * pbq * kcc[(size_t)c]; };
      // complex side of the z pass: chunk (q, c) of kzc entries (the last one takes the Nyquist entry)
      auto zside_c = [&](bool recv_layout, int otherbuf) {
        SideT sd;
        sd.chunk = (int)kzc;
        sd.nchunk = P1 * CH;
        sd.nphys = nk;
        for (int q = 0; q < P1; ++q)
          for (int c = 0; c < CH; ++c) {
            const int pidx = q * CH + c;
            if (recv_layout) {
              sd.base[pidx].buf = BUF_W2; sd.base[pidx].off = zoffc(c, q);
            } else {
              sd.base[pidx].buf = (q == c0) ? BUF_W1 : otherbuf;
              sd.base[pidx].off = (q == c0) ? roff(c, c0) : zoffc(c, q);
            }
            sd.sb[pidx] = kq(c, q);
            sd.si[pidx] = 1;
          }
        return sd;
      };
      if (!inverse) {  // pencil.py:730-754 (+ padded :853-881), pipelined
        b.use(BUF_W0, rowsz * nk);
        b.use(BUF_W1, P1 * blk1);
        b.use(BUF_W3, P2 * blk);
        b.use(BUF_W2, P2 * blk);
        b.fixed = 0;
        Step& z = b.rows(true, rowsz, pN2, nk, BUF_IN, zside_c(false, BUF_W0));
        const int zev = z.rec_ev = pg.nevents++;
        std::vector<int> eaev((size_t)CH), ebev((size_t)CH);
        for (int c = 0; c < CH; ++c) {  // all first exchanges only depend on the z pass
          b.fixed = 1;
          Step& xa = b.exch(1, P1, c0);
          xa.stream = 1;
          xa.wait_ev = (c == 0) ? zev : -1;
          eaev[(size_t)c] = xa.rec_ev = pg.nevents++;
          for (int q = 0; q < P1; ++q) {
            xa.send[q].buf = BUF_W0; xa.send[q].off = zoffc(c, q); xa.scnt[q] = rowsz * kq(c, q);
            xa.recv[q].buf = BUF_W1; xa.recv[q].off = roff(c, q); xa.rcnt[q] = rowsz * kcc[(size_t)c];
            xa.rpeer[q].buf = BUF_W1; xa.rpeer[q].off = (long long)P1 * rowsz * k0[(size_t)c] + (long long)c0 * rowsz * kq(c, q);
          }
        }
        for (int c = 0; c < CH; ++c) {
          const long long kc_ = kcc[(size_t)c];
          SideT g;
          g.chunk = pa; g.nchunk = P1; g.nphys = pN0;
          for (int q = 0; q < P1; ++q) { g.base[q].buf = BUF_W1; g.base[q].off = roff(c, q); g.sb[q] = 0; g.si[q] = pbq * kc_; }
          SideT o;
          o.chunk = (int)x2l; o.nchunk = P2; o.nphys = (int)N0;
          for (int q = 0; q < P2; ++q) {
            o.base[q].buf = (q == c1) ? BUF_W2 : BUF_W3; o.base[q].off = off2(c, q); o.sb[q] = 0; o.si[q] = pbq * kc_;
          }
          b.fixed = 2;
          Step& fx = b.strided(pN0, 1, pbq * kc_, 0, g, o, padded ? 1 : 0);
          fx.wait_ev = eaev[(size_t)c];
          const int xev = fx.rec_ev = pg.nevents++;
          b.fixed = 3;
          Step& xb = b.exch(2, P2, c1);
          xb.stream = 1;
          xb.wait_ev = xev;
          ebev[(size_t)c] = xb.rec_ev = pg.nevents++;
          for (int q = 0; q < P2; ++q) {
            xb.send[q].buf = BUF_W3; xb.send[q].off = off2(c, q); xb.scnt[q] = x2l * pbq * kc_;
            xb.recv[q].buf = BUF_W2; xb.recv[q].off = off2(c, q); xb.rcnt[q] = x2l * pbq * kc_;
            xb.rpeer[q].buf = BUF_W2; xb.rpeer[q].off = off2(c, c1);
          }
        }
        for (int c = 0; c < CH; ++c) {
          const long long kc_ = kcc[(size_t)c];
          SideT g;
          g.chunk = pbq; g.nchunk = P2; g.nphys = pN1;
          for (int q = 0; q < P2; ++q) { g.base[q].buf = BUF_W2; g.base[q].off = off2(c, q); g.sb[q] = pbq * kc_; g.si[q] = kc_; }
          b.fixed = 4;
          Step& fy = b.strided(pN1, x2l, kc_, 0, g, nat(BUF_OUT, k0[(size_t)c], N1 * kzl, kzl, (int)N1), padded ? 1 : 0,
                               padded ? 1.0 / p3 : 1.0);
          fy.wait_ev = ebev[(size_t)c];
        }
      } else {  // pencil.py:483-507 (+ padded :597-629), pipelined
        b.use(BUF_W0, P2 * blk);
        b.use(BUF_W1, P2 * blk);
        b.use(BUF_W3, P1 * blk1);
        b.use(BUF_W2, rowsz * nk);
        std::vector<int> ebev((size_t)CH);
        int last_ev = -1;
        for (int c = 0; c < CH; ++c) {
          const long long kc_ = kcc[(size_t)c];
          SideT o;
          o.chunk = pbq; o.nchunk = P2; o.nphys = pN1;
          for (int q = 0; q < P2; ++q) {
            o.base[q].buf = (q == c1) ? BUF_W1 : BUF_W0; o.base[q].off = off2(c, q); o.sb[q] = pbq * kc_; o.si[q] = kc_;
          }
          b.fixed = 0;
          Step& sy = b.strided(pN1, x2l, kc_, 1, nat(BUF_IN, k0[(size_t)c], N1 * kzl, kzl, (int)N1), o);
          if (masked) {
            sy.mask.on = 1;
            sy.mask.jdiv = 0x3fffffff;
            band(N0, false, sy.mask.b_lo, sy.mask.b_hi);
            sy.mask.b_off = (int)(c1 * x2l);
            band(N1, false, sy.mask.i_lo, sy.mask.i_hi);
            band(N2, true, sy.mask.jr_lo, sy.mask.jr_hi);
            sy.mask.jr_off = (int)(c0 * C + k0[(size_t)c]);
          }
          const int yev = sy.rec_ev = pg.nevents++;
          b.fixed = 1;
          Step& xb = b.exch(2, P2, c1);
          xb.stream = 1;
          xb.wait_ev = yev;
          ebev[(size_t)c] = xb.rec_ev = pg.nevents++;
          for (int q = 0; q < P2; ++q) {
            xb.send[q].buf = BUF_W0; xb.send[q].off = off2(c, q); xb.scnt[q] = x2l * pbq * kc_;
            xb.recv[q].buf = BUF_W1; xb.recv[q].off = off2(c, q); xb.rcnt[q] = x2l * pbq * kc_;
            xb.rpeer[q].buf = BUF_W1; xb.rpeer[q].off = off2(c, c1);
          }
        }
        for (int c = 0; c < CH; ++c) {
          const long long kc_ = kcc[(size_t)c];
          SideT g;
          g.chunk = (int)x2l; g.nchunk = P2; g.nphys = (int)N0;
          for (int q = 0; q < P2; ++q) { g.base[q].buf = BUF_W1; g.base[q].off = off2(c, q); g.sb[q] = 0; g.si[q] = pbq * kc_; }
          SideT o2;
          o2.chunk = pa; o2.nchunk = P1; o2.nphys = pN0;
          for (int q = 0; q < P1; ++q) {
            o2.base[q].buf = (q == c0) ? BUF_W2 : BUF_W3;
            o2.base[q].off = (q == c0) ? zoffc(c, c0) : roff(c, q);
            o2.sb[q] = 0; o2.si[q] = pbq * kc_;
          }
          b.fixed = 2;
          Step& sx = b.strided(pN0, 1, pbq * kc_, 1, g, o2);
          sx.wait_ev = ebev[(size_t)c];
          const int xev = sx.rec_ev = pg.nevents++;
          b.fixed = 3;
          Step& xa = b.exch(1, P1, c0);
          xa.stream = 1;
          xa.wait_ev = xev;
          last_ev = xa.rec_ev = pg.nevents++;
          for (int q = 0; q < P1; ++q) {
            xa.send[q].buf = BUF_W3; xa.send[q].off = roff(c, q); xa.scnt[q] = rowsz * kc_;
            xa.recv[q].buf = BUF_W2; xa.recv[q].off = zoffc(c, q); xa.rcnt[q] = rowsz * kq(c, q);
            xa.rpeer[q].buf = BUF_W2; xa.rpeer[q].off = zoffc(c, c0);
          }
        }
        b.fixed = 4;
        Step& z = b.rows(false, rowsz, pN2, nk, BUF_OUT, zside_c(true, BUF_W2), iscale);
        z.wait_ev = last_ev;  // exchanges complete in order on the second stream
      }
    } else if (!inverse) {  // pencil.py:730-754 (+ padded :853-881)
      b.rows(true, rowsz, pN2, nk, BUF_IN, zside(BUF_W1, zme * blk1, BUF_W0, false));
      b.use(BUF_W0, rowsz * nk);
      b.use(BUF_W1, P1 * blk1);
      Step& xa = b.exch(1, P1, c0);
      for (int q = 0; q < P1; ++q) {
        xa.send[q].buf = BUF_W0; xa.send[q].off = rowsz * zoff[q]; xa.scnt[q] = rowsz * zc[q];
        xa.recv[q].buf = BUF_W1; xa.recv[q].off = q * blk1; xa.rcnt[q] = blk1;
        xa.rpeer[q].buf = BUF_W1; xa.rpeer[q].off = c0 * rowsz * zc[q];
      }
      SideT o;
      o.chunk = (int)x2l; o.nchunk = P2; o.nphys = (int)N0;
      for (int q = 0; q < P2; ++q) {
        o.base[q].buf = (q == c1) ? BUF_W2 : BUF_W0; o.base[q].off = q * blk; o.sb[q] = 0; o.si[q] = pbq * kzl;
      }
      b.strided(pN0, 1, pbq * kzl, 0, nat(BUF_W1, 0, 0, pbq * kzl, pN0), o, padded ? 1 : 0);
      b.use(BUF_W0, P2 * blk);
      b.use(BUF_W2, P2 * blk);
      Step& xb = b.exch(2, P2, c1);
      for (int q = 0; q < P2; ++q) {
        xb.send[q].buf = BUF_W0; xb.send[q].off = q * blk; xb.scnt[q] = blk;
        xb.recv[q].buf = BUF_W2; xb.recv[q].off = q * blk; xb.rcnt[q] = blk;
        xb.rpeer[q].buf = BUF_W2; xb.rpeer[q].off = c1 * blk;
      }
      SideT g;
      g.chunk = pbq; g.nchunk = P2; g.nphys = pN1;
      for (int q = 0; q < P2; ++q) { g.base[q].buf = BUF_W2; g.base[q].off = q * blk; g.sb[q] = pbq * kzl; g.si[q] = kzl; }
      b.strided(pN1, x2l, kzl, 0, g, nat(BUF_OUT, 0, N1 * kzl, kzl, (int)N1), padded ? 1 : 0, padded ? 1.0 / p3 : 1.0);
    } else {  // pencil.py:483-507 (+ padded :597-629)
      SideT o;
      o.chunk = pbq; o.nchunk = P2; o.nphys = pN1;
      for (int q = 0; q < P2; ++q) {
        o.base[q].buf = (q == c1) ? BUF_W1 : BUF_W0; o.base[q].off = q * blk; o.sb[q] = pbq * kzl; o.si[q] = kzl;
      }
      Step& sy = b.strided(pN1, x2l, kzl, 1, nat(BUF_IN, 0, N1 * kzl, kzl, (int)N1), o);
      if (masked) {
        sy.mask.on = 1;
        sy.mask.jdiv = 0x3fffffff;
        band(N0, false, sy.mask.b_lo, sy.mask.b_hi);
        sy.mask.b_off = (int)(c1 * x2l);
        band(N1, false, sy.mask.i_lo, sy.mask.i_hi);
        band(N2, true, sy.mask.jr_lo, sy.mask.jr_hi);
        sy.mask.jr_off = (int)(c0 * C);
      }
      b.use(BUF_W0, P2 * blk);
      b.use(BUF_W1, P2 * blk);
      Step& xb = b.exch(2, P2, c1);
      for (int q = 0; q < P2; ++q) {
        xb.send[q].buf = BUF_W0; xb.send[q].off = q * blk; xb.scnt[q] = blk;
        xb.recv[q].buf = BUF_W1; xb.recv[q].off = q * blk; xb.rcnt[q] = blk;
        xb.rpeer[q].buf = BUF_W1; xb.rpeer[q].off = c1 * blk;
      }
      SideT o2;
      o2.chunk = pa; o2.nchunk = P1; o2.nphys = pN0;
      for (int q = 0; q < P1; ++q) {
        o2.base[q].buf = (q == c0) ? BUF_W2 : BUF_W0;
        o2.base[q].off = (q == c0) ? rowsz * zoff[c0] : q * blk1;
        o2.sb[q] = 0; o2.si[q] = pbq * kzl;
      }
      b.strided(pN0, 1, pbq * kzl, 1, nat(BUF_W1, 0, 0, pbq * kzl, (int)N0), o2);
      b.use(BUF_W0, P1 * blk1);
      b.use(BUF_W2, rowsz * nk);
      Step& xa = b.exch(1, P1, c0);
      for (int q = 0; q < P1; ++q) {
        xa.send[q].buf = BUF_W0; xa.send[q].off = q * blk1; xa.scnt[q] = blk1;
        xa.recv[q].buf = BUF_W2; xa.recv[q].off = rowsz * zoff[q]; xa.rcnt[q] = rowsz * zc[q];
        xa.rpeer[q].buf = BUF_W2; xa.rpeer[q].off = rowsz * zoff[c0];
      }
      b.rows(false, rowsz, pN2, nk, BUF_OUT, zside(0, 0, BUF_W2, true), iscale);
    }
  }
  return peer_mapped ? finish_peer_mapped(d, pg) : 0;
}

// ------------------------------------------------------------------------------------------------
// line.R2C programs (line.py:179-340)
// ------------------------------------------------------------------------------------------------
inline int build_line(const b200fft_plan_desc_t& d, int inverse, int dealias, Program& pg) {
  Builder b(pg);
  const bool padded = dealias == B200FFT_DEALIAS_3_2;
  const bool masked = inverse && dealias == B200FFT_DEALIAS_2_3;
  const double p = padded ? d.padsize : 1.0;
  const int P = d.nranks, me = d.rank;
  const long long N0 = d.N[0], N1 = d.N[1];
  const bool peer_mapped = (d.transport == B200FFT_TRANSPORT_P2P || d.transport == B200FFT_TRANSPORT_STORE) && P > 1;
  const long long Np0 = N0 / P, Np1 = N1 / P, Nf = N1 / 2 + 1;
  const int pN0 = ipad(p, N0), pNp0 = ipad(p, Np0), pN1 = ipad(p, N1);
  const long long kc = Np1 / 2;
  long long kcl[PMAXP], koff[PMAXP + 1];
  koff[0] = 0;
  for (int q = 0; q < P; ++q) {
    kcl[q] = kc + (q == P - 1 ? 1 : 0);
    koff[q + 1] = koff[q] + kcl[q];
  }
  const long long Npf = (P == 1) ? Nf : kcl[me];
  if (padded && (long long)pNp0 * P != pN0) return fail(B200FFT_ERR_ARG, "3/2-rule: padsize * N[0] / ranks must be an integer");
  // pipelined programs (d.chunks > 1; NCCL and copy-engine transports): the local rows are cut into CH
  // chunks whose exchange runs on the second stream beside the z pass of the next chunk
  int CH = P > 1 ? grid_chunks(d, N0 * (N1 / 2 + 1) / P) : 1;
  while (CH > 1 && pNp0 % CH) --CH;
  const long long rc = pNp0 / CH;
  const double p2 = p * p;
  const double iscale = (padded ? p2 : 1.0) / ((double)pN0 * (double)pN1);
  // The pass along x.  A column that needs 64 KB or more of shared memory (8192 points in single precision, 4096 in
  // double; BASELINE config 5a has 16384 single = 128 KB) leaves tiles one or two columns wide: 8- / 16-byte row
  // segments, a quarter / half of every 32-byte sector.  Such a transform runs as TWO launches instead
  // ("four-step", N = n1 * n2): A) n1-point transforms over rows n2 apart -- the n2 interleaved sub-columns side by
  // side as n2 * J columns, i.e. full-width tiles -- times the cross twiddles W_N^(x2 * k1) on store;
  // B) n2-point transforms over the n2 consecutive rows of each k1, output k2 scattered to row k1 + n1 * k2 (for
  // the inverse: to its peer's send block).  Twice the HBM traffic of one pass at full sector use instead of four
  // times.  One GPU, x pass of line 16384^2 single: 2.75 -> 0.99 ms (round trip 7.16 -> 3.61); 8192^2: 0.45 -> 0.26
  // (1.22 -> 0.83); at 32 KB columns it no longer pays (2048^2 double 0.095 -> 0.105 ms), profiles/r02_fourstep.
  // Plain and 2^k lengths only: padded / masked / 3 * 2^k transforms keep the single launch.
  const long long csz_line = d.precision == B200FFT_DOUBLE ? 16 : 8;
  auto xpass = [&](int inv, long long J, const SideT& in, const SideT& out, int fold, double scale, int tmpbuf,
                   Step** first, Step** last) {
    const long long n = pN0;
    const bool pow2 = n > 0 && (n & (n - 1)) == 0;
    bool split = !padded && !masked && pow2 && n * csz_line >= (64ll << 10) && in.nchunk == 1 && in.si[0] == J && fold == 0 &&
                 d.layout != B200FFT_LAYOUT_NATURAL;  // ("natural": the single launch, for A/B runs)
    long long n1 = 1;
    while (n1 * n1 < n) n1 *= 2;
    const long long n2 = n / n1;
    if (split && out.nchunk > 1 && out.chunk % n1) split = false;
    if (split && out.nchunk == 1 && out.si[0] != J) split = false;
    if (!split) {
      Step& s1 = b.strided((int)n, 1, J, inv, in, out, fold, scale);
      *first = *last = &s1;
      return;
    }
    const bool inplace = in.base[0].buf != BUF_IN && in.base[0].buf != BUF_OUT;
    const int abuf = inplace ? in.base[0].buf : tmpbuf;
    const long long aoff = inplace ? in.base[0].off : 0;
    if (!inplace) b.use(tmpbuf, n * J);
    const int pid = b.fixed;
    if (b.fixed < 0) b.fixed = b.next_pass++;  // both launches are one logical pass
    Step& sa = b.strided((int)n1, 1, n2 * J, inv, nat(in.base[0].buf, in.base[0].off, 0, n2 * J, (int)n1), nat(abuf, aoff, 0, n2 * J, (int)n1));
    sa.cross_n = (int)n;
    sa.cross_div = (int)J;
    SideT o2 = out;
    o2.nphys = (int)n2;
    if (out.nchunk == 1) {
      o2.chunk = (int)n2;
      o2.sb[0] = out.si[0];
      o2.si[0] = n1 * out.si[0];
    } else {
      o2.chunk = (int)(out.chunk / n1);
      for (int q = 0; q < out.nchunk; ++q) {
        o2.sb[q] = out.si[q];
        o2.si[q] = n1 * out.si[q];
      }
    }
    Step& sb2 = b.strided((int)n2, n1, J, inv, nat(abuf, aoff, n2 * J, J, (int)n2), o2, 0, scale);
    b.fixed = pid;
    *first = &sa;
    *last = &sb2;
  };
  Step *xs0 = nullptr, *xs1 = nullptr;
  if (!inverse) {
    if (P == 1) {  // line.py:182-191
      if (!padded) {
        b.rows(true, N0, (int)N1, (int)Nf, BUF_IN, nat(BUF_OUT, 0, Nf, 1, (int)Nf));
        xpass(0, Nf, nat(BUF_OUT, 0, 0, Nf, (int)N0), nat(BUF_OUT, 0, 0, Nf, (int)N0), 0, 1.0, BUF_W0, &xs0, &xs1);
      } else {
        b.rows(true, pN0, pN1, (int)Nf, BUF_IN, nat(BUF_W0, 0, Nf, 1, (int)Nf));
        b.use(BUF_W0, (long long)pN0 * Nf);
        b.strided(pN0, 1, Nf, 0, nat(BUF_W0, 0, 0, Nf, pN0), nat(BUF_OUT, 0, 0, Nf, (int)N0), 2, 1.0 / p2);
      }
    } else if (CH > 1) {  // line.py:193-258, pipelined: z(c) | exchange(c) over chunks of local rows, then the x pass
      const int recvbuf = (padded || peer_mapped) ? BUF_W1 : BUF_OUT;
      b.use(BUF_W0, (long long)pNp0 * Nf);
      b.use(recvbuf, (long long)pN0 * Npf);
      int last_ev = -1;
      for (int c = 0; c < CH; ++c) {
        const long long r0 = c * rc;
        SideT o;
        o.chunk = (int)kc; o.nchunk = P; o.nphys = (int)Nf;
        for (int q = 0; q < P; ++q) {
          o.base[q].buf = (q == me) ? recvbuf : BUF_W0;
          o.base[q].off = ((q == me) ? (long long)me * pNp0 * Npf : (long long)pNp0 * koff[q]) + r0 * kcl[q];
          o.sb[q] = kcl[q]; o.si[q] = 1;
        }
        b.fixed = 0;
        Step& z = b.rows(true, rc, pN1, (int)Nf, BUF_IN, o);
        z.real.off = r0 * pN1;
        const int zev = z.rec_ev = pg.nevents++;
        b.fixed = 1;
        Step& x = b.exch(0, P, me);
        x.stream = 1;
        x.wait_ev = zev;
        last_ev = x.rec_ev = pg.nevents++;
        for (int q = 0; q < P; ++q) {
          x.send[q].buf = BUF_W0; x.send[q].off = (long long)pNp0 * koff[q] + r0 * kcl[q]; x.scnt[q] = rc * kcl[q];
          x.recv[q].buf = recvbuf; x.recv[q].off = (long long)q * pNp0 * Npf + r0 * Npf; x.rcnt[q] = rc * Npf;
          x.rpeer[q].buf = recvbuf; x.rpeer[q].off = (long long)me * pNp0 * kcl[q] + r0 * kcl[q];
        }
      }
      b.fixed = 2;
      xpass(0, Npf, nat(recvbuf, 0, 0, Npf, pN0), nat(BUF_OUT, 0, 0, Npf, (int)N0), padded ? 1 : 0, padded ? 1.0 / p2 : 1.0, BUF_W2,
            &xs0, &xs1);
      xs0->wait_ev = last_ev;
    } else {  // line.py:193-258
      const int recvbuf = (padded || peer_mapped) ? BUF_W1 : BUF_OUT;
      SideT o;
      o.chunk = (int)kc; o.nchunk = P; o.nphys = (int)Nf;
      for (int q = 0; q < P; ++q) {
        o.base[q].buf = (q == me) ? recvbuf : BUF_W0;
        o.base[q].off = (q == me) ? (long long)me * pNp0 * Npf : (long long)pNp0 * koff[q];
        o.sb[q] = kcl[q]; o.si[q] = 1;
      }
      b.rows(true, pNp0, pN1, (int)Nf, BUF_IN, o);
      b.use(BUF_W0, (long long)pNp0 * Nf);
      b.use(recvbuf, (long long)pN0 * Npf);
      Step& x = b.exch(0, P, me);
      for (int q = 0; q < P; ++q) {
        x.send[q].buf = BUF_W0; x.send[q].off = (long long)pNp0 * koff[q]; x.scnt[q] = (long long)pNp0 * kcl[q];
        x.recv[q].buf = recvbuf; x.recv[q].off = (long long)q * pNp0 * Npf; x.rcnt[q] = (long long)pNp0 * Npf;
        x.rpeer[q].buf = recvbuf; x.rpeer[q].off = (long long)me * pNp0 * kcl[q];
      }
      xpass(0, Npf, nat(recvbuf, 0, 0, Npf, pN0), nat(BUF_OUT, 0, 0, Npf, (int)N0), padded ? 1 : 0, padded ? 1.0 / p2 : 1.0, BUF_W2,
            &xs0, &xs1);
    }
  } else {
    if (P == 1) {  // line.py:274-283
      xpass(1, Nf, nat(BUF_IN, 0, 0, Nf, (int)N0), nat(BUF_W0, 0, 0, Nf, pN0), 0, 1.0, BUF_W1, &xs0, &xs1);
      Step& sx = *xs0;
      b.use(BUF_W0, (long long)pN0 * Nf);
      if (masked) {
        sx.mask.on = 1;
        sx.mask.jdiv = 0x3fffffff;
        band(N0, false, sx.mask.i_lo, sx.mask.i_hi);
        band(N1, true, sx.mask.jr_lo, sx.mask.jr_hi);
      }
      b.rows(false, pN0, pN1, (int)Nf, BUF_OUT, nat(BUF_W0, 0, Nf, 1, (int)Nf), iscale);
    } else if (CH > 1) {  // line.py:285-338, pipelined: the x pass, then exchange(c) | z(c) over chunks of local rows
      const long long blk = (long long)pNp0 * Npf;
      SideT o;
      o.chunk = pNp0; o.nchunk = P; o.nphys = pN0;
      for (int q = 0; q < P; ++q) {
        o.base[q].buf = (q == me) ? BUF_W1 : BUF_W0;
        o.base[q].off = (q == me) ? (long long)pNp0 * koff[me] : q * blk;
        o.sb[q] = 0; o.si[q] = Npf;
      }
      b.fixed = 0;
      xpass(1, Npf, nat(BUF_IN, 0, 0, Npf, (int)N0), o, 0, 1.0, BUF_W2, &xs0, &xs1);
      Step& sx = *xs0;
      if (masked) {
        sx.mask.on = 1;
        sx.mask.jdiv = 0x3fffffff;
        band(N0, false, sx.mask.i_lo, sx.mask.i_hi);
        band(N1, true, sx.mask.jr_lo, sx.mask.jr_hi);
        sx.mask.jr_off = (int)(me * kc);
      }
      const int xev = xs1->rec_ev = pg.nevents++;
      b.use(BUF_W0, P * blk);
      b.use(BUF_W1, (long long)pNp0 * Nf);
      std::vector<int> eev((size_t)CH);
      for (int c = 0; c < CH; ++c) {
        const long long r0 = c * rc;
        b.fixed = 1;
        Step& x = b.exch(0, P, me);
        x.stream = 1;
        x.wait_ev = (c == 0) ? xev : -1;
        eev[(size_t)c] = x.rec_ev = pg.nevents++;
        for (int q = 0; q < P; ++q) {
          x.send[q].buf = BUF_W0; x.send[q].off = q * blk + r0 * Npf; x.scnt[q] = rc * Npf;
          x.recv[q].buf = BUF_W1; x.recv[q].off = (long long)pNp0 * koff[q] + r0 * kcl[q]; x.rcnt[q] = rc * kcl[q];
          x.rpeer[q].buf = BUF_W1; x.rpeer[q].off = (long long)pNp0 * koff[me] + r0 * Npf;
        }
      }
      for (int c = 0; c < CH; ++c) {
        const long long r0 = c * rc;
        SideT g;
        g.chunk = (int)kc; g.nchunk = P; g.nphys = (int)Nf;
        for (int q = 0; q < P; ++q) { g.base[q].buf = BUF_W1; g.base[q].off = (long long)pNp0 * koff[q] + r0 * kcl[q]; g.sb[q] = kcl[q]; g.si[q] = 1; }
        b.fixed = 2;
        Step& z = b.rows(false, rc, pN1, (int)Nf, BUF_OUT, g, iscale);
        z.real.off = r0 * pN1;
        z.wait_ev = eev[(size_t)c];
      }
    } else {  // line.py:285-338
      const long long blk = (long long)pNp0 * Npf;
      SideT o;
      o.chunk = pNp0; o.nchunk = P; o.nphys = pN0;
      for (int q = 0; q < P; ++q) {
        o.base[q].buf = (q == me) ? BUF_W1 : BUF_W0;
        o.base[q].off = (q == me) ? (long long)pNp0 * koff[me] : q * blk;
        o.sb[q] = 0; o.si[q] = Npf;
      }
      xpass(1, Npf, nat(BUF_IN, 0, 0, Npf, (int)N0), o, 0, 1.0, BUF_W2, &xs0, &xs1);
      Step& sx = *xs0;
      if (masked) {
        sx.mask.on = 1;
        sx.mask.jdiv = 0x3fffffff;
        band(N0, false, sx.mask.i_lo, sx.mask.i_hi);
        band(N1, true, sx.mask.jr_lo, sx.mask.jr_hi);
        sx.mask.jr_off = (int)(me * kc);
      }
      b.use(BUF_W0, P * blk);
      b.use(BUF_W1, (long long)pNp0 * Nf);
      Step& x = b.exch(0, P, me);
      for (int q = 0; q < P; ++q) {
        x.send[q].buf = BUF_W0; x.send[q].off = q * blk; x.scnt[q] = blk;
        x.recv[q].buf = BUF_W1; x.recv[q].off = (long long)pNp0 * koff[q]; x.rcnt[q] = (long long)pNp0 * kcl[q];
        x.rpeer[q].buf = BUF_W1; x.rpeer[q].off = (long long)pNp0 * koff[me];
      }
      SideT g;
      g.chunk = (int)kc; g.nchunk = P; g.nphys = (int)Nf;
      for (int q = 0; q < P; ++q) { g.base[q].buf = BUF_W1; g.base[q].off = (long long)pNp0 * koff[q]; g.sb[q] = kcl[q]; g.si[q] = 1; }
      b.rows(false, pNp0, pN1, (int)Nf, BUF_OUT, g, iscale);
    }
  }
  return peer_mapped ? finish_peer_mapped(d, pg) : 0;
}

// Build the step list of one (direction, dealias) program.  Mirrors oracle/slab.py etc.
inline int build_program(const b200fft_plan_desc_t& d, int inverse, int dealias, Program& pg) {
  switch (d.kind) {
    case B200FFT_SLAB:
    case B200FFT_SLAB_C2C:
      return build_slab(d, inverse, dealias, pg);
    case B200FFT_PENCIL_X:
    case B200FFT_PENCIL_Y:
      return build_pencil(d, inverse, dealias, pg);
    case B200FFT_LINE:
      return build_line(d, inverse, dealias, pg);
    default:
      return fail(B200FFT_ERR_ARG, "unknown plan kind %d", d.kind);
  }
}


}  // namespace b200fft
