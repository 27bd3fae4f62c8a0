// Batched 1D FFT kernels for sm_100a with the reference's pack / pad / truncate / mask copies
// fused into their load and store index maps.
//
//   strided C2C   : [B][n][J] -> [B'][n'][J]   FFT along the middle (strided) axis; replaces
//                   serialFFT fft/ifft(axis=0|1) (pyfftw_fft.py:26-39,115-128; numpy_fft.py:25-37),
//                   copy_to_padded/copy_from_padded (slab.py:516-536, pencil.py:351-379),
//                   transpose_Uc / rollaxis packs (maths.pyx:21-31, pencil.py:109-143),
//                   dealias_filter (maths.pyx:9-19) and the Alltoallw subarray datatypes
//                   (slab.py:199-211, pencil.py:218-246,971-999).
//   row R2C / C2R : contiguous rows; replaces rfft/irfft(axis=-1) (pyfftw_fft.py:71-83,160-173;
//                   numpy_fft.py:39-51) plus the z pad / truncate copies and the z-chunk pack.
//
// Algorithm: in-place mixed-radix decimation-in-frequency in shared memory.  Stage 0 reads its
// butterfly inputs straight from HBM into registers, the last stage writes straight from
// registers to HBM; the digit reversal is absorbed by assigning last-stage butterflies to
// threads in digit-reversed order so that both HBM sides stay coalesced.  One __syncthreads per
// stage boundary.  Tensor cores are not used: the work is a butterfly network bound by HBM.
//
// All phase bodies are __host__ __device__ and free of CUDA builtins so that tests/emu can run
// the identical index math on the CPU (never part of the product path).
#pragma once
#include "fft_radix.cuh"

namespace b200fft {

constexpr int MAXP = 16;  // max peers (chunks) of one exchange: 8 GPUs per box, headroom for 16

// ------------------------------------------------------------------------------------------
// compile-time radix plans
// ------------------------------------------------------------------------------------------
template <int I, int R0, int... Rs> struct NthRadix { static constexpr int value = NthRadix<I - 1, Rs...>::value; };
template <int R0, int... Rs> struct NthRadix<0, R0, Rs...> { static constexpr int value = R0; };

template <int I, int R0, int... Rs> struct PrefixProd { static constexpr int value = R0 * PrefixProd<I - 1, Rs...>::value; };
template <int R0, int... Rs> struct PrefixProd<0, R0, Rs...> { static constexpr int value = 1; };

template <int... Rs>
struct Plan {
  static constexpr int S = sizeof...(Rs);
  static constexpr int N = (1 * ... * Rs);
  template <int s> static constexpr int R = NthRadix<s, Rs...>::value;
  template <int s> static constexpr int L = N / PrefixProd<s, Rs..., 1>::value;  // sub-transform length at stage s
  template <int s> static constexpr int M = L<s> / R<s>;                         // butterfly stride at stage s
  static constexpr int RMAX = []() { int m = 1; for (int r : {Rs...}) m = r > m ? r : m; return m; }();
  static constexpr int RLAST = NthRadix<S - 1, Rs...>::value;

  // position (in the in-place DIF array) that holds frequency k after the last stage
  template <int s = 0> B2_HD static int pos(int k) {
    if constexpr (s >= S) {
      return 0;
    } else {
      constexpr int r = R<s>;
      return (k % r) * M<s> + pos<s + 1>(k / r);
    }
  }
};

// shared-memory swizzle: element `row` of a tile whose rows are ROWB bytes; keeps every access
// pattern of the DIF stages at the minimum number of 128-byte wavefronts (kernels).
template <int M0, int SW>
B2_HD int swz(int row) {
  if constexpr (SW > 1 && (M0 % SW) == 0) return row ^ ((row / M0) & (SW - 1));
  else return row;
}

// Twiddles of one butterfly: w[c] = W^c for c = 1..R-1 from ONE table load (W = tw[step]) and a
// product tree of depth <= log2(R) (w[c] = w[c/2] * w[c - c/2]).  A per-element table gather
// costs up to 32 L1 wavefronts per warp instruction and saturated the LSU pipe (profiles/r01a);
// the multiplies ride on the idle FP pipe.  Rounding: <= 4 products, ~5e-16 (double) per twiddle.
template <int R, class real>
B2_HD void twiddle_powers(cx<real>* w, const cx<real>* tw, int step) {
  if constexpr (R > 1) {
    w[1] = tw[step];
#pragma unroll
    for (int c = 2; c < R; ++c) w[c] = cmul(w[c / 2], w[c - c / 2]);
  }
}

// One DIF stage for the butterflies owned by thread `t` (of TC threads cooperating on one
// transform).  `sm` points at this transform's element 0 (column offset included); consecutive
// positions are RS elements apart.
template <class real, class P, int s, int TC, int RS, int SW, bool IN_FN, bool OUT_FN, class In, class Out>
B2_HD void fft_stage(int t, cx<real>* sm, const cx<real>* tw, int tws, In&& in, Out&& out, int fold_mode) {
  constexpr int S = P::S;
  constexpr int R = P::template R<s>;
  constexpr int L = P::template L<s>;
  constexpr int M = L / R;
  constexpr int M0 = P::template M<0>;
  constexpr int NB = P::N / R;
  constexpr int ROUNDS = (NB + TC - 1) / TC;
  using C = cx<real>;
#pragma unroll
  for (int rr = 0; rr < ROUNDS; ++rr) {
    const int u = t + rr * TC;
    if ((NB % TC) != 0 && u >= NB) break;
    int B, q;
    if constexpr (s == S - 1) {  // M == 1: butterflies taken in digit-reversed order
      B = (S > 1) ? P::template pos<0>(u) : 0;
      q = 0;
    } else {
      B = (u / M) * L;
      q = u % M;
    }
    C v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if constexpr (IN_FN) v[r] = in(B + q + r * M);
      else v[r] = sm[swz<M0, SW>(B + q + r * M) * RS];
    }
    Dft<R>::template run<1>(v);
    if constexpr (s < S - 1) {
      if constexpr (R >= 16 && sizeof(real) == 8) {
        // double-precision radix 16: 16 values + 16 twiddles would need > 128 registers; run the
        // powers as a chain (2 live twiddles) so that three CTAs stay resident per SM
        const C w1 = tw[q * tws * (P::N / L)];
        C wc = w1;
#pragma unroll
        for (int c = 1; c < R; ++c) {
          v[c] = cmul(v[c], wc);
          if (c + 1 < R) wc = cmul(wc, w1);
        }
      } else {
        C w[R];
        twiddle_powers<R>(w, tw, q * tws * (P::N / L));  // W_L^q = W_NTW^(q*(NTW/L))
#pragma unroll
        for (int c = 1; c < R; ++c) v[c] = cmul(v[c], w[c]);
      }
    }
    if constexpr (OUT_FN) {
      if constexpr (R % 3 == 0) {  // Nyquist fold of the 3/2-rule truncation (slab.py:480-482,529-533)
        if (fold_mode != 0 && u == 0) {  // slots R/3 and 2R/3 hold frequencies n/3 and 2n/3 (= -n/3)
          if (fold_mode == 1) v[R / 3] = v[2 * R / 3] = cadd(v[R / 3], v[2 * R / 3]);
          else if (fold_mode == 2) v[R / 3] = v[2 * R / 3];
          else v[2 * R / 3] = v[R / 3];
        }
      }
#pragma unroll
      for (int c = 0; c < R; ++c) out(u + c * NB, v[c]);
    } else {
#pragma unroll
      for (int c = 0; c < R; ++c) sm[swz<M0, SW>(B + q + c * M) * RS] = v[c];
    }
  }
}

// ------------------------------------------------------------------------------------------
// index maps
// ------------------------------------------------------------------------------------------
// One side (load or store) of a pass.  The transformed axis index i (after pad / truncate
// mapping to the physical extent nphys) is cut into `nchunk` chunks of `chunk` entries (the
// last one takes the remainder); chunk p lives at base[p] with its own pitches -- this is the
// per-peer block of an exchange (send layout on stores, receive layout on loads) or, for
// nchunk == 1, a plain strided array.  Offsets in elements:
//     base[p] + b*sb[p] + (i - p*chunk)*si[p] + j
struct Side {
  void* base[MAXP];
  long long sb[MAXP];
  long long si[MAXP];
  int chunk;
  int nchunk;
  int nphys;
  int jc;        // > 0: column j lives at (j / jc) * sj + (j % jc)  (kz-blocked arrays; strided passes only)
  long long sj;
};

// element offset of column j (each thread maps its own column, so a tile may straddle blocks)
B2_HD long long side_jmap(const Side& sd, long long j) { return sd.jc > 0 ? (j / sd.jc) * sd.sj + (j % sd.jc) : j; }

// 2/3-rule mask folded into a load (maths.pyx:9-19; masks slab.py:191-197, pencil.py:343-349,
// line.py:131-136).  An element is zeroed when any enabled band contains its index:
//   i + i_off in [i_lo, i_hi] ; b + b_off in [b_lo, b_hi] ;
//   (j / jdiv) + jq_off in [jq_lo, jq_hi] ; (j % jdiv) + jr_off in [jr_lo, jr_hi]
struct Mask {
  int on;
  int i_off, i_lo, i_hi;
  int b_off, b_lo, b_hi;
  int jdiv;
  int jq_off, jq_lo, jq_hi;
  int jr_off, jr_lo, jr_hi;
};

B2_HD int chunk_of(int i, int chunk, int nchunk) {
  int p = i / chunk;
  return p < nchunk ? p : nchunk - 1;
}

typedef unsigned long long addr_t;

// byte address of element (b, physical row r, column j) of a side
template <int CB>
B2_HD addr_t side_addr(const Side& sd, long long b, int r, long long j) {
  const int pc = (sd.nchunk > 1) ? chunk_of(r, sd.chunk, sd.nchunk) : 0;
  return (addr_t)sd.base[pc] + (addr_t)((b * sd.sb[pc] + (long long)(r - pc * sd.chunk) * sd.si[pc] + j) * CB);
}

// 16 / 8-byte asynchronous global -> shared copy (LDGSTS): no registers hold the data, so a CTA
// keeps its whole tile in flight; src_bytes == 0 zero-fills (pad rows, masked entries, dead lanes).
// l2hint: L2 prefetch size of the copy (cp.async ... .L2::128B / .L2::256B): a miss brings the whole 128- /
// 256-byte chunk around the source into L2, so that the neighbouring column tiles (other CTAs, moments
// later) hit in L2 and DRAM sees one wide read per row instead of one per tile.
template <int BYTES>
B2_HD void async_copy(void* dst_smem, addr_t src, bool valid, int l2hint = 0) {
#if defined(__CUDA_ARCH__)
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  const int nb = valid ? BYTES : 0;
  if constexpr (BYTES == 16) {
    if (l2hint == 0)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(src), "r"(nb) : "memory");
    else if (l2hint == 1)
      asm volatile("cp.async.cg.shared.global.L2::128B [%0], [%1], 16, %2;\n" ::"r"(d), "l"(src), "r"(nb) : "memory");
    else
      asm volatile("cp.async.cg.shared.global.L2::256B [%0], [%1], 16, %2;\n" ::"r"(d), "l"(src), "r"(nb) : "memory");
  } else {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(src), "r"(nb) : "memory");
  }
#else
  (void)l2hint;
  unsigned char* d = reinterpret_cast<unsigned char*>(dst_smem);
  const unsigned char* g = reinterpret_cast<const unsigned char*>(src);
  for (int i = 0; i < BYTES; ++i) d[i] = valid ? g[i] : (unsigned char)0;
#endif
}
B2_HD void async_copy_wait() {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.wait_all;\n" ::: "memory");
#endif
}

// Streaming ("touched once", evict-first) global accesses.  Used by the fused z+y kernel for the data that
// only passes through -- the real input of the forward z pass, the real output of the inverse one, the y
// pass's final stores -- so that it does not push the intermediate the two passes share out of L2.
template <class real>
B2_HD cx<real> load_streaming(const cx<real>* p) {
#if defined(__CUDA_ARCH__)
  if constexpr (sizeof(real) == 8) {
    const double2 v = __ldcs(reinterpret_cast<const double2*>(p));
    return cx<real>{v.x, v.y};
  } else {
    const float2 v = __ldcs(reinterpret_cast<const float2*>(p));
    return cx<real>{v.x, v.y};
  }
#else
  return *p;
#endif
}
template <class real>
B2_HD void store_streaming(cx<real>* p, cx<real> v) {
#if defined(__CUDA_ARCH__)
  if constexpr (sizeof(real) == 8) __stcs(reinterpret_cast<double2*>(p), make_double2(v.x, v.y));
  else __stcs(reinterpret_cast<float2*>(p), make_float2(v.x, v.y));
#else
  *p = v;
#endif
}

// ------------------------------------------------------------------------------------------
// fused pair of passes through L2 (persistent kernel)
// ------------------------------------------------------------------------------------------
// Two consecutive passes A, B over the same planes (forward: z then y; inverse: y then z) are cut into
// G groups of planes.  One persistent kernel executes the blocks of BOTH grids from a single queue,
//     A(0) | A(1) B(0) | A(2) B(1) | ... | B(G-1)          ("super-steps" 0 .. G)
// so that B(g) reads what A(g) wrote while it is still in L2 (a group is a few planes, tens of MB), A
// runs one group ahead and no launch boundary drains the SMs.  Work is counted in units -- rows for a
// row pass, blocks for a strided pass -- because a block of a row pass (RPC rows) may straddle two
// groups: such a block of A runs with the earlier group and credits both, such a block of B runs with
// the later group and waits for both.  done[g] counts finished units of pass A in group g; a block of
// B spins until its groups are complete.  Every block it can wait for sits earlier in the queue, i.e.
// is already running on some SM and never waits itself, so the queue cannot deadlock.
struct FuseSide {
  unsigned n;      // blocks in the pass's grid
  unsigned upb;    // units per block
  unsigned upg;    // units per group (>= upb)
  unsigned units;  // units in total
  // first block whose first / last unit lies in a group >= k
  B2_HD unsigned lo_first(unsigned k) const {
    const unsigned long long v = ((unsigned long long)k * upg + upb - 1) / upb;
    return v < n ? (unsigned)v : n;
  }
  B2_HD unsigned lo_last(unsigned k) const {
    const unsigned long long v = ((unsigned long long)k * upg) / upb;
    return v < n ? (unsigned)v : n;
  }
  // groups touched by block `blk` and its units in the first of them
  B2_HD void groups(unsigned blk, unsigned& g1, unsigned& g2, unsigned& u_in_g1, unsigned& u_in_g2) const {
    const unsigned u0 = blk * upb, u1 = (u0 + upb < units) ? u0 + upb : units;
    g1 = u0 / upg;
    g2 = (u1 - 1) / upg;
    u_in_g1 = (g1 == g2) ? u1 - u0 : g2 * upg - u0;
    u_in_g2 = (g1 == g2) ? 0 : u1 - g2 * upg;
  }
  B2_HD unsigned need(unsigned g) const {  // units of group g
    const unsigned long long lo = (unsigned long long)g * upg;
    return (unsigned)((lo + upg <= units) ? upg : units - lo);
  }
};

struct FuseCtl {
  unsigned* ctr;   // next queue index
  unsigned* done;  // per group: finished units of pass A
  unsigned G;
  FuseSide a, b;
};

// queue start of super-step k (k in 0 .. G+1): A blocks starting in groups < k and B blocks ending in groups < k-1
B2_HD unsigned fuse_qstart(const FuseCtl& c, unsigned k) {
  if (k == 0) return 0;
  const unsigned na = (k > c.G) ? c.a.n : c.a.lo_first(k);
  const unsigned nb = (k - 1 >= c.G) ? c.b.n : c.b.lo_last(k - 1);
  return na + nb;
}

// queue index -> (pass B?, block of that pass's grid)
B2_HD void fuse_decode(const FuseCtl& c, unsigned i, bool& isB, unsigned& blk) {
  // blocks per super-step ~ upg_a/upb_a + upg_b/upb_b: estimate, then correct
  const unsigned long long per = (unsigned long long)c.a.upg * c.b.upb + (unsigned long long)c.b.upg * c.a.upb;
  unsigned long long k64 = ((unsigned long long)i * c.a.upb * c.b.upb) / (per ? per : 1);
  unsigned k = k64 > c.G ? c.G : (unsigned)k64;
  while (k > 0 && fuse_qstart(c, k) > i) --k;
  while (k < c.G && fuse_qstart(c, k + 1) <= i) ++k;
  const unsigned off = i - fuse_qstart(c, k);
  const unsigned a0 = (k > c.G) ? c.a.n : c.a.lo_first(k);
  const unsigned a1 = (k + 1 > c.G) ? c.a.n : c.a.lo_first(k + 1);
  if (off < a1 - a0) {
    isB = false;
    blk = a0 + off;
  } else {
    isB = true;
    blk = c.b.lo_last(k - 1) + (off - (a1 - a0));  // (k >= 1 here: super-step 0 holds blocks of A only)
  }
}

// ------------------------------------------------------------------------------------------
// strided C2C pass
// ------------------------------------------------------------------------------------------
template <class real>
struct StridedParams {
  Side in, out;
  long long B;
  int J;
  int n;
  int inverse;
  int fold_mode;  // 0 none, 1 add (slab/pencil/line P>1), 2 replace (line P==1, line.py:189)
  real scale;
  Mask mask;
  const cx<real>* tw;
  int tws;  // table length / n
  int l2hint;  // L2 prefetch size of the tile loads: 0 none, 1 = 128 bytes, 2 = 256 bytes
};

// Tile geometry.  A CTA owns T adjacent columns (ROWB = T*sizeof(complex) contiguous bytes per
// row) of one batch entry, all n rows: n*ROWB bytes of shared memory plus an n-entry row-address
// table.  ROWB is the widest of 128 / 64 / 32 / 16 bytes that still leaves room for 3 (128 B) or
// 2 resident CTAs per SM -- while one CTA computes, the asynchronous loads of the others are in
// flight, which is what keeps HBM busy.
template <class real, class P, int MB = 0, int RB = 0>
struct StridedCfg {
  static constexpr int CB = (int)sizeof(cx<real>);
  static constexpr int LIM = 227 * 1024;
  static constexpr bool fits(int rowb, int k, bool tab) { return k * (P::N * rowb + (tab ? P::N * 8 : 0) + 1024) <= LIM; }
  static constexpr int ROWB = RB > 0                    ? RB
                              : (P::N * 128 <= 32 * 1024) ? 128
                              : fits(128, 3, true)      ? 128
                              : fits(64, 2, true)       ? 64
                              : fits(32, 2, true)       ? 32
                              : fits(32, 1, false)      ? 32
                              : fits(16, 1, false)      ? 16
                                                        : 8;  // one single-precision column (n = 16384)
  static constexpr bool TAB = fits(ROWB, 1, true);
  static constexpr int T0 = ROWB / CB < 1 ? 1 : ROWB / CB;
  static constexpr int NBMIN = P::N / P::RMAX;
  static constexpr int pow2floor(int x) { int p = 1; while (2 * p <= x) p *= 2; return p; }
  static constexpr int TC_ = pow2floor(NBMIN) < (256 / T0) ? pow2floor(NBMIN) : (256 / T0);
  static constexpr int TC = TC_ < 1 ? 1 : TC_;
  static constexpr int T = (T0 * TC >= 128) ? T0 : (128 / TC);
  static constexpr int NT = T * TC;
  static constexpr int SW = (T * CB >= 128) ? 1 : 128 / (T * CB);
  static constexpr int TILE = P::N * T * CB;
  static constexpr int SMEM = TILE + (TAB ? P::N * 8 : 0);
  static constexpr int NPHASE = P::S + 3;
  static constexpr int MINB_ = LIM / (SMEM + 1024);
  static constexpr int MINB = MB > 0 ? MB : (MINB_ < 1 ? 1 : (MINB_ > 3 ? 3 : MINB_));
};

// JS: the sides may use the blocked column map (Side::jc); compiled for the benchmark lengths only
template <class real, class P, int MB = 0, int RB = 0, bool STREAM_ST = false, bool JS = false>
struct StridedK {
  using Cfg = StridedCfg<real, P, MB, RB>;
  using C = cx<real>;
  using Params = StridedParams<real>;
  static constexpr int NPHASE = Cfg::NPHASE;
  static constexpr int NT = Cfg::NT;
  static constexpr int SMEM = Cfg::SMEM;
  static constexpr int SMEM1 = Cfg::SMEM;
  static constexpr bool PIPE = false;
  static constexpr int MINB = Cfg::MINB;

  // 1D grid: consecutive blocks walk the column tiles of one batch entry (adjacent ROWB-byte
  // segments of the same rows -> neighbouring CTAs share DRAM pages and L2 lines).
  B2_HD static unsigned long long blocks(const Params& p) {
    return (unsigned long long)((p.J + Cfg::T - 1) / Cfg::T) * (unsigned long long)p.B;
  }
  B2_HD static void decode(const Params& p, unsigned blk, int& bx, int& by) {
    const unsigned nt = (unsigned)((p.J + Cfg::T - 1) / Cfg::T);
    by = (int)(blk / nt);
    bx = (int)(blk - (unsigned)by * nt);
  }
  // work units of the fused pair kernel: one per block, (column tiles) per plane
  B2_HD static FuseSide fuse_side(const Params& p, long long planes, long long planes_per_group) {
    const unsigned nt = (unsigned)((p.J + Cfg::T - 1) / Cfg::T);
    (void)planes;
    return FuseSide{(unsigned)blocks(p), 1u, (unsigned)(planes_per_group * nt), (unsigned)blocks(p)};
  }

  // address of load row i (logical index of the transform input) for column j0; 0 = zero fill:
  // copy_to_padded (slab.py:517-523) and the kx / ky band of the 2/3-rule mask
  B2_HD static addr_t in_row(const Params& p, long long b, int i, long long j0) { return in_row_n<P::N>(p, b, i, j0); }
  B2_HD static addr_t out_row(const Params& p, long long b, int k, long long j0) { return out_row_n<P::N>(p, b, k, j0); }
  // (n: length of the whole transform -- the cluster kernel runs half-length plans on each CTA)
  template <int n>
  B2_HD static addr_t in_row_n(const Params& p, long long b, int i, long long j0) {
    int ip = i;
    if (p.in.nphys < n) {
      const int h = p.in.nphys / 2;
      if (i >= h) {
        if (i < n - h) return 0;
        ip = i - (n - p.in.nphys);
      }
    }
    if (p.mask.on) {
      const int ii = ip + p.mask.i_off;
      if (ii >= p.mask.i_lo && ii <= p.mask.i_hi) return 0;
    }
    return side_addr<Cfg::CB>(p.in, b, ip, j0);
  }
  // address of store row for output frequency k; 0 = dropped: copy_from_padded (slab.py:529-533).
  // The inverse transform is the forward one with the output index reversed (k -> -k mod n).
  template <int n>
  B2_HD static addr_t out_row_n(const Params& p, long long b, int k, long long j0) {
    int kp = p.inverse ? (k == 0 ? 0 : n - k) : k;
    if (p.out.nphys < n) {
      const int h = p.out.nphys / 2;
      if (kp > h) {
        if (kp <= n - h) return 0;
        kp -= n - p.out.nphys;
      }
    }
    return side_addr<Cfg::CB>(p.out, b, kp, j0);
  }

  template <int s>
  B2_HD static void phase(const Params& p, void* smraw, int tid, int bx, int by) {
    constexpr int T = Cfg::T, n = P::N, CB = Cfg::CB;
    constexpr int M0 = P::template M<0>;
    const int c = tid % T;
    const int t = tid / T;
    const int j0 = bx * T;
    const long long b = by;
    const bool live = j0 + c < p.J;
    addr_t* tab = reinterpret_cast<addr_t*>(reinterpret_cast<unsigned char*>(smraw) + Cfg::TILE);

    // row addresses are taken at column jin / jout and every thread adds its own column offset cin / cout;
    // with blocked column layouts (JS) the rows are taken at column 0 and each column is mapped by itself
    const long long jin = JS ? 0 : j0, jout = JS ? 0 : j0;
    const long long cin = JS ? side_jmap(p.in, (long long)j0 + c) : c, cout = JS ? side_jmap(p.out, (long long)j0 + c) : c;
    if constexpr (s == 0) {  // load-row address table
      if constexpr (Cfg::TAB)
        for (int i = tid; i < n; i += Cfg::NT) tab[i] = in_row(p, b, i, jin);
    } else if constexpr (s == 1) {  // whole tile global -> shared, asynchronously
      bool colzero = !live;
      if (p.mask.on && live) {
        const Mask& m = p.mask;
        const int j = j0 + c;
        const int bb = (int)b + m.b_off, jq = j / m.jdiv + m.jq_off, jr = j % m.jdiv + m.jr_off;
        if ((bb >= m.b_lo && bb <= m.b_hi) || (jq >= m.jq_lo && jq <= m.jq_hi) || (jr >= m.jr_lo && jr <= m.jr_hi))
          colzero = true;
      }
      C* sm = reinterpret_cast<C*>(smraw) + c;
      const addr_t fallback = (addr_t)p.tw;  // any valid address: never dereferenced when size == 0
#pragma unroll 4
      for (int i = t; i < n; i += Cfg::TC) {
        addr_t a;
        if constexpr (Cfg::TAB) a = tab[i];
        else a = in_row(p, b, i, jin);
        const bool ok = (a != 0) && !colzero;
        async_copy<CB>(sm + swz<M0, Cfg::SW>(i) * T, ok ? a + (addr_t)(cin * CB) : fallback, ok, p.l2hint);
      }
    } else if constexpr (s == 2) {  // store-row address table (overlaps the loads in flight)
      if constexpr (Cfg::TAB)
        for (int k = tid; k < n; k += Cfg::NT) tab[k] = out_row(p, b, k, jout);
      async_copy_wait();
    } else {
      constexpr int st = s - 3;
      C* sm = reinterpret_cast<C*>(smraw) + c;
      auto in = [](int) -> C { return C{0, 0}; };
      auto out = [&](int k, C v) {
        if (!live) return;
        addr_t a;
        if constexpr (Cfg::TAB) a = tab[k];
        else a = out_row(p, b, k, jout);
        if (a == 0) return;
        if (p.scale != (real)1) v = cscale(v, p.scale);
        if constexpr (STREAM_ST) store_streaming(reinterpret_cast<C*>(a + (addr_t)(cout * CB)), v);
        else *reinterpret_cast<C*>(a + (addr_t)(cout * CB)) = v;
      };
      // reversed output index: the slot that survives a "keep -N/2" truncation is the other one
      const int fold = (p.inverse && p.fold_mode == 2) ? 3 : p.fold_mode;
      fft_stage<real, P, st, Cfg::TC, T, Cfg::SW, false, (st == P::S - 1)>(t, sm, p.tw, p.tws, in, out, fold);
    }
  }
};


// Strided pass with its first stage fed straight from HBM ("direct" form, B200FFT_VARIANT=35).  StridedK
// stages the whole tile in shared memory with asynchronous copies before the first butterflies read it
// back; ncu r01c shows the pass bound by LSU wavefronts (77 %), two of whose six shared-memory accesses
// per element are that staging.  Here a thread loads the R inputs of its first-stage butterflies from
// the row-address table straight into registers (all loads issued before the first use: 16 x 16 bytes
// per thread in flight for a radix-16 stage), so the tile makes its first trip through shared memory as
// the first stage's OUTPUT.  Same tile geometry, index maps and later stages; one barrier less.
template <class real, class P, int MB = 0, int RB = 0>
struct StridedDK {
  using SK = StridedK<real, P, MB, RB>;
  using Cfg = typename SK::Cfg;
  using C = cx<real>;
  using Params = StridedParams<real>;
  static_assert(Cfg::TAB, "the direct form keeps its row-address table in shared memory");
  static_assert(P::S >= 2, "the direct form needs a later stage to store from");
  static constexpr int NPHASE = P::S + 2;
  static constexpr int NT = Cfg::NT;
  static constexpr int SMEM = Cfg::SMEM;
  static constexpr int SMEM1 = Cfg::SMEM;
  static constexpr bool PIPE = false;
  static constexpr int MINB = Cfg::MINB;
  B2_HD static unsigned long long blocks(const Params& p) { return SK::blocks(p); }
  B2_HD static void decode(const Params& p, unsigned blk, int& bx, int& by) { SK::decode(p, blk, bx, by); }

  // phases: 0 load-row table; 1 stage 0 (HBM -> registers -> shared); 2 store-row table; 3.. stages 1..S-1.
  // Plans with at least two stages (a single stage would have to store from phase 1).
  template <int s>
  B2_HD static void phase(const Params& p, void* smraw, int tid, int bx, int by) {
    constexpr int T = Cfg::T, n = P::N, CB = Cfg::CB;
    const int c = tid % T;
    const int t = tid / T;
    const int j0 = bx * T;
    const long long b = by;
    const bool live = j0 + c < p.J;
    addr_t* tab = reinterpret_cast<addr_t*>(reinterpret_cast<unsigned char*>(smraw) + Cfg::TILE);
    C* sm = reinterpret_cast<C*>(smraw) + c;
    auto no_in = [](int) -> C { return C{0, 0}; };
    auto no_out = [](int, C) {};
    auto store = [&](int k, C v) {
      if (!live) return;
      const addr_t a = tab[k];
      if (a == 0) return;
      if (p.scale != (real)1) v = cscale(v, p.scale);
      *reinterpret_cast<C*>(a + (addr_t)c * CB) = v;
    };
    const int fold = (p.inverse && p.fold_mode == 2) ? 3 : p.fold_mode;
    if constexpr (s == 0) {
      for (int i = tid; i < n; i += Cfg::NT) tab[i] = SK::in_row(p, b, i, j0);
    } else if constexpr (s == 1) {
      bool colzero = !live;
      if (p.mask.on && live) {
        const Mask& m = p.mask;
        const int j = j0 + c;
        const int bb = (int)b + m.b_off, jq = j / m.jdiv + m.jq_off, jr = j % m.jdiv + m.jr_off;
        if ((bb >= m.b_lo && bb <= m.b_hi) || (jq >= m.jq_lo && jq <= m.jq_hi) || (jr >= m.jr_lo && jr <= m.jr_hi))
          colzero = true;
      }
      auto in = [&](int i) -> C {  // pad rows, masked rows / columns and dead lanes read as zero
        const addr_t a = tab[i];
        return (a != 0 && !colzero) ? *reinterpret_cast<const C*>(a + (addr_t)c * CB) : C{0, 0};
      };
      // (a single-stage plan would need the store table here: such lengths stay with StridedK, see dispatch)
      fft_stage<real, P, 0, Cfg::TC, T, Cfg::SW, true, false>(t, sm, p.tw, p.tws, in, no_out, 0);
    } else if constexpr (s == 2) {
      for (int k = tid; k < n; k += Cfg::NT) tab[k] = SK::out_row(p, b, k, j0);
    } else {
      constexpr int st = s - 2;
      fft_stage<real, P, st, Cfg::TC, T, Cfg::SW, false, (st == P::S - 1)>(t, sm, p.tw, p.tws, no_in, store, fold);
    }
  }
};

// ------------------------------------------------------------------------------------------
// strided C2C pass on a 2-CTA cluster ("far" strides: the x pass of a slab)
// ------------------------------------------------------------------------------------------
// When rows of the transformed axis are >= 1 MB apart every row segment of a tile lies in its own
// 2 MB page and the pass is bound by address translations per byte: 128-byte row
// segments halve them, but a whole n x 128 B column tile leaves room for one CTA per SM only.
// Here a column tile of n = 2H rows x 128 bytes is split over the two CTAs of a cluster BY ROWS:
// CTA r loads rows [rH, (r+1)H) (H x 128 B, the footprint of today's 64-byte tiles, so the same
// number of CTAs stay resident), the pair runs the first radix-2 DIF stage across distributed shared
// memory,
//     s[i] = a[i] + a[i+H]            -> CTA 0 (even output frequencies 2k')
//     d[i] = (a[i] - a[i+H]) W_n^i    -> CTA 1 (odd output frequencies 2k'+1)
// and each CTA finishes with an independent H-point transform of its half and stores H rows of 128
// bytes.  Every (i, column) pair of the cross stage is read and rewritten by exactly one thread
// (CTA r takes i in [rH/2, (r+1)H/2)), so it needs no barrier between its loads and stores: one
// cluster barrier before it (both tiles have landed) and one after it (all remote writes are done).
// Pad / truncate / fold / mask / inverse index maps are those of StridedK, evaluated for the full
// length n; the two folded modes +-N/2 are even frequencies and stay in one butterfly of CTA 0.
// RB = 64 gives long columns (n >= 2048) near-stride tiles of today's width but half the rows per CTA,
// i.e. three resident CTAs where the whole column in one CTA leaves room for one.
template <class real, class PS, int RB = 128>
struct ClusterStridedK {
  using SK = StridedK<real, PS, 0, RB>;
  using Cfg = typename SK::Cfg;
  using C = cx<real>;
  using Params = StridedParams<real>;
  static constexpr int H = PS::N, N = 2 * PS::N;
  static constexpr int CLUSTER = 2;
  static constexpr int NPHASE = PS::S + 4;
  static constexpr int SYNC_BEFORE = 3;  // cluster-wide barriers before and after this phase
  static constexpr int NT = Cfg::NT;
  static constexpr int SMEM = Cfg::SMEM;
  static constexpr int MINB = Cfg::MINB;
  static_assert(Cfg::TAB, "cluster tiles keep their row-address table in shared memory");
  static_assert(Cfg::T * Cfg::CB == RB, "tile rows are RB bytes");

  B2_HD static unsigned long long blocks(const Params& p) { return 2ull * SK::blocks(p); }
  B2_HD static void decode(const Params& p, unsigned blk, int& bx, int& by) { SK::decode(p, blk / 2, bx, by); }

  // `sm`: this CTA's tile, `peer`: the other CTA's tile (distributed shared memory), `rank`: 0 / 1
  template <int s>
  B2_HD static void phase(const Params& p, void* smraw, void* peerraw, int tid, int rank, int bx, int by) {
    constexpr int T = Cfg::T, CB = Cfg::CB, TC = Cfg::TC, SW = Cfg::SW;
    constexpr int M0 = PS::template M<0>;
    const int c = tid % T;
    const int t = tid / T;
    const int j0 = bx * T;
    const long long b = by;
    const bool live = j0 + c < p.J;
    addr_t* tab = reinterpret_cast<addr_t*>(reinterpret_cast<unsigned char*>(smraw) + Cfg::TILE);

    if constexpr (s == 0) {  // load-row addresses of this CTA's half
      for (int i = tid; i < H; i += NT) tab[i] = SK::template in_row_n<N>(p, b, rank * H + i, j0);
    } else if constexpr (s == 1) {  // H rows x 128 bytes, global -> shared, asynchronously
      bool colzero = !live;
      if (p.mask.on && live) {
        const Mask& m = p.mask;
        const int j = j0 + c;
        const int bb = (int)b + m.b_off, jq = j / m.jdiv + m.jq_off, jr = j % m.jdiv + m.jr_off;
        if ((bb >= m.b_lo && bb <= m.b_hi) || (jq >= m.jq_lo && jq <= m.jq_hi) || (jr >= m.jr_lo && jr <= m.jr_hi))
          colzero = true;
      }
      C* sm = reinterpret_cast<C*>(smraw) + c;
      const addr_t fallback = (addr_t)p.tw;
#pragma unroll 4
      for (int i = t; i < H; i += TC) {
        const addr_t a = tab[i];
        const bool ok = (a != 0) && !colzero;
        async_copy<CB>(sm + swz<M0, SW>(i) * T, ok ? a + (addr_t)c * CB : fallback, ok);
      }
    } else if constexpr (s == 2) {  // store-row addresses: CTA `rank` owns output frequencies 2k' + rank
      for (int k = tid; k < H; k += NT) tab[k] = SK::template out_row_n<N>(p, b, 2 * k + rank, j0);
      async_copy_wait();
    } else if constexpr (s == 3) {  // radix-2 stage across the pair
      C* own = reinterpret_cast<C*>(smraw) + c;
      C* oth = reinterpret_cast<C*>(peerraw) + c;
      C* lo = rank == 0 ? own : oth;  // rows i      (CTA 0's tile)
      C* hi = rank == 0 ? oth : own;  // rows i + H  (CTA 1's tile)
#pragma unroll 4
      for (int i = rank * (H / 2) + t; i < (rank + 1) * (H / 2); i += TC) {
        const int ps = swz<M0, SW>(i) * T;  // both tiles use the sub-plan's swizzle
        const C a = lo[ps], bq = hi[ps];
        lo[ps] = cadd(a, bq);
        hi[ps] = cmul(csub(a, bq), p.tw[i * p.tws]);
      }
    } else {
      constexpr int st = s - 4;
      C* sm = reinterpret_cast<C*>(smraw) + c;
      auto in = [](int) -> C { return C{0, 0}; };
      auto out = [&](int k, C v) {
        if (!live) return;
        const addr_t a = tab[k];
        if (a == 0) return;
        if (p.scale != (real)1) v = cscale(v, p.scale);
        *reinterpret_cast<C*>(a + (addr_t)c * CB) = v;
      };
      const int fold = rank != 0 ? 0 : (p.inverse && p.fold_mode == 2) ? 3 : p.fold_mode;
      fft_stage<real, PS, st, TC, T, SW, false, (st == PS::S - 1)>(t, sm, p.tw, 2 * p.tws, in, out, fold);
    }
  }
};

// ------------------------------------------------------------------------------------------
// contiguous-row R2C / C2R passes (half-length complex FFT + split / merge step)
// ------------------------------------------------------------------------------------------
template <class real>
struct RowParams {
  Side cside;          // complex side: out for R2C, in for C2R; chunked along k, b = row
  const void* rin;     // R2C: real input rows
  void* rout;          // C2R: real output rows
  long long rpitch;    // real row pitch (elements)
  long long rows;
  int n;               // real length (= 2*P::N)
  int nk;              // number of complex entries kept (R2C, truncation) / present (C2R, zero pad)
  real scale;
  const cx<real>* tw;  // table of W_NTW^j
  int tws;             // NTW / n
};

// Threads per row TC and rows per CTA.  TC is the largest lane count (a multiple of a warp where
// the row is long enough) that keeps >= 85% of the butterfly slots of every stage busy; the CTA
// takes as many rows as fit a ~32 KB buffer (<= 256 threads).  Row kernels are persistent and
// double-buffered: while a CTA transforms one group of rows, the asynchronous loads of its next
// group are in flight, so every resident CTA always has a buffer's worth of HBM reads outstanding.
// CAPV > 0 overrides the resident-CTA target that sets the register budget (see MINB_CAP)
template <class real, class P, int CAPV = 0>
struct RowCfg {
  static constexpr int CB = (int)sizeof(cx<real>);
  static constexpr int H = P::N;
  static constexpr int SROW = H + 16 / CB;  // H slots (swizzled) + one for X[H] (C2R), rows stay 16-byte aligned
  template <int s = 0>
  static constexpr long long slots(int tc) {  // sum over stages of rounds * radix
    if constexpr (s >= P::S) {
      return 0;
    } else {
      constexpr int R = P::template R<s>;
      return (long long)(((H / R) + tc - 1) / tc) * R + slots<s + 1>(tc);
    }
  }
  static constexpr bool good(int tc) {  // S*H useful element slots out of tc * slots(tc)
    return tc >= 1 && tc <= 256 && (tc == 1 || tc * 4 <= H) && 100LL * P::S * H >= 85LL * tc * slots(tc);
  }
  static constexpr int pick() {
    constexpr int cand[] = {256, 192, 128, 96, 64, 48, 32, 24, 16, 12, 8, 6, 4, 3, 2, 1};
    for (int c : cand)
      if (good(c)) return c;
    return 1;
  }
  static constexpr int TC = pick();
  static constexpr int RPC_T = (256 / TC) < 1 ? 1 : (256 / TC);                          // rows per CTA by threads
  static constexpr int RPC_S = (32 * 1024) / (SROW * CB) < 1 ? 1 : (32 * 1024) / (SROW * CB);  // by shared memory
  static constexpr int RPC_ = RPC_T < RPC_S ? RPC_T : RPC_S;
  static constexpr int RPC = (RPC_ * TC >= 32) ? ((RPC_ * TC) / 32 * 32) / TC : RPC_;  // whole warps
  static constexpr int NT = TC * RPC;
  // threads sharing one row; they may use a barrier of their own when they are whole warps
  static constexpr bool ROWBAR_OK = RPC > 1 && TC % 32 == 0 && RPC <= 15;
  static constexpr int SW = 128 / CB;
  static constexpr int SMEM1 = RPC * SROW * CB;
  static constexpr bool PIPE = 2 * SMEM1 + 1024 <= 227 * 1024;
  static constexpr int SMEM = PIPE ? 2 * SMEM1 : SMEM1;
  static constexpr int MINB_ = (227 * 1024) / (SMEM + 1024);
  // register budget: double-precision radix-16 butterflies need ~128 registers, radix-12 ~96
  static constexpr int MINB_CAP = CAPV > 0 ? CAPV : (P::RMAX >= 16 && CB == 16) ? 2 : (P::RMAX >= 12 ? 3 : 4);
  static constexpr int MINB = MINB_ < 1 ? 1 : (MINB_ > MINB_CAP ? MINB_CAP : MINB_);
};

template <class real, class P, bool STREAM_LD = false, int CAPV = 0>
struct R2CK {  // forward: real rows -> complex rows
  static constexpr int GROUP = RowCfg<real, P>::TC;
  // Stage 0 loads its butterfly inputs straight from HBM into registers (coalesced 16-byte loads,
  // all issued before the first use): staging the row through shared memory first, as C2R does,
  // measured 20% slower here because the kernel is bound by LSU wavefronts, not by load latency.
  using Cfg = RowCfg<real, P, CAPV>;
  using C = cx<real>;
  using Params = RowParams<real>;
  static constexpr int NPHASE = P::S + 1;
  static constexpr int NT = Cfg::NT;
  static constexpr int SMEM = Cfg::SMEM1;
  static constexpr int SMEM1 = Cfg::SMEM1;
  static constexpr bool PIPE = false;
  static constexpr int MINB_ = (227 * 1024) / (SMEM + 1024);
  static constexpr int MINB = MINB_ < 1 ? 1 : (MINB_ > Cfg::MINB_CAP ? Cfg::MINB_CAP : MINB_);
  B2_HD static unsigned long long blocks(const Params& p) {
    return (unsigned long long)((p.rows + Cfg::RPC - 1) / Cfg::RPC);
  }
  B2_HD static void decode(const Params&, unsigned blk, int& bx, int& by) {
    bx = (int)blk;
    by = 0;
  }
  // work units of the fused pair kernel: rows (RPC per block)
  B2_HD static FuseSide fuse_side(const Params& p, long long planes, long long planes_per_group) {
    return FuseSide{(unsigned)blocks(p), (unsigned)Cfg::RPC, (unsigned)(planes_per_group * (p.rows / planes)), (unsigned)p.rows};
  }

  template <int s>
  B2_HD static void phase(const Params& p, void* smraw, int tid, int bx, int) {
    constexpr int H = Cfg::H, TC = Cfg::TC;
    constexpr int M0 = P::template M<0>;
    const int rl = tid / TC, t = tid % TC;
    const long long row = (long long)bx * Cfg::RPC + rl;
    const bool live = row < p.rows;
    C* sm = reinterpret_cast<C*>(smraw) + rl * Cfg::SROW;
    if constexpr (s < P::S) {
      const C* src = reinterpret_cast<const C*>(reinterpret_cast<const real*>(p.rin) + row * p.rpitch);
      auto in = [&](int i) -> C {  // z[i] = x[2i] + i*x[2i+1]
        if (!live) return C{0, 0};
        if constexpr (STREAM_LD) return load_streaming(src + i);
        else return src[i];
      };
      auto out = [](int, C) {};
      // every stage writes shared memory (the split step needs all of F)
      fft_stage<real, P, s, TC, 1, Cfg::SW, (s == 0), false>(t, sm, p.tw, 2 * p.tws, in, out, 0);
    } else {
      // split step: X[k] = E[k] + W_n^k O[k],  X[H-k] = conj(E[k] - W_n^k O[k])
      if (!live) return;
      const Side& o = p.cside;
      auto store = [&](int k, C v) {
        if (k >= p.nk) return;  // z truncation of the 3/2-rule: copy_from_padded axis 2 (slab.py:535)
        const int pc = (o.nchunk > 1) ? chunk_of(k, o.chunk, o.nchunk) : 0;
        C* ptr = reinterpret_cast<C*>(o.base[pc]) + row * o.sb[pc] + (k - pc * o.chunk);
        *ptr = (p.scale != (real)1) ? cscale(v, p.scale) : v;
      };
      constexpr int NK = H / 2 + 1;
      constexpr int ROUNDS = (NK + TC - 1) / TC;
#pragma unroll
      for (int rr = 0; rr < ROUNDS; ++rr) {
        const int k = t + rr * TC;
        if (k >= NK) break;
        if (k == 0) {
          C f = sm[swz<M0, Cfg::SW>(0)];
          store(0, C{f.x + f.y, 0});
          store(H, C{f.x - f.y, 0});
        } else {
          const C a = sm[swz<M0, Cfg::SW>(P::template pos<0>(k))];
          const C bq = sm[swz<M0, Cfg::SW>(P::template pos<0>(H - k))];
          const C e = C{(real)0.5 * (a.x + bq.x), (real)0.5 * (a.y - bq.y)};   // (a + conj b)/2
          const C d = C{(real)0.5 * (a.x - bq.x), (real)0.5 * (a.y + bq.y)};   // (a - conj b)/2
          const C o2 = mul_mi(d);                                              // O = -i (a - conj b)/2
          const C wo = cmul(p.tw[k * p.tws], o2);
          store(k, cadd(e, wo));
          if (k != H - k) store(H - k, cconj(csub(e, wo)));
        }
      }
    }
  }
};

// R2C with the split step folded into the last stage ("paired" form, B200FFT_VARIANT=33).  R2CK writes the
// half-length spectrum F to shared memory after its last stage only to read F[k] and F[H-k] back in the split
// step.  The last-stage butterfly u holds the frequencies u + c*NB, their mirrors H - (u + c*NB) sit in
// butterfly NB - u at slot R-1-c: a thread that runs BOTH butterflies has every pair in registers and stores
// X[k], X[H-k] straight to HBM.  That removes one shared-memory write and one read of the whole row (a third
// of the shared-memory traffic of an LSU-bound kernel, ncu r01c) and one barrier; the price is 2R values in
// registers and half the threads idle in the last phase.  Plans with at least two stages.
template <class real, class P>
struct R2CPK {
  static_assert(P::S >= 2, "the paired form needs a last stage that reads shared memory");
  using Cfg = RowCfg<real, P, (sizeof(real) == 8 ? 3 : 0)>;  // double: 2R = 16..24 complex values live -> three-CTA register budget
  static constexpr int GROUP = Cfg::TC;
  using C = cx<real>;
  using Params = RowParams<real>;
  static constexpr int NPHASE = P::S;
  static constexpr int NT = Cfg::NT;
  static constexpr int SMEM = Cfg::SMEM1;
  static constexpr int SMEM1 = Cfg::SMEM1;
  static constexpr bool PIPE = false;
  static constexpr int MINB_ = (227 * 1024) / (SMEM + 1024);
  static constexpr int MINB = MINB_ < 1 ? 1 : (MINB_ > Cfg::MINB_CAP ? Cfg::MINB_CAP : MINB_);
  B2_HD static unsigned long long blocks(const Params& p) { return (unsigned long long)((p.rows + Cfg::RPC - 1) / Cfg::RPC); }
  B2_HD static void decode(const Params&, unsigned blk, int& bx, int& by) {
    bx = (int)blk;
    by = 0;
  }

  template <int s>
  B2_HD static void phase(const Params& p, void* smraw, int tid, int bx, int) {
    constexpr int H = Cfg::H, TC = Cfg::TC;
    constexpr int M0 = P::template M<0>;
    const int rl = tid / TC, t = tid % TC;
    const long long row = (long long)bx * Cfg::RPC + rl;
    const bool live = row < p.rows;
    C* sm = reinterpret_cast<C*>(smraw) + rl * Cfg::SROW;
    if constexpr (s < P::S - 1) {
      const C* src = reinterpret_cast<const C*>(reinterpret_cast<const real*>(p.rin) + row * p.rpitch);
      auto in = [&](int i) -> C { return live ? src[i] : C{0, 0}; };
      auto out = [](int, C) {};
      fft_stage<real, P, s, TC, 1, Cfg::SW, (s == 0), false>(t, sm, p.tw, 2 * p.tws, in, out, 0);
    } else {
      if (!live) return;
      constexpr int R = P::template R<P::S - 1>;
      constexpr int NB = H / R;
      constexpr int ITEMS = NB / 2 + 1;
      constexpr int ROUNDS = (ITEMS + TC - 1) / TC;
      const Side& o = p.cside;
      auto store = [&](int k, C v) {
        if (k >= p.nk) return;  // z truncation of the 3/2-rule (slab.py:535)
        const int pc = (o.nchunk > 1) ? chunk_of(k, o.chunk, o.nchunk) : 0;
        C* ptr = reinterpret_cast<C*>(o.base[pc]) + row * o.sb[pc] + (k - pc * o.chunk);
        *ptr = (p.scale != (real)1) ? cscale(v, p.scale) : v;
      };
      // X[k] = E + W_n^k O, X[H-k] = conj(E - W_n^k O) from a = F[k], b = F[H-k]   (0 < k < H, k != H-k handled by caller)
      auto split = [&](int k, C a, C bq) {
        const C e = C{(real)0.5 * (a.x + bq.x), (real)0.5 * (a.y - bq.y)};
        const C d = C{(real)0.5 * (a.x - bq.x), (real)0.5 * (a.y + bq.y)};
        const C wo = cmul(p.tw[k * p.tws], mul_mi(d));
        store(k, cadd(e, wo));
        if (k != H - k) store(H - k, cconj(csub(e, wo)));
      };
#pragma unroll
      for (int rr = 0; rr < ROUNDS; ++rr) {
        const int it = t + rr * TC;
        if (it >= ITEMS) break;
        const int u1 = it, u2 = (NB - it) % NB;
        C v1[R], v2[R];
        const int B1 = P::template pos<0>(u1), B2 = P::template pos<0>(u2);
#pragma unroll
        for (int r = 0; r < R; ++r) v1[r] = sm[swz<M0, Cfg::SW>(B1 + r)];
        Dft<R>::template run<1>(v1);
        if (u2 != u1) {
#pragma unroll
          for (int r = 0; r < R; ++r) v2[r] = sm[swz<M0, Cfg::SW>(B2 + r)];
          Dft<R>::template run<1>(v2);
#pragma unroll
          for (int c = 0; c < R; ++c) split(u1 + c * NB, v1[c], v2[R - 1 - c]);
        } else if (u1 == 0) {  // frequencies c*NB: mirror (R-c)*NB in the same butterfly; k = 0 carries X[0] and X[H]
          store(0, C{v1[0].x + v1[0].y, 0});
          store(H, C{v1[0].x - v1[0].y, 0});
#pragma unroll
          for (int c = 1; c <= R / 2; ++c) split(c * NB, v1[c], v1[R - c]);
        } else {  // u = NB/2: frequencies NB/2 + c*NB, mirror at slot R-1-c of the same butterfly
#pragma unroll
          for (int c = 0; c < (R + 1) / 2; ++c) split(u1 + c * NB, v1[c], v1[R - 1 - c]);
        }
      }
    }
  }
};

template <class real, class P, bool STREAM_ST = false, int CAPV = 0>
struct C2RK {  // inverse: complex rows -> real rows (unnormalised; caller's scale carries 1/n)
  static constexpr int GROUP = RowCfg<real, P>::TC;
  using Cfg = RowCfg<real, P, CAPV>;
  using C = cx<real>;
  using Params = RowParams<real>;
  static constexpr int NPHASE = P::S + 2;
  static constexpr int NT = Cfg::NT;
  static constexpr int SMEM = Cfg::SMEM;
  static constexpr int SMEM1 = Cfg::SMEM1;
  static constexpr bool PIPE = Cfg::PIPE;
  static constexpr int MINB = Cfg::MINB;
  B2_HD static unsigned long long blocks(const Params& p) {
    return (unsigned long long)((p.rows + Cfg::RPC - 1) / Cfg::RPC);
  }
  B2_HD static void decode(const Params&, unsigned blk, int& bx, int& by) {
    bx = (int)blk;
    by = 0;
  }
  // work units of the fused pair kernel: rows (RPC per block)
  B2_HD static FuseSide fuse_side(const Params& p, long long planes, long long planes_per_group) {
    return FuseSide{(unsigned)blocks(p), (unsigned)Cfg::RPC, (unsigned)(planes_per_group * (p.rows / planes)), (unsigned)p.rows};
  }

  template <int s>
  B2_HD static void phase(const Params& p, void* smraw, int tid, int bx, int) {
    constexpr int H = Cfg::H, TC = Cfg::TC, CB = Cfg::CB;
    constexpr int M0 = P::template M<0>;
    const int rl = tid / TC, t = tid % TC;
    const long long row = (long long)bx * Cfg::RPC + rl;
    const bool live = row < p.rows;
    C* sm = reinterpret_cast<C*>(smraw) + rl * Cfg::SROW;
    if constexpr (s == 0) {
      // spectrum row -> shared, asynchronously; entries k >= nk are the z zero pad
      // (copy_to_padded axis 2, slab.py:524-525); chunks are the receive blocks of an exchange
      const Side& o = p.cside;
#pragma unroll 4
      for (int k = t; k <= H; k += TC) {
        const bool ok = live && k < p.nk;
        addr_t a = (addr_t)p.tw;
        if (ok) {
          const int pc = (o.nchunk > 1) ? chunk_of(k, o.chunk, o.nchunk) : 0;
          a = (addr_t)o.base[pc] + (addr_t)((row * o.sb[pc] + (k - pc * o.chunk)) * CB);
        }
        async_copy<CB>(sm + (k < H ? swz<M0, Cfg::SW>(k) : H), a, ok);
      }
    } else if constexpr (s == 1) {
      // merge step, in place: G[k] = (X[k] + conj X[H-k]) + i W_n^-k (X[k] - conj X[H-k]); stored swapped
      constexpr int NK = H / 2 + 1;
      constexpr int ROUNDS = (NK + TC - 1) / TC;
#pragma unroll
      for (int rr = 0; rr < ROUNDS; ++rr) {
        const int k = t + rr * TC;
        if (k >= NK) break;
        if (k == 0) {
          const real x0 = sm[swz<M0, Cfg::SW>(0)].x, xh = sm[H].x;  // imaginary parts of DC / Nyquist ignored (C2R)
          sm[swz<M0, Cfg::SW>(0)] = cswap(C{x0 + xh, x0 - xh});
        } else {
          const C a = sm[swz<M0, Cfg::SW>(k)], bq = sm[swz<M0, Cfg::SW>(H - k)];
          const C e = C{a.x + bq.x, a.y - bq.y};            // a + conj b
          const C d = C{a.x - bq.x, a.y + bq.y};            // a - conj b
          const C o2 = cmul(cconj(p.tw[k * p.tws]), d);    // W_n^-k (a - conj b)
          const C io = mul_pi(o2);
          sm[swz<M0, Cfg::SW>(k)] = cswap(cadd(e, io));
          if (k != H - k) sm[swz<M0, Cfg::SW>(H - k)] = cswap(cadd(cconj(e), mul_pi(cconj(o2))));
        }
      }
    } else {
      constexpr int st = s - 2;
      real* dst = reinterpret_cast<real*>(p.rout) + row * p.rpitch;
      auto in = [](int) -> C { return C{0, 0}; };
      auto out = [&](int m, C v) {
        if (!live) return;
        // z = swap(FFT(swap G)) ; x[2m] = Re z, x[2m+1] = Im z
        C z = C{v.y * p.scale, v.x * p.scale};
        if constexpr (STREAM_ST) store_streaming(reinterpret_cast<C*>(dst + 2 * (long long)m), z);
        else *reinterpret_cast<C*>(dst + 2 * (long long)m) = z;
      };
      fft_stage<real, P, st, TC, 1, Cfg::SW, false, (st == P::S - 1)>(t, sm, p.tw, 2 * p.tws, in, out, 0);
    }
  }
};

// C2R, register-staged ("direct") form -- the mirror image of R2CK: each thread loads the spectrum pairs
// (X[k], X[H-k]) it merges straight from HBM into registers (all loads issued before the first use, both
// streams coalesced), merges them there and writes G to shared memory once.  Against C2RK this drops the
// staging copy's shared-memory round trip (one write + two reads per pair) and one barrier; it has no
// asynchronous double buffer and relies on CTA residency for overlap, exactly like R2CK (which reaches
// 0.84 of the HBM figure where C2RK reaches 0.67).  Opt-in: B200FFT_VARIANT=31.
template <class real, class P>
struct C2RDK {
  static constexpr int GROUP = RowCfg<real, P>::TC;
  using Cfg = RowCfg<real, P>;
  using C = cx<real>;
  using Params = RowParams<real>;
  static constexpr int NPHASE = P::S + 1;
  static constexpr int NT = Cfg::NT;
  static constexpr int SMEM = Cfg::SMEM1;
  static constexpr int SMEM1 = Cfg::SMEM1;
  static constexpr bool PIPE = false;
  static constexpr int MINB_ = (227 * 1024) / (SMEM + 1024);
  static constexpr int MINB = MINB_ < 1 ? 1 : (MINB_ > Cfg::MINB_CAP ? Cfg::MINB_CAP : MINB_);
  B2_HD static unsigned long long blocks(const Params& p) {
    return (unsigned long long)((p.rows + Cfg::RPC - 1) / Cfg::RPC);
  }
  B2_HD static void decode(const Params&, unsigned blk, int& bx, int& by) {
    bx = (int)blk;
    by = 0;
  }

  template <int s>
  B2_HD static void phase(const Params& p, void* smraw, int tid, int bx, int) {
    constexpr int H = Cfg::H, TC = Cfg::TC;
    constexpr int M0 = P::template M<0>;
    const int rl = tid / TC, t = tid % TC;
    const long long row = (long long)bx * Cfg::RPC + rl;
    const bool live = row < p.rows;
    C* sm = reinterpret_cast<C*>(smraw) + rl * Cfg::SROW;
    if constexpr (s == 0) {
      const Side& o = p.cside;
      // X[k], or 0 in the z zero pad (copy_to_padded axis 2, slab.py:524-525) / for rows past the end
      auto load = [&](int k) -> C {
        if (!live || k >= p.nk) return C{0, 0};
        const int pc = (o.nchunk > 1) ? chunk_of(k, o.chunk, o.nchunk) : 0;
        return *(reinterpret_cast<const C*>(o.base[pc]) + row * o.sb[pc] + (k - pc * o.chunk));
      };
      constexpr int NK = H / 2 + 1;
      constexpr int ROUNDS = (NK + TC - 1) / TC;
      C a[ROUNDS], bq[ROUNDS];
#pragma unroll
      for (int rr = 0; rr < ROUNDS; ++rr) {
        const int k = t + rr * TC;
        if (k < NK) {
          a[rr] = load(k);
          bq[rr] = load(H - k);
        }
      }
#pragma unroll
      for (int rr = 0; rr < ROUNDS; ++rr) {
        const int k = t + rr * TC;
        if (k >= NK) break;
        if (k == 0) {  // imaginary parts of DC / Nyquist ignored (C2R)
          sm[swz<M0, Cfg::SW>(0)] = cswap(C{a[rr].x + bq[rr].x, a[rr].x - bq[rr].x});
        } else {
          // G[k] = (X[k] + conj X[H-k]) + i W_n^-k (X[k] - conj X[H-k]); stored swapped (inverse by swapping)
          const C e = C{a[rr].x + bq[rr].x, a[rr].y - bq[rr].y};
          const C d = C{a[rr].x - bq[rr].x, a[rr].y + bq[rr].y};
          const C o2 = cmul(cconj(p.tw[k * p.tws]), d);
          sm[swz<M0, Cfg::SW>(k)] = cswap(cadd(e, mul_pi(o2)));
          if (k != H - k) sm[swz<M0, Cfg::SW>(H - k)] = cswap(cadd(cconj(e), mul_pi(cconj(o2))));
        }
      }
    } else {
      constexpr int st = s - 1;
      real* dst = reinterpret_cast<real*>(p.rout) + row * p.rpitch;
      auto in = [](int) -> C { return C{0, 0}; };
      auto out = [&](int m, C v) {
        if (!live) return;
        *reinterpret_cast<C*>(dst + 2 * (long long)m) = C{v.y * p.scale, v.x * p.scale};
      };
      fft_stage<real, P, st, TC, 1, Cfg::SW, false, (st == P::S - 1)>(t, sm, p.tw, 2 * p.tws, in, out, 0);
    }
  }
};

// C2R with the merge step folded into the first stage ("paired" form, B200FFT_VARIANT=34) -- the mirror image
// of R2CPK.  The first-stage butterfly q reads G[q + r*M]; G[k] needs X[k] and X[H-k], and H - (q + r*M) sits in
// butterfly M - q at slot R-1-r: a thread that runs BOTH butterflies loads their 2R spectrum entries straight
// from HBM (coalesced in q), merges the pairs in registers, runs the two butterflies and writes the stage's
// result to shared memory.  Against C2RK: no staging copy, no separate merge pass (two shared-memory round
// trips and two barriers less); against C2RDK one more round trip less.  Plans with a first radix <= 8, at
// least two stages and an even first stride (the 1024- and 1536-point benchmark rows).
template <class real, class P>
struct C2RPK {
  static constexpr int R0 = P::template R<0>;
  static constexpr int M = P::template M<0>;
  static_assert(P::S >= 2 && R0 <= 8 && M % 2 == 0, "paired C2R: first radix <= 8, two or more stages, even first stride");
  using Cfg = RowCfg<real, P, (sizeof(real) == 8 ? 3 : 0)>;
  static constexpr int GROUP = Cfg::TC;
  using C = cx<real>;
  using Params = RowParams<real>;
  static constexpr int NPHASE = P::S;
  static constexpr int NT = Cfg::NT;
  static constexpr int SMEM = Cfg::SMEM1;
  static constexpr int SMEM1 = Cfg::SMEM1;
  static constexpr bool PIPE = false;
  static constexpr int MINB_ = (227 * 1024) / (SMEM + 1024);
  static constexpr int MINB = MINB_ < 1 ? 1 : (MINB_ > Cfg::MINB_CAP ? Cfg::MINB_CAP : MINB_);
  B2_HD static unsigned long long blocks(const Params& p) { return (unsigned long long)((p.rows + Cfg::RPC - 1) / Cfg::RPC); }
  B2_HD static void decode(const Params&, unsigned blk, int& bx, int& by) {
    bx = (int)blk;
    by = 0;
  }

  template <int s>
  B2_HD static void phase(const Params& p, void* smraw, int tid, int bx, int) {
    constexpr int H = Cfg::H, TC = Cfg::TC;
    const int rl = tid / TC, t = tid % TC;
    const long long row = (long long)bx * Cfg::RPC + rl;
    const bool live = row < p.rows;
    C* sm = reinterpret_cast<C*>(smraw) + rl * Cfg::SROW;
    if constexpr (s == 0) {
      const Side& o = p.cside;
      auto load = [&](int k) -> C {  // X[k], 0 in the z zero pad (slab.py:524-525) and for rows past the end
        if (!live || k >= p.nk) return C{0, 0};
        const int pc = (o.nchunk > 1) ? chunk_of(k, o.chunk, o.nchunk) : 0;
        return *(reinterpret_cast<const C*>(o.base[pc]) + row * o.sb[pc] + (k - pc * o.chunk));
      };
      // G[k], G[H-k] (stored swapped: inverse by swapping) from a = X[k], b = X[H-k], 0 < k < H
      auto merge = [&](int k, C a, C bq, C& gk, C& ghk) {
        const C e = C{a.x + bq.x, a.y - bq.y};
        const C d = C{a.x - bq.x, a.y + bq.y};
        const C o2 = cmul(cconj(p.tw[k * p.tws]), d);
        gk = cswap(cadd(e, mul_pi(o2)));
        ghk = cswap(cadd(cconj(e), mul_pi(cconj(o2))));
      };
      // one first-stage butterfly on G values in registers: DFT, twiddles W_H^(q*c), result to shared memory
      auto butterfly = [&](int q, C* v) {
        Dft<R0>::template run<1>(v);
        C w[R0];
        twiddle_powers<R0>(w, p.tw, q * 2 * p.tws);  // W_H^q = W_n^(2q)
        sm[swz<M, Cfg::SW>(q)] = v[0];
#pragma unroll
        for (int c = 1; c < R0; ++c) sm[swz<M, Cfg::SW>(q + c * M)] = cmul(v[c], w[c]);
      };
      constexpr int ITEMS = M / 2 + 1;
      constexpr int ROUNDS = (ITEMS + TC - 1) / TC;
#pragma unroll
      for (int rr = 0; rr < ROUNDS; ++rr) {
        const int it = t + rr * TC;
        if (it >= ITEMS) break;
        const int q1 = it, q2 = (M - it) % M;
        C v1[R0], v2[R0];
#pragma unroll
        for (int r = 0; r < R0; ++r) v1[r] = load(q1 + r * M);
        if (q2 != q1) {
#pragma unroll
          for (int r = 0; r < R0; ++r) v2[r] = load(q2 + r * M);
#pragma unroll
          for (int r = 0; r < R0; ++r) {  // pair (k, H-k) = (q1 + r*M, q2 + (R0-1-r)*M)
            C gk, ghk;
            merge(q1 + r * M, v1[r], v2[R0 - 1 - r], gk, ghk);
            v1[r] = gk;
            v2[R0 - 1 - r] = ghk;
          }
          butterfly(q1, v1);
          butterfly(q2, v2);
        } else if (q1 == 0) {  // entries r*M: mirror (R0-r)*M in the same butterfly; k = 0 pairs with X[H]
          const C xh = load(H);
          v2[0] = cswap(C{v1[0].x + xh.x, v1[0].x - xh.x});  // imaginary parts of DC / Nyquist ignored (C2R)
#pragma unroll
          for (int r = 1; r <= R0 / 2; ++r) {
            C gk, ghk;
            merge(r * M, v1[r], v1[R0 - r], gk, ghk);
            v2[r] = gk;
            if (r != R0 - r) v2[R0 - r] = ghk;
          }
          butterfly(0, v2);
        } else {  // q = M/2: entries M/2 + r*M, mirror at slot R0-1-r of the same butterfly
#pragma unroll
          for (int r = 0; r < (R0 + 1) / 2; ++r) {
            C gk, ghk;
            merge(q1 + r * M, v1[r], v1[R0 - 1 - r], gk, ghk);
            v2[r] = gk;
            if (r != R0 - 1 - r) v2[R0 - 1 - r] = ghk;
          }
          butterfly(q1, v2);
        }
      }
    } else {
      real* dst = reinterpret_cast<real*>(p.rout) + row * p.rpitch;
      auto in = [](int) -> C { return C{0, 0}; };
      auto out = [&](int m, C v) {
        if (!live) return;
        *reinterpret_cast<C*>(dst + 2 * (long long)m) = C{v.y * p.scale, v.x * p.scale};
      };
      fft_stage<real, P, s, TC, 1, Cfg::SW, false, (s == P::S - 1)>(t, sm, p.tw, 2 * p.tws, in, out, 0);
    }
  }
};

// ------------------------------------------------------------------------------------------
// contiguous-row C2C pass (the z pass of slab.C2C, slab.py:538-825): the strided pass's index maps
// (zero pad on load, truncate / fold on store, reversed output index for the inverse, scale) on
// rows that are contiguous in memory (J == 1, unit element stride).  Threads walk along the row, so
// HBM accesses are coalesced; geometry and pipelining are those of the C2R kernel.
// ------------------------------------------------------------------------------------------
template <class real, class P>
struct RowC2CK {
  using Cfg = RowCfg<real, P>;
  static constexpr int GROUP = Cfg::TC;
  using SK = StridedK<real, P>;
  using C = cx<real>;
  using Params = StridedParams<real>;
  static constexpr int NPHASE = P::S + 1;
  static constexpr int NT = Cfg::NT;
  static constexpr int SMEM = Cfg::SMEM;
  static constexpr int SMEM1 = Cfg::SMEM1;
  static constexpr bool PIPE = Cfg::PIPE;
  static constexpr int MINB = Cfg::MINB;
  B2_HD static unsigned long long blocks(const Params& p) {
    return (unsigned long long)((p.B + Cfg::RPC - 1) / Cfg::RPC);
  }
  B2_HD static void decode(const Params&, unsigned blk, int& bx, int& by) {
    bx = (int)blk;
    by = 0;
  }

  template <int s>
  B2_HD static void phase(const Params& p, void* smraw, int tid, int bx, int) {
    constexpr int n = P::N, TC = Cfg::TC, CB = Cfg::CB;
    constexpr int M0 = P::template M<0>;
    const int rl = tid / TC, t = tid % TC;
    const long long row = (long long)bx * Cfg::RPC + rl;
    const bool live = row < p.B;
    C* sm = reinterpret_cast<C*>(smraw) + rl * Cfg::SROW;
    if constexpr (s == 0) {
#pragma unroll 4
      for (int i = t; i < n; i += TC) {
        const addr_t a = live ? SK::in_row(p, row, i, 0) : 0;
        async_copy<CB>(sm + swz<M0, Cfg::SW>(i), a ? a : (addr_t)p.tw, a != 0);
      }
    } else {
      constexpr int st = s - 1;
      auto in = [](int) -> C { return C{0, 0}; };
      auto out = [&](int k, C v) {
        if (!live) return;
        const addr_t a = SK::out_row(p, row, k, 0);
        if (a == 0) return;
        if (p.scale != (real)1) v = cscale(v, p.scale);
        *reinterpret_cast<C*>(a) = v;
      };
      const int fold = (p.inverse && p.fold_mode == 2) ? 3 : p.fold_mode;
      fft_stage<real, P, st, TC, 1, Cfg::SW, false, (st == P::S - 1)>(t, sm, p.tw, p.tws, in, out, fold);
    }
  }
};


// ------------------------------------------------------------------------------------------
// device entry + host launcher
// ------------------------------------------------------------------------------------------
#if defined(__CUDACC__)
// Barrier between two phases.  RB ("row barriers", row kernels only): the TC threads that share a row
// are whole warps and touch nothing but their own row of shared memory, so they synchronise among
// themselves on a named barrier (id 1 + row) instead of stalling the whole CTA -- the rows of a CTA
// then drift apart and overlap one row's butterflies with another's loads and stores.
template <class K, bool RB>
__device__ __forceinline__ void phase_barrier() {
  if constexpr (RB) {
    static_assert(K::GROUP % 32 == 0 && K::NT / K::GROUP <= 15, "row groups must be whole warps, at most 15 per CTA");
    asm volatile("bar.sync %0, %1;\n" ::"r"((int)threadIdx.x / K::GROUP + 1), "n"(K::GROUP) : "memory");
  } else {
    __syncthreads();
  }
}

template <class K, int s, bool RB = false>
__device__ __forceinline__ void run_phases(const typename K::Params& p, void* sm, int bx, int by) {
  K::template phase<s>(p, sm, (int)threadIdx.x, bx, by);
  if constexpr (s == 0) async_copy_wait();  // phase 0 of the row kernels only issues its loads
  if constexpr (s + 1 < K::NPHASE) {
    phase_barrier<K, RB>();
    run_phases<K, s + 1, RB>(p, sm, bx, by);
  }
}

template <class K, bool RB = false>
__global__ void __launch_bounds__(K::NT, K::MINB) fft_kernel(const __grid_constant__ typename K::Params p) {
  extern __shared__ __align__(128) unsigned char smraw[];
  if constexpr (K::PIPE) {
    // persistent, double-buffered: phase 0 (asynchronous loads) of the CTA's next work item is
    // issued before the current item's compute phases; launched with one CTA per resident slot
    const unsigned nblk = (unsigned)K::blocks(p);
    unsigned g = blockIdx.x;
    int buf = 0, bx, by;
    if (g < nblk) {
      K::decode(p, g, bx, by);
      K::template phase<0>(p, smraw, (int)threadIdx.x, bx, by);
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    for (; g < nblk; g += gridDim.x) {
      const unsigned gn = g + gridDim.x;
      if (gn < nblk) {
        K::decode(p, gn, bx, by);
        K::template phase<0>(p, smraw + (buf ^ 1) * K::SMEM1, (int)threadIdx.x, bx, by);
      }
      asm volatile("cp.async.commit_group;\n" ::: "memory");
      asm volatile("cp.async.wait_group 1;\n" ::: "memory");
      phase_barrier<K, RB>();
      K::decode(p, g, bx, by);
      run_phases<K, 1, RB>(p, smraw + buf * K::SMEM1, bx, by);
      phase_barrier<K, RB>();  // every read of this buffer is done before the next iteration refills it
      buf ^= 1;
    }
  } else {
    int bx, by;
    K::decode(p, blockIdx.x, bx, by);
    run_phases<K, 0, RB>(p, smraw, bx, by);
  }
}

// Cluster kernel: the two CTAs of a cluster share one column tile (ClusterStridedK).  Cluster-wide
// barriers (which also order distributed-shared-memory accesses) surround the cross stage; every
// thread of both CTAs reaches every barrier (no early exits), and no CTA touches its partner's
// shared memory after the second one, so either may retire first.
template <class K, int s>
__device__ __forceinline__ void run_cluster_phases(const typename K::Params& p, void* sm, void* peer, int rank, int bx, int by) {
  K::template phase<s>(p, sm, peer, (int)threadIdx.x, rank, bx, by);
  if constexpr (s + 1 < K::NPHASE) {
    if constexpr (s + 1 == K::SYNC_BEFORE || s == K::SYNC_BEFORE) {
      asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
      asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    } else {
      __syncthreads();
    }
    run_cluster_phases<K, s + 1>(p, sm, peer, rank, bx, by);
  }
}


// phases of pass K inside a CTA of NTL >= K::NT threads: the surplus threads only keep the barriers
template <class K, int s, int NTL>
__device__ __forceinline__ void run_phases_in(const typename K::Params& p, void* sm, int bx, int by) {
  if (NTL == K::NT || (int)threadIdx.x < K::NT) {
    K::template phase<s>(p, sm, (int)threadIdx.x, bx, by);
    if constexpr (s == 0) async_copy_wait();
  }
  if constexpr (s + 1 < K::NPHASE) {
    __syncthreads();
    run_phases_in<K, s + 1, NTL>(p, sm, bx, by);
  }
}

// one block of pass K; not inlined, so that each pass keeps the register allocation of its own kernel
// instead of the union of both (the 1536-point pair spilled 384 bytes when inlined)
template <class K, int NTL>
__device__ __noinline__ void run_block_in(const typename K::Params& p, void* sm, unsigned blk) {
  int bx, by;
  K::decode(p, blk, bx, by);
  run_phases_in<K, 0, NTL>(p, sm, bx, by);
}

template <class KA, class KB>
struct FusePair {
  static constexpr int NT = KA::NT > KB::NT ? KA::NT : KB::NT;
  static constexpr int SMEM = KA::SMEM1 > KB::SMEM1 ? KA::SMEM1 : KB::SMEM1;  // one work item at a time: no double buffer
  static constexpr int MINB_ = (227 * 1024) / (SMEM + 1024);
  static constexpr int MINB_K = KA::MINB < KB::MINB ? KA::MINB : KB::MINB;
  static constexpr int MINB = MINB_ < 1 ? 1 : (MINB_ < MINB_K ? MINB_ : MINB_K);
};

template <class KA, class KB>
__global__ void __launch_bounds__(FusePair<KA, KB>::NT, FusePair<KA, KB>::MINB)
    fused_pair_kernel(const __grid_constant__ typename KA::Params pa, const __grid_constant__ typename KB::Params pb,
                      const __grid_constant__ FuseCtl c) {
  extern __shared__ __align__(128) unsigned char smraw[];
  __shared__ unsigned s_blk;  // block to run: bit 31 = pass B, 0xffffffff = queue exhausted
  const unsigned total = c.a.n + c.b.n;
  for (;;) {
    __syncthreads();  // the previous item is done with shared memory and with s_blk
    if (threadIdx.x == 0) {  // one thread takes the next item, decodes it and, for pass B, waits for its groups
      const unsigned i = atomicAdd(c.ctr, 1u);
      unsigned v = 0xffffffffu;
      if (i < total) {
        bool isB;
        unsigned blk, g1, g2, u1, u2;
        fuse_decode(c, i, isB, blk);
        if (isB) {
          c.b.groups(blk, g1, g2, u1, u2);
          // (bounded: a protocol error must end as a launch failure, never as a GPU that spins forever --
          // 2^26 polls of >= 200 ns are more than ten seconds, a healthy wait is microseconds)
          unsigned polls = 0;
          for (unsigned g = g1; g <= g2; ++g)
            while (*reinterpret_cast<volatile unsigned*>(c.done + g) < c.a.need(g)) {
              __nanosleep(200);
              if (++polls > (1u << 26)) __trap();
            }
          __threadfence();
        }
        v = blk | (isB ? 0x80000000u : 0u);
      }
      s_blk = v;
    }
    __syncthreads();
    const unsigned v = s_blk;
    if (v == 0xffffffffu) break;
    const unsigned blk = v & 0x7fffffffu;
    if (!(v & 0x80000000u)) {
      run_block_in<KA, FusePair<KA, KB>::NT>(pa, smraw, blk);
      __threadfence();  // this thread's stores are visible device-wide ...
      __syncthreads();  // ... for every thread of the block, before its units are counted
      if (threadIdx.x == 0) {
        unsigned g1, g2, u1, u2;
        c.a.groups(blk, g1, g2, u1, u2);
        atomicAdd(c.done + g1, u1);
        if (g2 != g1) atomicAdd(c.done + g2, u2);
      }
    } else {
      run_block_in<KB, FusePair<KA, KB>::NT>(pb, smraw, blk);
    }
  }
}

template <class K>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(K::NT, K::MINB)
    fft_cluster_kernel(const __grid_constant__ typename K::Params p) {
  extern __shared__ __align__(128) unsigned char smraw[];
  unsigned rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(rank));
  // generic address of the partner CTA's copy of smraw (distributed shared memory window)
  void* peer;
  asm volatile("mapa.u64 %0, %1, %2;\n" : "=l"(peer) : "l"(smraw), "r"(rank ^ 1u));
  int bx, by;
  K::decode(p, blockIdx.x, bx, by);
  run_cluster_phases<K, 0>(p, smraw, peer, (int)rank, bx, by);
}
#endif

}  // namespace b200fft
