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
};

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
template <int BYTES>
B2_HD void async_copy(void* dst_smem, addr_t src, bool valid) {
#if defined(__CUDA_ARCH__)
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  const int nb = valid ? BYTES : 0;
  // (an L2 prefetch size on these copies -- .L2::128B / .L2::256B -- changes nothing, near or far:
  // profiles/r02_single/ab_l2hint.txt)
  if constexpr (BYTES == 16)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(src), "r"(nb) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(src), "r"(nb) : "memory");
#else
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
  const cx<real>* tw2;  // cross twiddles W_N^j of a four-step transform (b200fft_strided_desc_t::cross_n), or null
  int tw2_div;          // sub-column of a column: column / tw2_div
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
  // (96 threads per column for the 1536-point double-precision columns -- one full round of first-stage radix-16
  // butterflies, 24 warps per SM instead of 16 at 80 registers -- measured 3-6 % SLOWER than 64, round 2.)
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

template <class real, class P, int MB = 0, int RB = 0>
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

    // row addresses are taken at column j0; every thread adds its own column offset c
    const long long jin = j0, jout = j0;
    const long long cin = c, cout = c;
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
        async_copy<CB>(sm + swz<M0, Cfg::SW>(i) * T, ok ? a + (addr_t)(cin * CB) : fallback, ok);
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
        if (p.tw2 != nullptr) {  // four-step: W_N^(x2 * k1), k1 = the frequency this slot really holds
          const int k1 = p.inverse ? (k == 0 ? 0 : n - k) : k;
          const C w = p.tw2[(long long)((j0 + c) / p.tw2_div) * k1];
          v = cmul(v, p.inverse ? cconj(w) : w);
        }
        *reinterpret_cast<C*>(a + (addr_t)(cout * CB)) = v;
      };
      // reversed output index: the slot that survives a "keep -N/2" truncation is the other one
      const int fold = (p.inverse && p.fold_mode == 2) ? 3 : p.fold_mode;
      fft_stage<real, P, st, Cfg::TC, T, Cfg::SW, false, (st == P::S - 1)>(t, sm, p.tw, p.tws, in, out, fold);
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
  long long rm_period, rm_block, rm_planes;  // complex-side row permutation (b200fft_rows_desc_t), rm_block == 0: none
};

// complex-side row of real row r: [x][y] kept as [y block][x][y in block] (include/b200fft.h)
template <class real>
B2_HD long long crow_of(const RowParams<real>& p, long long r) {
  if (p.rm_block <= 0) return r;
  const long long x = r / p.rm_period, y = r - x * p.rm_period;
  const long long yb = y / p.rm_block;
  return (yb * p.rm_planes + x) * p.rm_block + (y - yb * p.rm_block);
}

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
  static constexpr int SW = 128 / CB;
  static constexpr int SMEM1 = RPC * SROW * CB;
  static constexpr bool PIPE = 2 * SMEM1 + 1024 <= 227 * 1024;
  static constexpr int SMEM = PIPE ? 2 * SMEM1 : SMEM1;
  static constexpr int MINB_ = (227 * 1024) / (SMEM + 1024);
  // register budget: double-precision radix-16 butterflies need ~128 registers, radix-12 ~96
  static constexpr int MINB_CAP = CAPV > 0 ? CAPV : (P::RMAX >= 16 && CB == 16) ? 2 : (P::RMAX >= 12 ? 3 : 4);
  static constexpr int MINB = MINB_ < 1 ? 1 : (MINB_ > MINB_CAP ? MINB_CAP : MINB_);
};

template <class real, class P, int CAPV = 0>
struct R2CK {  // forward: real rows -> complex rows
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
        return live ? src[i] : C{0, 0};
      };
      auto out = [](int, C) {};
      // every stage writes shared memory (the split step needs all of F)
      fft_stage<real, P, s, TC, 1, Cfg::SW, (s == 0), false>(t, sm, p.tw, 2 * p.tws, in, out, 0);
    } else {
      // split step: X[k] = E[k] + W_n^k O[k],  X[H-k] = conj(E[k] - W_n^k O[k])
      if (!live) return;
      const Side& o = p.cside;
      const long long crow = crow_of(p, row);
      auto store = [&](int k, C v) {
        if (k >= p.nk) return;  // z truncation of the 3/2-rule: copy_from_padded axis 2 (slab.py:535)
        const int pc = (o.nchunk > 1) ? chunk_of(k, o.chunk, o.nchunk) : 0;
        C* ptr = reinterpret_cast<C*>(o.base[pc]) + crow * o.sb[pc] + (k - pc * o.chunk);
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

template <class real, class P, int CAPV = 0>
struct C2RK {  // inverse: complex rows -> real rows (unnormalised; caller's scale carries 1/n)
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
      const long long crow = live ? crow_of(p, row) : 0;
#pragma unroll 4
      for (int k = t; k <= H; k += TC) {
        const bool ok = live && k < p.nk;
        addr_t a = (addr_t)p.tw;
        if (ok) {
          const int pc = (o.nchunk > 1) ? chunk_of(k, o.chunk, o.nchunk) : 0;
          a = (addr_t)o.base[pc] + (addr_t)((crow * o.sb[pc] + (k - pc * o.chunk)) * CB);
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
        *reinterpret_cast<C*>(dst + 2 * (long long)m) = z;
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
// 0.84 of the HBM figure where C2RK reaches 0.67).  Measured 3.31 ms against 3.93 ms at 1024^3 double
// (profiles/r02_single/ab_single.txt): the default for rows of 512 ... 3072 reals (k_rows.inc).
template <class real, class P, int CAPV = 0>
struct C2RDK {
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
      const long long crow = live ? crow_of(p, row) : 0;
      // X[k], or 0 in the z zero pad (copy_to_padded axis 2, slab.py:524-525) / for rows past the end
      auto load = [&](int k) -> C {
        if (!live || k >= p.nk) return C{0, 0};
        const int pc = (o.nchunk > 1) ? chunk_of(k, o.chunk, o.nchunk) : 0;
        return *(reinterpret_cast<const C*>(o.base[pc]) + crow * o.sb[pc] + (k - pc * o.chunk));
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

// ------------------------------------------------------------------------------------------
// contiguous-row C2C pass (the z pass of slab.C2C, slab.py:538-825): the strided pass's index maps
// (zero pad on load, truncate / fold on store, reversed output index for the inverse, scale) on
// rows that are contiguous in memory (J == 1, unit element stride).  Threads walk along the row, so
// HBM accesses are coalesced; geometry and pipelining are those of the C2R kernel.
// ------------------------------------------------------------------------------------------
template <class real, class P>
struct RowC2CK {
  using Cfg = RowCfg<real, P>;
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
template <class K, int s>
__device__ __forceinline__ void run_phases(const typename K::Params& p, void* sm, int bx, int by) {
  K::template phase<s>(p, sm, (int)threadIdx.x, bx, by);
  if constexpr (s == 0) async_copy_wait();  // phase 0 of the row kernels only issues its loads
  if constexpr (s + 1 < K::NPHASE) {
    __syncthreads();
    run_phases<K, s + 1>(p, sm, bx, by);
  }
}

template <class K>
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
      __syncthreads();
      K::decode(p, g, bx, by);
      run_phases<K, 1>(p, smraw + buf * K::SMEM1, bx, by);
      __syncthreads();  // every read of this buffer is done before the next iteration refills it
      buf ^= 1;
    }
  } else {
    int bx, by;
    K::decode(p, blockIdx.x, bx, by);
    run_phases<K, 0>(p, smraw, bx, by);
  }
}
#endif

}  // namespace b200fft
