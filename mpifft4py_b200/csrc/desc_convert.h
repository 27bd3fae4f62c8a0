// C-ABI descriptor -> kernel parameter conversion and twiddle generation (host side, shared by
// libb200fft.so and the CPU emulator in tests/emu).
#pragma once
#include <cmath>
#include <cstring>
#include <vector>
#include "../../include/b200fft.h"
#include "fft_kernels.cuh"

namespace b200fft {

static_assert(MAXP == B200FFT_MAXP, "peer table sizes must agree");
static_assert(sizeof(Side) == sizeof(b200fft_side_t), "Side mirrors b200fft_side_t");
static_assert(sizeof(Mask) == sizeof(b200fft_mask_t), "Mask mirrors b200fft_mask_t");

// W_len^j = exp(-2*pi*i*j/len), j in [0, len): long double evaluation with octant reduction so
// that every entry is the correctly rounded value of the exact root (1e-12 parity budget).
template <class real>
std::vector<cx<real>> make_twiddles(int len) {
  std::vector<cx<real>> t((size_t)len);
  const long double pi = 3.14159265358979323846264338327950288L;
  for (int j = 0; j < len; ++j) {
    // theta = 2*pi*j/len = oct*(pi/4) + phi, phi in [0, pi/4): evaluate cos/sin only on
    // [0, pi/4] and rotate by exact quarter turns.
    const long long j8 = 8LL * j;
    const int oct = (int)(j8 / len);
    const long long r = j8 - (long long)oct * len;
    long double c, s;
    int q;
    if (oct % 2 == 0) {
      const long double phi = 2.0L * pi * (long double)r / (8.0L * (long double)len);
      c = cosl(phi);
      s = sinl(phi);
      q = oct / 2;
    } else {
      const long double beta = 2.0L * pi * (long double)(len - r) / (8.0L * (long double)len);
      c = cosl(beta);
      s = -sinl(beta);
      q = ((oct + 1) / 2) % 4;
    }
    long double C, S;
    switch (q) {
      case 0: C = c; S = s; break;
      case 1: C = -s; S = c; break;
      case 2: C = -c; S = -s; break;
      default: C = s; S = -c; break;
    }
    t[(size_t)j] = cx<real>{(real)C, (real)(-S)};  // exp(-i theta)
  }
  return t;
}

inline void convert_side(Side& o, const b200fft_side_t& s) { std::memcpy(&o, &s, sizeof(Side)); }

template <class real>
StridedParams<real> convert_strided(const b200fft_strided_desc_t& d, const cx<real>* tw, int tws) {
  StridedParams<real> p;
  std::memset(&p, 0, sizeof(p));
  convert_side(p.in, d.in);
  convert_side(p.out, d.out);
  p.B = d.B;
  p.J = d.J;
  p.n = d.n;
  p.inverse = d.inverse;
  p.fold_mode = d.fold_mode;
  p.scale = (real)d.scale;
  std::memcpy(&p.mask, &d.mask, sizeof(Mask));
  if (p.mask.jdiv <= 0) p.mask.jdiv = 1;
  p.tw = tw;
  p.tws = tws;
  return p;
}

template <class real>
RowParams<real> convert_rows(const b200fft_rows_desc_t& d, const cx<real>* tw, int tws, bool forward) {
  RowParams<real> p;
  std::memset(&p, 0, sizeof(p));
  convert_side(p.cside, d.cside);
  p.rin = forward ? d.real_base : nullptr;
  p.rout = forward ? nullptr : d.real_base;
  p.rpitch = d.rpitch;
  p.rows = d.rows;
  p.n = d.n;
  p.nk = d.nk;
  p.scale = (real)d.scale;
  p.tw = tw;
  p.tws = tws;
  p.rm_period = d.rm_period;
  p.rm_block = d.rm_block;
  p.rm_planes = d.rm_planes;
  return p;
}

// argument validation shared by both back ends; returns nullptr if fine, else a message
inline const char* check_side(const b200fft_side_t& s, int n, bool strided) {
  if (s.nchunk < 1 || s.nchunk > B200FFT_MAXP) return "side.nchunk out of range";
  if (s.nphys < 1 || s.nphys > n) return "side.nphys out of range";
  if (s.nchunk > 1 && s.chunk < 1) return "side.chunk must be positive";
  if (s.nchunk > 1 && (long long)s.chunk * (s.nchunk - 1) >= s.nphys) return "side chunks exceed extent";
  for (int p = 0; p < s.nchunk; ++p)
    if (!s.base[p]) return "side.base is null";
  (void)strided;
  return nullptr;
}

inline const char* check_strided(const b200fft_strided_desc_t& d) {
  if (d.precision != B200FFT_SINGLE && d.precision != B200FFT_DOUBLE) return "bad precision";
  if (d.n < 2 || d.B < 0 || d.J < 0) return "bad sizes";
  if (const char* e = check_side(d.in, d.n, true)) return e;
  if (const char* e = check_side(d.out, d.n, true)) return e;
  if (d.fold_mode < 0 || d.fold_mode > 2) return "bad fold_mode";
  if (d.out.nphys < d.n) {
    // the fold needs modes +N/2 and -N/2 in one last-stage butterfly: n = 3N/2 exactly
    if (d.fold_mode != 0 && 2 * d.n != 3 * d.out.nphys) return "fold needs n == 1.5 * nphys";
    if (d.out.nphys % 2) return "odd truncated extent";
  }
  if (d.in.nphys < d.n && d.in.nphys % 2) return "odd padded extent";
  if (d.cross_n < 0 || (d.cross_n > 0 && (d.cross_div < 1 || d.cross_n % d.n || (long long)(d.J / d.cross_div + (d.J % d.cross_div != 0)) * d.n > d.cross_n ||
                                          d.in.nphys != d.n || d.out.nphys != d.n || d.fold_mode != 0)))
    return "bad cross twiddle (cross_n = n * sub-columns, no pad / truncation in a four-step launch)";
  return nullptr;
}

// a strided pass whose "columns" are single elements of contiguous rows: served by the row C2C kernel
inline bool contiguous_rows(const b200fft_strided_desc_t& d) {
  return d.J == 1 && d.in.nchunk == 1 && d.out.nchunk == 1 && d.in.si[0] == 1 && d.out.si[0] == 1 && !d.mask.on && d.cross_n == 0;
}

inline const char* check_rows(const b200fft_rows_desc_t& d) {
  if (d.precision != B200FFT_SINGLE && d.precision != B200FFT_DOUBLE) return "bad precision";
  if (d.n < 4 || d.n % 2 || d.rows < 0) return "bad sizes";
  if (d.nk < 1 || d.nk > d.n / 2 + 1) return "bad nk";
  if (d.rpitch % 2) return "real row pitch must be even";
  if (!d.real_base) return "real_base is null";
  b200fft_side_t s = d.cside;
  if (s.nchunk < 1 || s.nchunk > B200FFT_MAXP) return "cside.nchunk out of range";
  for (int p = 0; p < s.nchunk; ++p)
    if (!s.base[p]) return "cside.base is null";
  if (d.rm_block < 0 || (d.rm_block > 0 && (d.rm_period < 1 || d.rm_planes < 1 || d.rm_period % d.rm_block || d.rows > d.rm_period * d.rm_planes)))
    return "bad row map (rm_block must divide rm_period; rows <= rm_period * rm_planes)";
  return nullptr;
}

}  // namespace b200fft
