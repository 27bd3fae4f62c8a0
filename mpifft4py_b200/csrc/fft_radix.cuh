// Register-resident DFT butterflies (radix 2,3,4 and composites 6,8,12,16,24) for sm_100a.
//
// Everything here is __host__ __device__ so that the very same index math and arithmetic can
// be executed by the CPU emulator used in the `-m "not gpu"` tests (tests/emu); the product
// path only ever runs the __global__ kernels in fft_kernels.cuh.
//
// Convention: forward DFT, X[k] = sum_j x[j] exp(-2*pi*i*j*k/R) (numpy.fft.fft, which the
// reference's serialFFT/numpy_fft.py:25-30 wraps).  Inverse transforms are obtained by swapping
// re/im on load and store (IFFT(x) = swap(FFT(swap(x)))), so only forward butterflies exist.
#pragma once

#if defined(__CUDACC__)
#define B2_HD __host__ __device__ __forceinline__
#else
#define B2_HD inline
#endif

namespace b200fft {

template <class R>
struct alignas(2 * sizeof(R)) cx {
  R x, y;
};

template <class R> B2_HD cx<R> cadd(cx<R> a, cx<R> b) { return cx<R>{a.x + b.x, a.y + b.y}; }
template <class R> B2_HD cx<R> csub(cx<R> a, cx<R> b) { return cx<R>{a.x - b.x, a.y - b.y}; }
template <class R> B2_HD cx<R> cmul(cx<R> a, cx<R> b) {
  return cx<R>{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
template <class R> B2_HD cx<R> cconj(cx<R> a) { return cx<R>{a.x, -a.y}; }
template <class R> B2_HD cx<R> cswap(cx<R> a) { return cx<R>{a.y, a.x}; }
template <class R> B2_HD cx<R> cscale(cx<R> a, R s) { return cx<R>{a.x * s, a.y * s}; }
// multiply by -i and +i
template <class R> B2_HD cx<R> mul_mi(cx<R> a) { return cx<R>{a.y, -a.x}; }
template <class R> B2_HD cx<R> mul_pi(cx<R> a) { return cx<R>{-a.y, a.x}; }

// cos(2*pi*m/48), m = 0..12, to 25 digits (48 = lcm(16, 24) covers every internal twiddle).
B2_HD constexpr double cos48_q(int m) {
  return m == 0 ? 1.0
       : m == 1 ? 0.9914448613738104111445575
       : m == 2 ? 0.9659258262890682867497432
       : m == 3 ? 0.9238795325112867561281832
       : m == 4 ? 0.8660254037844386467637232
       : m == 5 ? 0.7933533402912351645797770
       : m == 6 ? 0.7071067811865475244008444
       : m == 7 ? 0.6087614290087206394160975
       : m == 8 ? 0.5
       : m == 9 ? 0.3826834323650897717284600
       : m == 10 ? 0.2588190451025207623488988
       : m == 11 ? 0.1305261922200515915484062
       : 0.0;
}
// cos / sin of 2*pi*m/48 for any integer m >= 0
B2_HD constexpr double cos48(int m) {
  m %= 48;
  return m <= 12 ? cos48_q(m) : m <= 24 ? -cos48_q(24 - m) : m <= 36 ? -cos48_q(m - 24) : cos48_q(48 - m);
}
B2_HD constexpr double sin48(int m) { return cos48(m + 36); }  // sin(a) = cos(a - 90deg) = cos(a + 270deg)

// v *= exp(-2*pi*i*M48/48), with the cheap forms for multiples of 45 degrees.
template <int M48, class R>
B2_HD cx<R> mul_w48(cx<R> a) {
  constexpr int m = ((M48 % 48) + 48) % 48;
  if constexpr (m == 0) {
    return a;
  } else if constexpr (m == 12) {  // -i
    return mul_mi(a);
  } else if constexpr (m == 24) {
    return cx<R>{-a.x, -a.y};
  } else if constexpr (m == 36) {  // +i
    return mul_pi(a);
  } else if constexpr (m == 6) {  // (1 - i)/sqrt2
    constexpr R h = (R)0.7071067811865475244008444;
    return cx<R>{(a.x + a.y) * h, (a.y - a.x) * h};
  } else if constexpr (m == 18) {  // (-1 - i)/sqrt2
    constexpr R h = (R)0.7071067811865475244008444;
    return cx<R>{(a.y - a.x) * h, -(a.x + a.y) * h};
  } else if constexpr (m == 30) {  // (-1 + i)/sqrt2
    constexpr R h = (R)0.7071067811865475244008444;
    return cx<R>{-(a.x + a.y) * h, (a.x - a.y) * h};
  } else if constexpr (m == 42) {  // (1 + i)/sqrt2
    constexpr R h = (R)0.7071067811865475244008444;
    return cx<R>{(a.x - a.y) * h, (a.x + a.y) * h};
  } else {
    constexpr R c = (R)cos48(m);
    constexpr R s = (R)(-sin48(m));  // exp(-i t) = cos t - i sin t
    return cx<R>{a.x * c - a.y * s, a.x * s + a.y * c};
  }
}

// ---- prime butterflies on strided register arrays: elements v[0], v[ST], v[2ST], ... -------
template <int R> struct Dft;

template <> struct Dft<1> {
  template <int ST, class T> B2_HD static void run(cx<T>*) {}
};

template <> struct Dft<2> {
  template <int ST, class T> B2_HD static void run(cx<T>* v) {
    cx<T> a = v[0], b = v[ST];
    v[0] = cadd(a, b);
    v[ST] = csub(a, b);
  }
};

template <> struct Dft<3> {
  template <int ST, class T> B2_HD static void run(cx<T>* v) {
    constexpr T s = (T)0.8660254037844386467637232;
    cx<T> a = v[0], b = v[ST], c = v[2 * ST];
    cx<T> t1 = cadd(b, c);
    cx<T> t2 = cx<T>{a.x - (T)0.5 * t1.x, a.y - (T)0.5 * t1.y};
    cx<T> d = csub(b, c);
    cx<T> t3 = cx<T>{s * d.x, s * d.y};
    v[0] = cadd(a, t1);
    v[ST] = cadd(t2, mul_mi(t3));      // t2 - i*t3
    v[2 * ST] = cadd(t2, mul_pi(t3));  // t2 + i*t3
  }
};

template <> struct Dft<4> {
  template <int ST, class T> B2_HD static void run(cx<T>* v) {
    cx<T> a = v[0], b = v[ST], c = v[2 * ST], d = v[3 * ST];
    cx<T> s0 = cadd(a, c), d0 = csub(a, c), s1 = cadd(b, d), d1 = csub(b, d);
    v[0] = cadd(s0, s1);
    v[2 * ST] = csub(s0, s1);
    v[ST] = cadd(d0, mul_mi(d1));      // d0 - i*d1
    v[3 * ST] = cadd(d0, mul_pi(d1));  // d0 + i*d1
  }
};

// ---- Cooley-Tukey composite R = R1*R2 entirely in registers ----------------------------------
// input index i = R2*i1 + i2, output index k = k1 + R1*k2 (natural order in and out):
//   step 1: for each i2, DFT_R1 over i1 (stride R2*ST)         -> y[k1][i2] at v[(R2*k1+i2)*ST]
//   step 2: y[k1][i2] *= W_R^(i2*k1)
//   step 3: for each k1, DFT_R2 over i2 (stride ST)            -> X[k1+R1*k2] at v[(R2*k1+k2)*ST]
//   step 4: transpose to natural order.
template <int R1, int R2, int K1, int I2> struct CompTw {
  template <int ST, class T> B2_HD static void run(cx<T>* v) {
    constexpr int R = R1 * R2;
    v[(R2 * K1 + I2) * ST] = mul_w48<(I2 * K1) * (48 / R)>(v[(R2 * K1 + I2) * ST]);
    if constexpr (I2 + 1 < R2) CompTw<R1, R2, K1, I2 + 1>::template run<ST>(v);
    else if constexpr (K1 + 1 < R1) CompTw<R1, R2, K1 + 1, 0>::template run<ST>(v);
  }
};

template <int R1, int R2> struct Composite {
  template <int ST, class T> B2_HD static void run(cx<T>* v) {
    constexpr int R = R1 * R2;
    static_assert(48 % R == 0, "internal twiddles come from the 48-th roots table");
#pragma unroll
    for (int i2 = 0; i2 < R2; ++i2) Dft<R1>::template run<R2 * ST>(v + i2 * ST);
    CompTw<R1, R2, 0, 0>::template run<ST>(v);
#pragma unroll
    for (int k1 = 0; k1 < R1; ++k1) Dft<R2>::template run<ST>(v + R2 * k1 * ST);
    cx<T> t[R];
#pragma unroll
    for (int k1 = 0; k1 < R1; ++k1)
#pragma unroll
      for (int k2 = 0; k2 < R2; ++k2) t[k1 + R1 * k2] = v[(R2 * k1 + k2) * ST];
#pragma unroll
    for (int k = 0; k < R; ++k) v[k * ST] = t[k];
  }
};

template <> struct Dft<6> {
  template <int ST, class T> B2_HD static void run(cx<T>* v) { Composite<2, 3>::run<ST>(v); }
};
template <> struct Dft<8> {
  template <int ST, class T> B2_HD static void run(cx<T>* v) { Composite<2, 4>::run<ST>(v); }
};
template <> struct Dft<12> {
  template <int ST, class T> B2_HD static void run(cx<T>* v) { Composite<4, 3>::run<ST>(v); }
};
template <> struct Dft<16> {
  template <int ST, class T> B2_HD static void run(cx<T>* v) { Composite<4, 4>::run<ST>(v); }
};
template <> struct Dft<24> {
  template <int ST, class T> B2_HD static void run(cx<T>* v) { Composite<8, 3>::run<ST>(v); }
};

}  // namespace b200fft
