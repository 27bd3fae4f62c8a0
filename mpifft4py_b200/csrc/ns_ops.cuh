// Pointwise operations of the pseudo-spectral Navier-Stokes right-hand side -- the caller of the transforms
// (/root/reference/demo/spectral_dns_solver.py:53-77, 87-98): curl in spectral space,
// cross product in physical space, pressure projection + viscous term + the Runge-Kutta updates.
//
// The reference keeps three dense wavenumber meshes K[i] of the complex shape and K2, K_over_K2 next to them
// (demo :44-47) and runs a dozen numpy passes per stage.  Here a point's wavenumbers come from the three 1D
// vectors (kx[n0], ky[n1], kz[n2]) through its linear index, so no mesh is ever read from HBM, and each of the
// three operations is ONE pass over its operands.
//
// Bodies are __host__ __device__ and free of CUDA builtins: tests/emu/host_shim.cpp runs the same functions in
// plain loops (never part of the product path).
#pragma once
#include "fft_radix.cuh"

namespace b200fft {

template <class real>
struct NsMesh {
  long long n0, n1, n2;  // local complex shape, C order
  const real *kx, *ky, *kz;
};

// wavenumbers of linear index i of an (n0, n1, n2) array
template <class real>
B2_HD void ns_wavenumbers(const NsMesh<real>& m, long long i, real& kx, real& ky, real& kz) {
  const long long k = i % m.n2, r = i / m.n2;
  kx = m.kx[r / m.n1];
  ky = m.ky[r % m.n1];
  kz = m.kz[k];
}

// curl_hat = i K x u_hat   (demo :60-64: z[0] = 1j*(K[1]*x[2] - K[2]*x[1]), ...); arrays are [3][n]
template <class real>
B2_HD void ns_curl_point(const NsMesh<real>& m, long long n, long long i, const cx<real>* u, cx<real>* c) {
  real kx, ky, kz;
  ns_wavenumbers(m, i, kx, ky, kz);
  const cx<real> u0 = u[i], u1 = u[n + i], u2 = u[2 * n + i];
  // 1j * (a + ib) = -b + ia
  const cx<real> t0{ky * u2.x - kz * u1.x, ky * u2.y - kz * u1.y};
  const cx<real> t1{kz * u0.x - kx * u2.x, kz * u0.y - kx * u2.y};
  const cx<real> t2{kx * u1.x - ky * u0.x, kx * u1.y - ky * u0.y};
  c[i] = cx<real>{-t0.y, t0.x};
  c[n + i] = cx<real>{-t1.y, t1.x};
  c[2 * n + i] = cx<real>{-t2.y, t2.x};
}

// w = a x b in physical space (demo :53-58); arrays are [3][n] reals
template <class real>
B2_HD void ns_cross_point(long long n, long long i, const real* a, const real* b, real* w) {
  const real a0 = a[i], a1 = a[n + i], a2 = a[2 * n + i];
  const real b0 = b[i], b1 = b[n + i], b2 = b[2 * n + i];
  w[i] = a1 * b2 - a2 * b1;
  w[n + i] = a2 * b0 - a0 * b2;
  w[2 * n + i] = a0 * b1 - a1 * b0;
}

// Right-hand side from the transformed cross product `du` (demo :72-76):
//     P = sum_i du_i K_i / K2 (K2 == 0 -> 1);   rhs_i = du_i - P K_i - nu K2 u_i
// followed by the Runge-Kutta bookkeeping of the stage (demo :91-97), fused so that no operand makes a second trip:
//     u1_i += a_dt * rhs_i;   u_i = (last ? u1_i : u0_i + b_dt * rhs_i)
// With u0 == nullptr only rhs is formed and written back to du (the plain compute_rhs).
template <class real>
B2_HD void ns_rhs_point(const NsMesh<real>& m, long long n, long long i, real nu, cx<real>* du, cx<real>* u, const cx<real>* u0,
                        cx<real>* u1, real a_dt, real b_dt, int last) {
  real k[3];
  ns_wavenumbers(m, i, k[0], k[1], k[2]);
  const real k2 = k[0] * k[0] + k[1] * k[1] + k[2] * k[2];
  const real inv = (real)1 / (k2 == (real)0 ? (real)1 : k2);
  cx<real> d[3], uu[3];
  for (int c = 0; c < 3; ++c) {
    d[c] = du[c * n + i];
    uu[c] = u[c * n + i];
  }
  cx<real> p{(d[0].x * k[0] + d[1].x * k[1] + d[2].x * k[2]) * inv, (d[0].y * k[0] + d[1].y * k[1] + d[2].y * k[2]) * inv};
  for (int c = 0; c < 3; ++c) {
    const cx<real> r{d[c].x - p.x * k[c] - nu * k2 * uu[c].x, d[c].y - p.y * k[c] - nu * k2 * uu[c].y};
    if (u0 == nullptr) {
      du[c * n + i] = r;
    } else {
      cx<real> s = u1[c * n + i];
      s.x += a_dt * r.x;
      s.y += a_dt * r.y;
      u1[c * n + i] = s;
      const cx<real> b = u0[c * n + i];
      u[c * n + i] = last ? s : cx<real>{b.x + b_dt * r.x, b.y + b_dt * r.y};
    }
  }
}

}  // namespace b200fft
