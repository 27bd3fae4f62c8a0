"""Pseudo-spectral Navier-Stokes right-hand side on the device -- the caller of the transforms (SURVEY.md section 8
row f-2), ``/root/reference/demo/spectral_dns_solver.py:53-98``.

The demo's ``ComputeRHS`` is nine transforms and about a dozen numpy passes over dense wavenumber meshes per
Runge-Kutta stage.  Here a stage is the same nine transforms (this package's fused passes) plus THREE elementwise
kernels of ``libb200fft.so`` (``b200fft_ns_curl`` / ``_cross`` / ``_rhs``, csrc/ns_ops.cuh): the wavenumbers of a
point come from three 1D vectors, the pressure projection, the viscous term and the RK4 bookkeeping share one pass,
and on a single rank the whole stage can be replayed from a CUDA graph (``graph=True``): 30 launches, one submit.

    FFT = m.Slab_R2C(N, L, comm, "double")
    S = m.ns.Solver(FFT, nu=0.000625, dt=0.01)
    S.set_velocity(U)            # (3,) + real_shape() CUDA tensor
    S.step(); ...; k = S.kinetic_energy()

There is no CPU path: the kernels are the library's (tests reach the same point functions through the host build).
"""
import ctypes as C

import numpy as np

from . import _cdefs as D
from . import _lib

A_RK = (1. / 6., 1. / 3., 1. / 3., 1. / 6.)  # demo :48-49
B_RK = (0.5, 0.5, 1.)


class Solver(object):
    """Taylor-Green style incompressible Navier-Stokes integrator (RK4, rotational form, 3/2-rule or no dealiasing)
    around a slab transform object.  State: ``U_hat`` of shape (3,) + complex_shape()."""

    def __init__(self, FFT, nu, dt, dealias="3/2-rule", graph=False, lib=None, xp=None):
        import torch
        self.torch = xp or torch
        self.FFT, self.nu, self.dt, self.dealias = FFT, float(nu), float(dt), dealias
        self.L = lib or _lib.lib()
        self.double = FFT.float is np.float64
        self.prec = D.DOUBLE if self.double else D.SINGLE
        t = self.torch
        self.rdt, self.cdt = (t.float64, t.complex128) if self.double else (t.float32, t.complex64)
        self.cshape = tuple(int(s) for s in FFT.complex_shape())
        self.rshape = tuple(int(s) for s in FFT.real_shape())
        self.wshape = tuple(int(s) for s in FFT.work_shape(dealias))
        dev = self._device()
        z = lambda shape, dt_: t.zeros(shape, dtype=dt_, device=dev)  # noqa: E731
        self.U_hat, self.U_hat0, self.U_hat1, self.dU = (z((3,) + self.cshape, self.cdt) for _ in range(4))
        self.curl_hat = z((3,) + self.cshape, self.cdt)
        self.U_d, self.curl_d, self.cross_d = (z((3,) + self.wshape, self.rdt) for _ in range(3))
        # this rank's scaled wavenumbers as three device vectors (get_local_wavenumbermesh(scaled=True), slab.py:160-189)
        K = FFT.get_local_wavenumbermesh(scaled=True)
        self.k = [t.from_numpy(np.ascontiguousarray(np.asarray(k, dtype=FFT.float).ravel())).to(dev) for k in K]
        assert tuple(len(k) for k in self.k) == self.cshape
        self.mesh = D.NsMesh(self.prec, self.cshape[0], self.cshape[1], self.cshape[2],
                             self.k[0].data_ptr(), self.k[1].data_ptr(), self.k[2].data_ptr())
        self.n = int(np.prod(self.cshape))
        self.graph = None
        self._want_graph = bool(graph) and FFT.num_processes == 1
        self.steps_done = 0

    def _device(self):
        t = self.torch
        return t.device("cuda", t.cuda.current_device()) if t.cuda.is_available() else t.device("cpu")

    def _stream(self):
        t = self.torch
        return C.c_void_p(t.cuda.current_stream().cuda_stream) if t.cuda.is_available() else None

    # ------------------------------------------------------------------------------------------ state
    def set_velocity(self, U):
        """U: (3,) + real_shape() tensor on the transform's device (demo :80-85)."""
        for i in range(3):
            self.FFT.fftn(U[i].contiguous(), self.U_hat[i])

    def velocity(self):
        U = self.torch.zeros((3,) + self.rshape, dtype=self.rdt, device=self.U_hat.device)
        for i in range(3):
            self.FFT.ifftn(self.U_hat[i], U[i])
        return U

    def kinetic_energy(self):
        """This rank's share of sum(U*U) / (N0*N1*N2) / 2 (demo :101-102); reduce over ranks for the total."""
        U = self.velocity()
        N = self.FFT.N
        return float((U * U).sum()) / float(N[0]) / float(N[1]) / float(N[2]) / 2.0

    # ------------------------------------------------------------------------------------------ one RK stage
    def _stage(self, rk):
        """demo :66-77 + :91-97 for stage rk: 3 inverse transforms of U_hat, curl (1 kernel) + 3 inverse transforms,
        cross (1 kernel) + 3 forward transforms, then projection / viscous term / RK update (1 kernel)."""
        F, L, st, d = self.FFT, self.L, self._stream(), self.dealias
        for i in range(3):
            F.ifftn(self.U_hat[i], self.U_d[i], d)
        _lib.check(L.b200fft_ns_curl(C.byref(self.mesh), self.U_hat.data_ptr(), self.curl_hat.data_ptr(), st))
        for i in range(3):
            F.ifftn(self.curl_hat[i], self.curl_d[i], d)
        _lib.check(L.b200fft_ns_cross(self.prec, int(np.prod(self.wshape)), self.U_d.data_ptr(), self.curl_d.data_ptr(),
                                      self.cross_d.data_ptr(), st))
        for i in range(3):
            F.fftn(self.cross_d[i], self.dU[i], d)
        last = rk == 3
        _lib.check(L.b200fft_ns_rhs(C.byref(self.mesh), self.nu, self.dU.data_ptr(), self.U_hat.data_ptr(),
                                    self.U_hat0.data_ptr(), self.U_hat1.data_ptr(), A_RK[rk] * self.dt,
                                    0.0 if last else B_RK[rk] * self.dt, 1 if last else 0, st))

    def _one_step(self):
        self.U_hat0.copy_(self.U_hat)
        self.U_hat1.copy_(self.U_hat)
        for rk in range(4):
            self._stage(rk)

    def step(self):
        """One RK4 time step (demo :87-98)."""
        t = self.torch
        if self._want_graph and self.graph is None and self.steps_done >= 1:
            # plans and work buffers exist after the first eager step: capture the whole step once, replay afterwards
            s = t.cuda.Stream()
            s.wait_stream(t.cuda.current_stream())
            with t.cuda.stream(s):
                g = t.cuda.CUDAGraph()
                with t.cuda.graph(g, stream=s):
                    self._one_step()
            t.cuda.current_stream().wait_stream(s)
            self.graph = g
        if self.graph is not None:
            self.graph.replay()
        else:
            self._one_step()
        self.steps_done += 1

    def rhs(self, out=None):
        """The plain right-hand side of the current state (demo ``ComputeRHS``), for tests: (3,) + complex_shape()."""
        F, L, st, d = self.FFT, self.L, self._stream(), self.dealias
        for i in range(3):
            F.ifftn(self.U_hat[i], self.U_d[i], d)
        _lib.check(L.b200fft_ns_curl(C.byref(self.mesh), self.U_hat.data_ptr(), self.curl_hat.data_ptr(), st))
        for i in range(3):
            F.ifftn(self.curl_hat[i], self.curl_d[i], d)
        _lib.check(L.b200fft_ns_cross(self.prec, int(np.prod(self.wshape)), self.U_d.data_ptr(), self.curl_d.data_ptr(),
                                      self.cross_d.data_ptr(), st))
        for i in range(3):
            F.fftn(self.cross_d[i], self.dU[i], d)
        _lib.check(L.b200fft_ns_rhs(C.byref(self.mesh), self.nu, self.dU.data_ptr(), self.U_hat.data_ptr(), None, None,
                                    0.0, 0.0, 0, st))
        if out is not None:
            out.copy_(self.dU)
            return out
        return self.dU
