"""2D row decomposition: drop-in for ``mpiFFT4py.line.R2C`` (reference ``mpiFFT4py/line.py:41-340``).

Real rows are split along x (``real_shape = (N0/P, N1)``), spectral columns along ky
(``complex_shape = (N0, Npf)``, the last rank owning the Nyquist column).  The reference's
Nyquist pack trick, Scatter and Send/Recv (``line.py:206,217-223,302-306``) are replaced by an
exchange with one extra column for the last rank; see DESIGN.md (deviation D3) for the one input
class where the pack trick itself is inexact upstream.
"""
from collections import defaultdict

import numpy as np
from numpy.fft import fftfreq, rfftfreq

from . import _cdefs as D
from ._engine import Transform
from .mpibase import datatypes, work_arrays, zeros


class R2C(Transform):
    """2D real-to-complex FFT (``fft2``/``ifft2``), row decomposition (``line.py:41-75``)."""

    def __init__(self, N, L, comm, precision, padsize=1.5, threads=1,
                 planner_effort=defaultdict(lambda: "FFTW_MEASURE")):
        self.N = N
        self.L = L
        assert len(L) == 2
        assert len(N) == 2
        self.comm = comm
        self.float, self.complex, self.mpitype = datatypes(precision)
        self.num_processes = comm.Get_size()
        self.rank = comm.Get_rank()
        self.padsize = padsize
        self.threads = threads
        self.planner_effort = planner_effort
        self.Np = N // self.num_processes
        self.Nf = N[1]//2+1
        self.Npf = self.Np[1]//2+1 if self.rank+1 == self.num_processes else self.Np[1]//2
        self.Nfp = int(padsize*self.N[1]/2+1)
        self.ks = (fftfreq(N[0])*N[0]).astype(int)
        self.dealias = np.zeros(0)
        self.work_arrays = work_arrays()
        self._create_plan(D.LINE, N, self.num_processes, self.rank, comm=comm)

    def real_shape(self):
        """The local shape of the real data"""
        return (self.Np[0], self.N[1])

    def complex_shape(self):
        """The local shape of the complex data"""
        return (self.N[0], self.Npf)

    def global_complex_shape(self):
        return (self.N[0], self.Nf)

    def global_real_shape(self):
        return (self.N[0], self.N[1])

    def real_local_slice(self, padsize=1):
        return (slice(int(padsize*self.rank*self.Np[0]),
                      int(padsize*(self.rank+1)*self.Np[0]), 1),
                slice(0, int(padsize*self.N[1])))

    def complex_local_slice(self):
        return (slice(0, self.N[0]),
                slice(self.rank*self.Np[1]//2, self.rank*self.Np[1]//2+self.Npf, 1))

    def get_N(self):
        return self.N

    def get_local_mesh(self):
        X = np.mgrid[self.rank*self.Np[0]:(self.rank+1)*self.Np[0], :self.N[1]].astype(self.float)
        X[0] *= self.L[0]/self.N[0]
        X[1] *= self.L[1]/self.N[1]
        return X

    def get_local_wavenumbermesh(self, scaled=True, broadcast=False,
                                 eliminate_highest_freq=False):
        """``line.py:112-129`` (note scaled=True is the default here, unlike slab/pencil)."""
        kx = fftfreq(self.N[0], 1./self.N[0])
        ky = rfftfreq(self.N[1], 1./self.N[1])
        if eliminate_highest_freq:
            for i, k in enumerate((kx, ky)):
                if self.N[i] % 2 == 0:
                    k[self.N[i]//2] = 0

        Ks = list(np.meshgrid(kx, ky[self.rank*self.Np[1]//2:(self.rank*self.Np[1]//2+self.Npf)], indexing='ij', sparse=True))
        if scaled is True:
            Lp = 2*np.pi/self.L
            Ks[0] *= Lp[0]
            Ks[1] *= Lp[1]
        K = Ks
        if broadcast is True:
            K = [np.broadcast_to(k, self.complex_shape()) for k in Ks]
        return K

    def get_dealias_filter(self):
        """``line.py:131-136``.  The engine's fused mask uses the unscaled wavenumbers, which is
        the same thing for the 2*pi-periodic box the reference's mask assumes."""
        K = self.get_local_wavenumbermesh()
        kmax = 2./3.*(self.N//2+1)
        dealias = np.array((abs(K[0]) < kmax[0])*(abs(K[1]) < kmax[1]), dtype=np.uint8)
        return dealias

    def global_complex_shape_padded(self):
        return (int(self.padsize*self.N[0]), int(self.padsize*self.N[1]/2+1))

    def real_shape_padded(self):
        return (int(self.padsize*self.Np[0]), int(self.padsize*self.N[1]))

    def work_shape(self, dealias):
        if dealias == '3/2-rule':
            return self.real_shape_padded()
        else:
            return self.real_shape()

    def fft2(self, u, fu, dealias=None):
        """Forward 2D transform (``line.py:179-260``)."""
        assert dealias in ('3/2-rule', '2/3-rule', 'None', None)
        ushape = self.real_shape_padded() if dealias == '3/2-rule' else self.real_shape()
        return self._run(0, u, fu, dealias, ushape, self.float, self.complex_shape(), self.complex)

    def ifft2(self, fu, u, dealias=None):
        """Inverse 2D transform (``line.py:262-340``).  fu is not modified."""
        assert dealias in ('3/2-rule', '2/3-rule', 'None', None)
        ushape = self.real_shape_padded() if dealias == '3/2-rule' else self.real_shape()
        return self._run(1, fu, u, dealias, self.complex_shape(), self.complex, ushape, self.float)
