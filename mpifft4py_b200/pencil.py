"""Pencil decomposition: drop-in for ``mpiFFT4py.pencil`` (reference ``mpiFFT4py/pencil.py:76-1484``).

``R2CY`` leaves the spectral data aligned in y (``pencil.py:145-883``), ``R2CX`` in x
(``:885-1477``); ``R2C(...)`` is the reference's factory (``:1479-1484``).  Real blocks are
``(N0/P1, N1/P2, N2)``; the process grid is ``comm0`` (P1 consecutive ranks, rank ``r % P1``) x
``comm1`` (P2 ranks with equal ``r % P1``, rank ``r // P1``), ``pencil.py:184-195``.

communication: 'Alltoall' and 'Alltoallw' give the same arrays in the same layout upstream and
share one NCCL path here (uneven last z-chunk carries the Nyquist plane, no pack trick and no
Scatter/Send/Recv); 'AlltoallN' keeps its own layout with the Nyquist plane dropped.
"""
from collections import defaultdict

import numpy as np
from numpy.fft import fftfreq, rfftfreq

from . import _cdefs as D
from ._engine import Transform
from .mpibase import datatypes, work_arrays

__all__ = ['R2C', 'R2CX', 'R2CY']


def _compute_dims(P):
    """MPI.Compute_dims(P, 2) (``pencil.py:185``): balanced factors, larger first (8 -> [4, 2])."""
    best = (P, 1)
    for a in range(1, int(P ** 0.5) + 1):
        if P % a == 0:
            best = (P // a, a)
    return best


def _subsize(N, size, rank):
    return N // size + ((N % size) * (rank == size - 1))


class R2CY(Transform):
    """3D R2C FFT, pencil decomposition, final alignment in y (``pencil.py:145-216``)."""

    _kind = D.PENCIL_Y

    def __init__(self, N, L, comm, precision, P1=None, communication='Alltoallw', padsize=1.5, threads=1,
                 planner_effort=defaultdict(lambda: "FFTW_MEASURE")):
        self.N = N
        assert len(L) == 3
        assert len(N) == 3
        self.Nf = N[2]//2+1
        self.comm = comm
        self.float, self.complex, self.mpitype = datatypes(precision)
        self.num_processes = comm.Get_size()
        assert self.num_processes > 1
        self.L = L.astype(self.float)   # pencil.py:173,177: `float` is rebound to the numpy type there
        self.dealias = np.zeros(0)
        self.communication = communication
        self.padsize = padsize
        self.threads = threads
        self.planner_effort = planner_effort
        self.rank = comm.Get_rank()
        if P1 is None:
            P1, P2 = _compute_dims(self.num_processes)
            self.P1, self.P2 = P1, P2
        else:
            self.P1 = P1
            self.P2 = P2 = self.num_processes // P1
        self.N1 = N // P1
        self.N2 = N // P2
        if not (self.num_processes % 2 == 0 or self.num_processes == 1):
            raise IOError("Number of cpus must be even")
        if (P1 % 2 != 0) or (P2 % 2 != 0):
            raise IOError("Number of cpus in each direction must be even power of 2")
        self.comm0 = comm.Split(self.rank // P1)   # pencil.py:192 (true division upstream, Q4)
        self.comm1 = comm.Split(self.rank % P1)
        self.comm0_rank = self.comm0.Get_rank()
        self.comm1_rank = self.comm1.Get_rank()
        self.work_arrays = work_arrays()
        self.N1f = self.N1[2]//2 if self.comm0_rank < self.P1-1 else self.N1[2]//2+1
        if self.communication == 'AlltoallN':
            self.N1f = self.N1[2]//2
        self._init_alignment()
        self._create_plan(self._kind, N, self.num_processes, self.rank, P1=self.P1, P2=self.P2,
                          drop_nyquist=int(self.communication == 'AlltoallN'),
                          comm=comm, comm0=self.comm0, comm1=self.comm1)

    def _init_alignment(self):
        pass

    def real_shape(self):
        """The local shape of the real data"""
        return (self.N1[0], self.N2[1], self.N[2])

    def complex_shape(self):
        """The local shape of the complex data"""
        return (self.N2[0], self.N[1], self.N1f)

    def real_shape_padded(self):
        return (int(self.padsize*self.N1[0]), int(self.padsize*self.N2[1]), int(self.padsize*self.N[2]))

    def work_shape(self, dealias):
        if dealias == '3/2-rule':
            return self.real_shape_padded()
        else:
            return self.real_shape()

    def real_local_slice(self, padsize=1):
        xzrank = self.comm0.Get_rank()
        xyrank = self.comm1.Get_rank()
        return (slice(int(padsize * xzrank * self.N1[0]), int(padsize * (xzrank+1) * self.N1[0]), 1),
                slice(int(padsize * xyrank * self.N2[1]), int(padsize * (xyrank+1) * self.N2[1]), 1),
                slice(0, int(padsize*self.N[2])))

    def complex_local_slice(self):
        xzrank = self.comm0.Get_rank()
        xyrank = self.comm1.Get_rank()
        return (slice(xyrank*self.N2[0], (xyrank+1)*self.N2[0], 1),
                slice(0, self.N[1]),
                slice(xzrank*self.N1[2]//2, xzrank*self.N1[2]//2 + self.N1f, 1))

    def complex_local_wavenumbers(self):
        s = self.complex_local_slice()
        return (fftfreq(self.N[0], 1./self.N[0]).astype(int)[s[0]],
                fftfreq(self.N[1], 1./self.N[1]).astype(int),
                rfftfreq(self.N[2], 1./self.N[2]).astype(int)[s[2]])

    def get_P(self):
        return self.P1, self.P2

    def get_local_mesh(self):
        xzrank = self.comm0.Get_rank()
        xyrank = self.comm1.Get_rank()
        x1 = slice(xzrank * self.N1[0], (xzrank+1) * self.N1[0], 1)
        x2 = slice(xyrank * self.N2[1], (xyrank+1) * self.N2[1], 1)
        X = list(np.ogrid[x1, x2, :self.N[2]])
        X[0] = (X[0]*self.L[0]/self.N[0]).astype(self.float)
        X[1] = (X[1]*self.L[1]/self.N[1]).astype(self.float)
        X[2] = (X[2]*self.L[2]/self.N[2]).astype(self.float)
        X = [np.broadcast_to(x, self.real_shape()) for x in X]
        return X

    def get_local_wavenumbermesh(self, scaled=False, broadcast=False,
                                 eliminate_highest_freq=False):
        """``pencil.py:311-341``: integer wavenumbers unless scaled."""
        s = self.complex_local_slice()
        kx = fftfreq(self.N[0], 1./self.N[0]).astype(int)
        ky = fftfreq(self.N[1], 1./self.N[1]).astype(int)
        kz = rfftfreq(self.N[2], 1./self.N[2]).astype(int)
        if eliminate_highest_freq:
            for i, k in enumerate((kx, ky, kz)):
                if self.N[i] % 2 == 0:
                    k[self.N[i]//2] = 0
        kx = kx[s[0]]
        kz = kz[s[2]]
        Ks = list(np.meshgrid(kx, ky, kz, indexing='ij', sparse=True))
        if scaled is True:
            Lp = 2*np.pi/self.L
            for i in range(3):
                Ks[i] = (Ks[i]*Lp[i]).astype(self.float)
        K = Ks
        if broadcast is True:
            K = [np.broadcast_to(k, self.complex_shape()) for k in Ks]
        return K

    def get_dealias_filter(self):
        """2/3-rule mask on the local spectral block (``pencil.py:343-349``)."""
        s = self.complex_local_slice()
        kx = fftfreq(self.N[0], 1./self.N[0]).astype(int)[s[0]]
        ky = fftfreq(self.N[1], 1./self.N[1]).astype(int)[s[1]]
        kz = rfftfreq(self.N[2], 1./self.N[2]).astype(int)[s[2]]
        K = np.meshgrid(kx, ky, kz, indexing='ij', sparse=True)
        kmax = 2./3.*(self.N//2+1)
        dealias = np.array((abs(K[0]) < kmax[0])*(abs(K[1]) < kmax[1])*
                           (abs(K[2]) < kmax[2]), dtype=np.uint8)
        return dealias

    # (copy_to_padded_* / copy_from_padded_* of pencil.py:351-379: index maps inside the FFT passes here)

    def global_complex_shape(self, padsize=1.0):
        """Global size of problem in complex wavenumber space"""
        return (int(padsize*self.N[0]), int(padsize*self.N[1]),
                int(padsize*self.N[2]//2+1))

    def ifftn(self, fu, u, dealias=None):
        """Inverse transform (Y: ``pencil.py:386-632``; X: ``:1001-1226``).  fu is not modified.
        2/3-rule follows the slab/R2CY semantics for both alignments (the reference's R2CX variant
        returns zeros)."""
        assert dealias in ('3/2-rule', '2/3-rule', 'None', None)
        ushape = self.real_shape_padded() if dealias == '3/2-rule' else self.real_shape()
        return self._run(1, fu, u, dealias, self.complex_shape(), self.complex, ushape, self.float)

    def fftn(self, u, fu, dealias=None):
        """Forward transform (Y: ``pencil.py:634-883``; X: ``:1228-1477``)."""
        assert dealias in ('3/2-rule', '2/3-rule', 'None', None)
        ushape = self.real_shape_padded() if dealias == '3/2-rule' else self.real_shape()
        return self._run(0, u, fu, dealias, ushape, self.float, self.complex_shape(), self.complex)


class R2CX(R2CY):
    """3D R2C FFT, pencil decomposition, final alignment in x (``pencil.py:885-969``)."""

    _kind = D.PENCIL_X

    def __init__(self, N, L, comm, precision, P1=None, communication='Alltoall',
                 padsize=1.5, threads=1,
                 planner_effort=defaultdict(lambda: "FFTW_MEASURE")):
        R2CY.__init__(self, N, L, comm, precision, P1=P1, communication=communication,
                      padsize=padsize, threads=threads, planner_effort=planner_effort)

    def _init_alignment(self):
        self.N2f = self.N2[2]//2 if self.comm1_rank < self.P2-1 else self.N2[2]//2+1
        if self.communication == 'AlltoallN':
            self.N2f = self.N2[2]//2
        if self.communication == 'Alltoallw':
            self.N2f = _subsize(self.Nf, self.P2, self.comm1_rank)

    def complex_shape(self):
        """The local shape of the complex data"""
        return (self.N[0], self.N1[1], self.N2f)

    def real_local_slice(self, padsize=1):
        xyrank = self.comm0.Get_rank()
        yzrank = self.comm1.Get_rank()
        return (slice(int(padsize * xyrank * self.N1[0]), int(padsize * (xyrank+1) * self.N1[0]), 1),
                slice(int(padsize * yzrank * self.N2[1]), int(padsize * (yzrank+1) * self.N2[1]), 1),
                slice(0, int(padsize * self.N[2])))

    def complex_local_slice(self):
        xyrank = self.comm0.Get_rank()
        yzrank = self.comm1.Get_rank()
        return (slice(0, self.N[0]),
                slice(xyrank*self.N1[1], (xyrank+1)*self.N1[1], 1),
                slice(yzrank*self.N2[2]//2, yzrank*self.N2[2]//2 + self.N2f, 1))

    def get_local_mesh(self):
        xyrank = self.comm0.Get_rank()
        yzrank = self.comm1.Get_rank()
        x1 = slice(xyrank * self.N1[0], (xyrank+1) * self.N1[0], 1)
        x2 = slice(yzrank * self.N2[1], (yzrank+1) * self.N2[1], 1)
        X = np.mgrid[x1, x2, :self.N[2]].astype(self.float)
        X[0] *= self.L[0]/self.N[0]
        X[1] *= self.L[1]/self.N[1]
        X[2] *= self.L[2]/self.N[2]
        return X

    def get_local_wavenumbermesh(self):
        """Dense float mesh of shape (3, N0, N1[1], N2[2]//2) -- Nyquist plane excluded, exactly
        as ``pencil.py:945-957``."""
        xyrank = self.comm0.Get_rank()
        yzrank = self.comm1.Get_rank()
        kx = fftfreq(self.N[0], 1./self.N[0]).astype(int)
        ky = fftfreq(self.N[1], 1./self.N[1]).astype(int)
        kz = fftfreq(self.N[2], 1./self.N[2]).astype(int)
        k2 = slice(xyrank*self.N1[1], (xyrank+1)*self.N1[1], 1)
        k1 = slice(yzrank*self.N2[2]//2, (yzrank+1)*self.N2[2]//2, 1)
        K = np.array(np.meshgrid(kx, ky[k2], kz[k1], indexing='ij'), dtype=self.float)
        return K


def R2C(N, L, comm, precision, P1=None, communication="Alltoall", padsize=1.5, threads=1,
        alignment="X", planner_effort=defaultdict(lambda: "FFTW_MEASURE")):
    """Factory of ``pencil.py:1479-1484``."""
    if alignment == 'X':
        return R2CX(N, L, comm, precision, P1, communication, padsize, threads, planner_effort)
    else:
        return R2CY(N, L, comm, precision, P1, communication, padsize, threads, planner_effort)
