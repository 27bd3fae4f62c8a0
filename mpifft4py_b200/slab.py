"""Slab decomposition: drop-in for ``mpiFFT4py.slab.R2C`` (reference ``mpiFFT4py/slab.py:49-536``).

Same constructor, methods, shapes and slices; the transforms run as CUDA kernels on the B200
(``libb200fft.so``) with NCCL exchanges instead of pyfftw/numpy + MPI.  Real space is split along
x (``real_shape = (N0/P, N1, N2)``), wavenumber space along y (``complex_shape = (N0, N1/P, N2/2+1)``).
"""
from collections import defaultdict

import numpy as np
from numpy.fft import fftfreq, rfftfreq

from . import _cdefs as D
from ._engine import Transform
from .mpibase import datatypes, work_arrays


class R2C(Transform):
    """3D real-to-complex FFT, slab decomposition (``slab.py:49-96``).

    Args are the reference's: N, L (numpy arrays), comm, precision ("single"/"double"),
    communication ('Alltoall', 'Sendrecv_replace', 'Alltoallw' -- all three map onto the same
    NCCL exchange and give identical results), padsize, threads and planner_effort (accepted,
    meaningless on a GPU).
    """

    _plan_kind = D.SLAB

    def __init__(self, N, L, comm, precision,
                 communication="Alltoallw",
                 padsize=1.5,
                 threads=1,
                 planner_effort=defaultdict(lambda: "FFTW_MEASURE")):
        assert len(L) == 3
        assert len(N) == 3
        self.N = N
        self.Nf = N[2]//2+1
        self.Nfp = int(padsize*N[2]//2+1)
        self.comm = comm
        self.float, self.complex, self.mpitype = datatypes(precision)
        self.communication = communication
        self.num_processes = comm.Get_size()
        self.rank = comm.Get_rank()
        self.Np = N // self.num_processes
        self.L = L.astype(self.float)
        self.dealias = np.zeros(0)
        self.padsize = padsize
        self.threads = threads
        self.planner_effort = planner_effort
        self.work_arrays = work_arrays()
        if not self.num_processes in [2**i for i in range(int(np.log2(N[0]))+1)]:
            raise IOError("Number of cpus must be in ",
                          [2**i for i in range(int(np.log2(N[0]))+1)])
        self._create_plan(self._plan_kind, N, self.num_processes, self.rank, comm=comm)

    def real_shape(self):
        """The local shape of the real data"""
        return (self.Np[0], self.N[1], self.N[2])

    def complex_shape(self):
        """The local shape of the complex data"""
        return (self.N[0], self.Np[1], self.Nf)

    def complex_shape_T(self):
        """The local transposed shape of the complex data"""
        return (self.Np[0], self.N[1], self.Nf)

    def global_real_shape(self):
        """Global size of problem in real physical space"""
        return (self.N[0], self.N[1], self.N[2])

    def global_complex_shape(self, padsize=1.):
        """Global size of problem in complex wavenumber space"""
        return (int(padsize*self.N[0]), int(padsize*self.N[1]),
                int(padsize*self.N[2]//2+1))

    def work_shape(self, dealias):
        """Shape of work arrays used in convection with dealiasing (``slab.py:118-127``)."""
        if dealias == '3/2-rule':
            return self.real_shape_padded()
        else:
            return self.real_shape()

    def real_local_slice(self, padsize=1):
        """Local slice in real space of the input array (``slab.py:129-138``)."""
        return (slice(int(padsize*self.rank*self.Np[0]),
                      int(padsize*(self.rank+1)*self.Np[0]), 1),
                slice(0, int(padsize*self.N[1]), 1),
                slice(0, int(padsize*self.N[2]), 1))

    def complex_local_slice(self):
        """Local slice of complex return array (``slab.py:140-144``)."""
        return (slice(0, self.N[0], 1),
                slice(self.rank*self.Np[1], (self.rank+1)*self.Np[1], 1),
                slice(0, self.Nf, 1))

    def complex_local_wavenumbers(self):
        """Returns local wavenumbers of complex space"""
        return (fftfreq(self.N[0], 1./self.N[0]).astype(self.float),
                fftfreq(self.N[1], 1./self.N[1])[self.complex_local_slice()[1]].astype(self.float),
                rfftfreq(self.N[2], 1./self.N[2]).astype(self.float))

    def get_local_mesh(self):
        """Returns the local decomposed physical mesh (``slab.py:152-160``)."""
        X = list(np.ogrid[self.rank*self.Np[0]:(self.rank+1)*self.Np[0],
                          :self.N[1], :self.N[2]])
        X[0] = (X[0]*self.L[0]/self.N[0]).astype(self.float)
        X[1] = (X[1]*self.L[1]/self.N[1]).astype(self.float)
        X[2] = (X[2]*self.L[2]/self.N[2]).astype(self.float)
        X = [np.broadcast_to(x, self.real_shape()) for x in X]
        return X

    def get_local_wavenumbermesh(self, scaled=False, broadcast=False, eliminate_highest_freq=False):
        """Returns (scaled) local decomposed wavenumbermesh (``slab.py:162-189``)."""
        kx, ky, kz = self.complex_local_wavenumbers()
        if eliminate_highest_freq:
            ky = fftfreq(self.N[1], 1./self.N[1].astype(self.float))
            for i, k in enumerate((kx, ky, kz)):
                if self.N[i] % 2 == 0:
                    k[self.N[i]//2] = 0
            ky = ky[self.complex_local_slice()[1]]

        Ks = list(np.meshgrid(kx, ky, kz, indexing='ij', sparse=True))
        for i in range(3):
            Ks[i] = Ks[i].astype(self.float)
        if scaled:
            Lp = 2*np.pi/self.L
            for i in range(3):
                Ks[i] *= Lp[i]
        K = Ks
        if broadcast is True:
            K = [np.broadcast_to(k, self.complex_shape()) for k in Ks]
        return K

    def get_dealias_filter(self):
        """Filter for dealiasing nonlinear convection (``slab.py:191-197``).  The transforms apply
        this mask inside the first inverse FFT pass; the array is returned for callers only."""
        K = self.get_local_wavenumbermesh()
        kmax = 2./3.*(self.N//2+1)
        dealias = np.array((abs(K[0]) < kmax[0])*(abs(K[1]) < kmax[1])*
                           (abs(K[2]) < kmax[2]), dtype=np.uint8)
        return dealias

    def ifftn(self, fu, u, dealias=None):
        """Inverse transform (``slab.py:214-346``): fu of complex_shape() -> u of real_shape(), or
        of real_shape_padded() for dealias='3/2-rule'.  fu is not modified."""
        assert dealias in ('3/2-rule', '2/3-rule', 'None', None)
        if dealias == '3/2-rule':
            assert self.num_processes <= self.N[0]//2 or self.num_processes == 1, \
                "Number of processors cannot be larger than N[0]//2 for 3/2-rule"
            ushape = self.real_shape_padded()
        else:
            ushape = self.real_shape()
        return self._run(1, fu, u, dealias, self.complex_shape(), self.complex, ushape, self.float)

    def fftn(self, u, fu, dealias=None):
        """Forward transform (``slab.py:349-485``): u of real_shape() [3/2-rule: real_shape_padded()]
        -> fu of complex_shape()."""
        assert dealias in ('3/2-rule', '2/3-rule', 'None', None)
        if dealias == '3/2-rule':
            assert self.num_processes <= self.N[0]//2 or self.num_processes == 1, \
                "Number of processors cannot be larger than N[0]//2 for 3/2-rule"
            ushape = self.real_shape_padded()
        else:
            ushape = self.real_shape()
        return self._run(0, u, fu, dealias, ushape, self.float, self.complex_shape(), self.complex)

    def real_shape_padded(self):
        """The local shape of the real data"""
        return (int(self.padsize*self.Np[0]), int(self.padsize*self.N[1]), int(self.padsize*self.N[2]))

    # The reference's intermediate shapes (complex_shape_padded_0 ... _I) and its copy_to_padded /
    # copy_from_padded helpers (slab.py:491-536) have no counterpart here: the pad / truncate copies are index
    # maps inside the FFT passes and the intermediates live in the plan's work buffers.


class C2C(R2C):
    """3D complex-to-complex FFT, slab decomposition: drop-in for ``mpiFFT4py.slab.C2C``
    (``slab.py:538-825``).  Reuses every R2C shape with ``Nf = N[2]`` (``:565-567``); both arrays are
    complex.  Same fused passes as R2C with a contiguous-row C2C kernel along z.

    3/2-rule semantics follow the reference: on several ranks the truncation folds the two Nyquist
    modes in y and z (``copy_from_padded``, ``:816-823``) and keeps mode -N/2 in x (``:796-797``); on
    one rank it keeps mode -N/2 in all three directions (the ``ks`` gather, ``:735-738``).
    ``dealias='2/3-rule'`` raises a broadcast ``ValueError`` upstream (the mask is inherited from
    R2C with an rfft-sized z extent); here it applies the intended mask over the full kz range.
    """

    _plan_kind = D.SLAB_C2C

    def __init__(self, N, L, comm, precision,
                 communication="Alltoall",
                 padsize=1.5,
                 threads=1,
                 planner_effort=defaultdict(lambda: "FFTW_MEASURE")):
        R2C.__init__(self, N, L, comm, precision, communication=communication, padsize=padsize,
                     threads=threads, planner_effort=planner_effort)
        self.Nf = N[2]
        self.Nfp = int(self.padsize*self.N[2])
        # Rename since there's no real space (slab.py:569-575)
        self.original_shape_padded = self.real_shape_padded
        self.original_shape = self.real_shape
        self.transformed_shape = self.complex_shape
        self.original_local_slice = self.real_local_slice
        self.transformed_local_slice = self.complex_local_slice
        self.ks = (fftfreq(N[2])*N[2]).astype(int)

    def global_shape(self, padsize=1.):
        """Global size of problem in transformed space"""
        return (int(padsize*self.N[0]), int(padsize*self.N[1]), int(padsize*self.N[2]))

    def transformed_local_wavenumbers(self):
        return (fftfreq(self.N[0], 1./self.N[0]),
                fftfreq(self.N[1], 1./self.N[1])[self.transformed_local_slice()[1]],
                fftfreq(self.N[2], 1./self.N[2]))

    def get_dealias_filter(self):
        kx, ky, kz = self.transformed_local_wavenumbers()
        K = np.meshgrid(kx, ky, kz, indexing='ij', sparse=True)
        kmax = 2./3.*(self.N//2+1)
        return np.array((abs(K[0]) < kmax[0])*(abs(K[1]) < kmax[1])*(abs(K[2]) < kmax[2]), dtype=np.uint8)

    def ifftn(self, fu, u, dealias=None):
        """``slab.py:587-698``: fu of transformed_shape() -> u of original_shape() [3/2-rule:
        original_shape_padded()], both complex.  fu is not modified."""
        assert dealias in ('3/2-rule', '2/3-rule', 'None', None)
        ushape = self.real_shape_padded() if dealias == '3/2-rule' else self.real_shape()
        return self._run(1, fu, u, dealias, self.complex_shape(), self.complex, ushape, self.complex)

    def fftn(self, u, fu, dealias=None):
        """``slab.py:700-800``."""
        assert dealias in ('3/2-rule', '2/3-rule', 'None', None)
        ushape = self.real_shape_padded() if dealias == '3/2-rule' else self.real_shape()
        return self._run(0, u, fu, dealias, ushape, self.complex, self.complex_shape(), self.complex)
