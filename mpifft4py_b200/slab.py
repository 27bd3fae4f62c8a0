"""Slab decomposition: drop-in for ``mpiFFT4py.slab.R2C`` (reference ``mpiFFT4py/slab.py:49-536``).

Same constructor, methods, shapes and slices; the transforms run as CUDA kernels on the B200
(``libb200fft.so``) with exchanges over NVLink (copy-engine pushes by default, NCCL send/recv as an option) instead of
pyfftw/numpy + MPI.  Real space is split along
x (``real_shape = (N0/P, N1, N2)``), wavenumber space along y (``complex_shape = (N0, N1/P, N2/2+1)``).
"""
from collections import defaultdict

import numpy as np
from numpy.fft import fftfreq

from . import _cdefs as D
from . import _geometry as G
from ._engine import Transform
from .mpibase import datatypes, work_arrays


class R2C(Transform):
    """3D real-to-complex FFT, slab decomposition (``slab.py:49-96``).

    Args are the reference's: N, L (numpy arrays), comm, precision ("single"/"double"),
    communication ('Alltoall', 'Sendrecv_replace', 'Alltoallw' -- all three map onto the same
    exchange and give identical results), padsize, threads and planner_effort (accepted,
    meaningless on a GPU).
    """

    _plan_kind = D.SLAB

    def __init__(self, N, L, comm, precision,
                 communication="Alltoallw",
                 padsize=1.5,
                 threads=1,
                 planner_effort=defaultdict(lambda: "FFTW_MEASURE")):
        assert len(L) == 3 and len(N) == 3
        self.float, self.complex, self.mpitype = datatypes(precision)
        self.N, self.L = N, L.astype(self.float)
        self.comm, self.communication = comm, communication
        self.num_processes, self.rank = comm.Get_size(), comm.Get_rank()
        self.padsize, self.threads, self.planner_effort = padsize, threads, planner_effort
        self.Np = N // self.num_processes          # points per rank along every axis (x in real, y in spectral space)
        self.Nf = N[2] // 2 + 1                    # kz entries of the half spectrum ...
        self.Nfp = int(padsize * N[2] // 2 + 1)    # ... and of the padded one
        self.dealias = np.zeros(0)
        self.work_arrays = work_arrays()
        legal = [2 ** i for i in range(int(np.log2(N[0])) + 1)]
        if self.num_processes not in legal:        # slab.py:89-91
            raise IOError("Number of cpus must be in ", legal)
        self._create_plan(self._plan_kind, N, self.num_processes, self.rank, comm=comm)

    # ---- shapes (slab.py:98-127, 487-489)
    def real_shape(self):
        """Local real block: this rank's x planes, all of y and z."""
        return (self.Np[0], self.N[1], self.N[2])

    def complex_shape(self):
        """Local spectral block: all of kx, this rank's ky range, the half spectrum in kz."""
        return (self.N[0], self.Np[1], self.Nf)

    def complex_shape_T(self):
        """The spectral block before the global transpose (x still local)."""
        return (self.Np[0], self.N[1], self.Nf)

    def global_real_shape(self):
        return (self.N[0], self.N[1], self.N[2])

    def global_complex_shape(self, padsize=1.):
        return (int(padsize * self.N[0]), int(padsize * self.N[1]), int(padsize * self.N[2] // 2 + 1))

    def real_shape_padded(self):
        return G.padded(self.real_shape(), self.padsize)

    def work_shape(self, dealias):
        """Real-space shape a transform with this dealias mode works on."""
        return self.real_shape_padded() if dealias == '3/2-rule' else self.real_shape()

    # ---- shapes of the 3/2-rule intermediates (slab.py:491-514).  Upstream allocates work arrays of these shapes; here
    # they describe the stages of the plan's fused passes (kept because they are public methods of the class):
    # after the z pass (_3: everything padded), after the y truncation (_2 -> _1), as per-peer blocks around the
    # exchange (_I, _0_I) and in front of the x pass (_0).
    def _pad(self, n):
        return int(self.padsize * n)

    def complex_shape_padded_0(self):
        return (self._pad(self.N[0]), self.Np[1], self.Nf)

    def complex_shape_padded_0_I(self):
        return (self.num_processes, self._pad(self.Np[0]), self.Np[1], self.Nf)

    def complex_shape_padded_1(self):
        return (self._pad(self.Np[0]), self.N[1], self.Nf)

    def complex_shape_padded_2(self):
        return (self._pad(self.Np[0]), self._pad(self.N[1]), self.Nf)

    def complex_shape_padded_3(self):
        return (self._pad(self.Np[0]), self._pad(self.N[1]), self.Nfp)

    def complex_shape_padded_I(self):
        return (self._pad(self.Np[0]), self.num_processes, self.Np[1], self.Nf)

    # ---- slices into the global arrays (slab.py:129-144): every step is 1
    def real_local_slice(self, padsize=1):
        return (G.block(self.Np[0], self.rank, padsize), G.whole(self.N[1], padsize), G.whole(self.N[2], padsize))

    def complex_local_slice(self):
        return (G.whole(self.N[0]), G.block(self.Np[1], self.rank), G.whole(self.Nf))

    # ---- meshes (slab.py:146-197)
    def _k_axes(self):
        return [G.frequencies(self.N[0]), G.frequencies(self.N[1]), G.frequencies(self.N[2], half=True)]

    def complex_local_wavenumbers(self):
        kx, ky, kz = self._k_axes()
        return (kx.astype(self.float), ky[self.complex_local_slice()[1]].astype(self.float), kz.astype(self.float))

    def get_local_mesh(self):
        """Physical coordinates of the local block as three broadcast views."""
        sl = (G.block(self.Np[0], self.rank), G.whole(self.N[1]), G.whole(self.N[2]))
        return G.sparse_physical_mesh(sl, self.N, self.L, self.float, self.real_shape())

    def get_local_wavenumbermesh(self, scaled=False, broadcast=False, eliminate_highest_freq=False):
        """Sparse (or broadcast) wavenumber mesh of the local spectral block, optionally scaled by 2 pi / L."""
        kx, ky, kz = self.complex_local_wavenumbers()
        if eliminate_highest_freq:
            ky = G.frequencies(self.N[1])   # the whole axis: the Nyquist entry may lie outside this rank's range
            G.drop_nyquist((kx, ky, kz), self.N)
            ky = ky[self.complex_local_slice()[1]]
        K = [k.astype(self.float) for k in G.sparse_spectral_mesh((kx, ky, kz))]
        if scaled:
            for k, f in zip(K, 2 * np.pi / self.L):
                k *= f
        return G.Vectors([np.broadcast_to(k, self.complex_shape()) for k in K] if broadcast is True else K)

    def get_dealias_filter(self):
        """2/3-rule mask on the local spectral block.  The transforms apply it inside the first inverse FFT
        pass (mask bands of the load index map); the array is for callers that filter spectra themselves."""
        return G.two_thirds_mask(self.get_local_wavenumbermesh(), self.N)

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

    # The reference's copy_to_padded / copy_from_padded helpers (slab.py:516-536) have no counterpart here: the pad /
    # truncate copies are index maps inside the FFT passes and the intermediates live in the plan's work buffers.


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
        kx, ky, kz = (G.frequencies(n) for n in self.N)
        return (kx, ky[self.transformed_local_slice()[1]], kz)

    def get_dealias_filter(self):
        return G.two_thirds_mask(G.sparse_spectral_mesh(self.transformed_local_wavenumbers()), self.N)

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
