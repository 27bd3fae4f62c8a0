"""Pencil decomposition: drop-in for ``mpiFFT4py.pencil`` (reference ``mpiFFT4py/pencil.py:76-1484``).

``R2CY`` leaves the spectral data aligned in y (``pencil.py:145-883``), ``R2CX`` in x
(``:885-1477``); ``R2C(...)`` is the reference's factory (``:1479-1484``).  Real blocks are
``(N0/P1, N1/P2, N2)``; the process grid is ``comm0`` (P1 consecutive ranks, rank ``r % P1``) x
``comm1`` (P2 ranks with equal ``r % P1``, rank ``r // P1``), ``pencil.py:184-195``.

communication: 'Alltoall' and 'Alltoallw' give the same arrays in the same layout upstream and
share one exchange path here (uneven last z-chunk carries the Nyquist plane, no pack trick and no
Scatter/Send/Recv); 'AlltoallN' keeps its own layout with the Nyquist plane dropped.
"""
from collections import defaultdict

import numpy as np
from . import _cdefs as D
from . import _geometry as G
from ._engine import Transform
from .mpibase import datatypes, work_arrays

__all__ = ['R2C', 'R2CX', 'R2CY']


class R2CY(Transform):
    """3D R2C FFT, pencil decomposition, final alignment in y (``pencil.py:145-216``).

    Real space is cut along x by ``comm0`` (P1 ranks) and along y by ``comm1`` (P2 ranks); the spectral block holds
    all of ky, the comm1-th part of kx and the comm0-th part of kz (the last part carrying the Nyquist plane)."""

    _kind = D.PENCIL_Y

    def __init__(self, N, L, comm, precision, P1=None, communication='Alltoallw', padsize=1.5, threads=1,
                 planner_effort=defaultdict(lambda: "FFTW_MEASURE")):
        assert len(L) == 3 and len(N) == 3
        self.float, self.complex, self.mpitype = datatypes(precision)
        self.N, self.L = N, L.astype(self.float)
        self.Nf = N[2] // 2 + 1
        self.comm, self.communication = comm, communication
        self.padsize, self.threads, self.planner_effort = padsize, threads, planner_effort
        self.num_processes, self.rank = comm.Get_size(), comm.Get_rank()
        assert self.num_processes > 1                                  # pencil.py:176
        self.dealias = np.zeros(0)
        self.P1, self.P2 = G.balanced_grid(self.num_processes) if P1 is None else (P1, self.num_processes // P1)
        self.N1, self.N2 = N // self.P1, N // self.P2                  # points per rank of either grid direction
        if self.num_processes % 2:                                     # pencil.py:201-205
            raise IOError("Number of cpus must be even")
        if self.P1 % 2 or self.P2 % 2:
            raise IOError("Number of cpus in each direction must be even power of 2")
        # grid: comm0 = P1 consecutive ranks (position rank % P1), comm1 = the P2 ranks with equal rank % P1
        # (pencil.py:192-195; upstream passes rank / P1, a float under true division)
        self.comm0 = comm.Split(self.rank // self.P1)
        self.comm1 = comm.Split(self.rank % self.P1)
        self.comm0_rank, self.comm1_rank = self.comm0.Get_rank(), self.comm1.Get_rank()
        self.work_arrays = work_arrays()
        drop = communication == 'AlltoallN'                            # that layout has no Nyquist plane at all
        self.N1f = self._kz_count(self.N1, self.comm0_rank, self.P1, drop)
        self._init_alignment(drop)
        self._create_plan(self._kind, N, self.num_processes, self.rank, P1=self.P1, P2=self.P2,
                          drop_nyquist=int(drop), comm=comm, comm0=self.comm0, comm1=self.comm1)

    @staticmethod
    def _kz_count(Npart, idx, parts, drop):
        """kz entries of part ``idx``: half the part's z extent, plus the Nyquist entry on the last part."""
        return Npart[2] // 2 + (0 if (drop or idx < parts - 1) else 1)

    def _init_alignment(self, drop):
        pass

    # which block of x, y and kz this rank owns (alignment Y: kx by comm1, kz by comm0)
    def _grid(self):
        return self.comm0.Get_rank(), self.comm1.Get_rank()

    def get_P(self):
        return self.P1, self.P2

    # ---- shapes
    def real_shape(self):
        return (self.N1[0], self.N2[1], self.N[2])

    def complex_shape(self):
        return (self.N2[0], self.N[1], self.N1f)

    def real_shape_padded(self):
        return G.padded(self.real_shape(), self.padsize)

    def work_shape(self, dealias):
        return self.real_shape_padded() if dealias == '3/2-rule' else self.real_shape()

    def global_complex_shape(self, padsize=1.0):
        return (int(padsize * self.N[0]), int(padsize * self.N[1]), int(padsize * self.N[2] // 2 + 1))

    # ---- slices (pencil.py:264-287; the z slice of the real block and the ky slice carry no step upstream)
    def real_local_slice(self, padsize=1):
        c0, c1 = self._grid()
        return (G.block(self.N1[0], c0, padsize), G.block(self.N2[1], c1, padsize), G.whole(self.N[2], padsize, None))

    def _kz_slice(self, Npart, idx, count):
        lo = idx * Npart[2] // 2
        return slice(lo, lo + count, 1)

    def complex_local_slice(self):
        c0, c1 = self._grid()
        return (G.block(self.N2[0], c1), G.whole(self.N[1], 1, None), self._kz_slice(self.N1, c0, self.N1f))

    # ---- meshes (pencil.py:289-349): integer wavenumbers unless scaled
    def _k_axes(self):
        return [G.frequencies(self.N[0]).astype(int), G.frequencies(self.N[1]).astype(int),
                G.frequencies(self.N[2], half=True).astype(int)]

    def complex_local_wavenumbers(self):
        s = self.complex_local_slice()
        kx, ky, kz = self._k_axes()
        return (kx[s[0]], ky, kz[s[2]])

    def get_local_mesh(self):
        c0, c1 = self._grid()
        sl = (G.block(self.N1[0], c0), G.block(self.N2[1], c1), G.whole(self.N[2]))
        return G.sparse_physical_mesh(sl, self.N, self.L, self.float, self.real_shape())

    def get_local_wavenumbermesh(self, scaled=False, broadcast=False, eliminate_highest_freq=False):
        s = self.complex_local_slice()
        ks = self._k_axes()
        if eliminate_highest_freq:
            G.drop_nyquist(ks, self.N)
        K = G.sparse_spectral_mesh((ks[0][s[0]], ks[1], ks[2][s[2]]))
        if scaled is True:
            K = [(k * f).astype(self.float) for k, f in zip(K, 2 * np.pi / self.L)]
        return G.Vectors([np.broadcast_to(k, self.complex_shape()) for k in K] if broadcast is True else K)

    def get_dealias_filter(self):
        """2/3-rule mask on the local spectral block (applied by the transforms inside their first inverse pass)."""
        s = self.complex_local_slice()
        K = G.sparse_spectral_mesh([k[sl] for k, sl in zip(self._k_axes(), s)])
        return G.two_thirds_mask(K, self.N)

    # ---- transforms
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
    """3D R2C FFT, pencil decomposition, final alignment in x (``pencil.py:885-969``): the spectral block holds all of
    kx, the comm0-th part of ky and the comm1-th part of kz."""

    _kind = D.PENCIL_X

    def __init__(self, N, L, comm, precision, P1=None, communication='Alltoall',
                 padsize=1.5, threads=1,
                 planner_effort=defaultdict(lambda: "FFTW_MEASURE")):
        R2CY.__init__(self, N, L, comm, precision, P1=P1, communication=communication,
                      padsize=padsize, threads=threads, planner_effort=planner_effort)

    def _init_alignment(self, drop):
        self.N2f = self._kz_count(self.N2, self.comm1_rank, self.P2, drop)
        if self.communication == 'Alltoallw':   # pencil.py:911-913: remainder of Nf / P2 on the last rank (the same number)
            self.N2f = self.Nf // self.P2 + (self.Nf % self.P2) * (self.comm1_rank == self.P2 - 1)

    def complex_shape(self):
        return (self.N[0], self.N1[1], self.N2f)

    def complex_local_slice(self):
        c0, c1 = self._grid()
        return (G.whole(self.N[0], 1, None), G.block(self.N1[1], c0), self._kz_slice(self.N2, c1, self.N2f))

    def get_local_mesh(self):
        """Dense (3, ...) coordinate array (``pencil.py:931-943``), unlike the sparse list of the Y alignment."""
        c0, c1 = self._grid()
        sl = (G.block(self.N1[0], c0), G.block(self.N2[1], c1), G.whole(self.N[2]))
        return G.dense_physical_mesh(sl, self.N, self.L, self.float)

    def get_local_wavenumbermesh(self):
        """Dense float mesh of shape (3, N0, N1[1], N2[2]//2): the Nyquist plane is not part of it and kz is taken
        from the full-axis frequency vector, exactly as ``pencil.py:945-957``."""
        c0, c1 = self._grid()
        kx, ky, kz = (G.frequencies(n).astype(int) for n in self.N)
        kzs = slice(c1 * self.N2[2] // 2, (c1 + 1) * self.N2[2] // 2, 1)
        return np.array(np.meshgrid(kx, ky[G.block(self.N1[1], c0)], kz[kzs], indexing='ij'), dtype=self.float)


def R2C(N, L, comm, precision, P1=None, communication="Alltoall", padsize=1.5, threads=1,
        alignment="X", planner_effort=defaultdict(lambda: "FFTW_MEASURE")):
    """Factory of ``pencil.py:1479-1484``."""
    if alignment == 'X':
        return R2CX(N, L, comm, precision, P1, communication, padsize, threads, planner_effort)
    else:
        return R2CY(N, L, comm, precision, P1, communication, padsize, threads, planner_effort)
