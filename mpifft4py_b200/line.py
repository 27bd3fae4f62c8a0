"""2D row decomposition: drop-in for ``mpiFFT4py.line.R2C`` (reference ``mpiFFT4py/line.py:41-340``).

Real rows are split along x (``real_shape = (N0/P, N1)``), spectral columns along ky
(``complex_shape = (N0, Npf)``, the last rank owning the Nyquist column).  The reference's
Nyquist pack trick, Scatter and Send/Recv (``line.py:206,217-223,302-306``) are replaced by an
exchange with one extra column for the last rank; see DESIGN.md (deviation D3) for the one input
class where the pack trick itself is inexact upstream.
"""
from collections import defaultdict

import numpy as np
from numpy.fft import fftfreq

from . import _cdefs as D
from . import _geometry as G
from ._engine import Transform
from .mpibase import datatypes, work_arrays


class R2C(Transform):
    """2D real-to-complex FFT (``fft2``/``ifft2``), row decomposition (``line.py:41-75``): rank r owns rows
    ``[r*Np0, (r+1)*Np0)`` of the real array and ``Npf`` columns of the half spectrum starting at ``r*Np1//2``."""

    def __init__(self, N, L, comm, precision, padsize=1.5, threads=1,
                 planner_effort=defaultdict(lambda: "FFTW_MEASURE")):
        assert len(L) == 2 and len(N) == 2
        self.N, self.L = N, L                      # (L keeps the caller's dtype here, line.py:57)
        self.comm = comm
        self.float, self.complex, self.mpitype = datatypes(precision)
        self.num_processes, self.rank = comm.Get_size(), comm.Get_rank()
        self.padsize, self.threads, self.planner_effort = padsize, threads, planner_effort
        self.Np = N // self.num_processes
        self.Nf = N[1] // 2 + 1
        last = self.rank + 1 == self.num_processes
        self.Npf = self.Np[1] // 2 + (1 if last else 0)   # the last rank also holds the Nyquist column
        self.Nfp = int(padsize * self.N[1] / 2 + 1)
        self.ks = (fftfreq(N[0]) * N[0]).astype(int)
        self.dealias = np.zeros(0)
        self.work_arrays = work_arrays()
        self._create_plan(D.LINE, N, self.num_processes, self.rank, comm=comm)

    def get_N(self):
        return self.N

    # ---- shapes (line.py:77-103, 138-164)
    def real_shape(self):
        return (self.Np[0], self.N[1])

    def complex_shape(self):
        return (self.N[0], self.Npf)

    def global_real_shape(self):
        return (self.N[0], self.N[1])

    def global_complex_shape(self):
        return (self.N[0], self.Nf)

    def global_complex_shape_padded(self):
        return (int(self.padsize * self.N[0]), int(self.padsize * self.N[1] / 2 + 1))

    def real_shape_padded(self):
        return G.padded(self.real_shape(), self.padsize)

    # shapes of the 3/2-rule intermediates (line.py:146-156; public methods upstream): the padded half spectrum of the
    # local rows, the local rows before the y pad, and this rank's spectral columns padded in x
    def complex_padded_xy(self):
        return (int(self.padsize * self.Np[0]), int(self.padsize * self.N[1] / 2 + 1))

    def complex_shape_padded_01(self):
        return (int(self.padsize * self.Np[0]), self.Nf)

    def complex_padded_x(self):
        return (int(self.padsize * self.N[0]), self.Npf)

    def work_shape(self, dealias):
        return self.real_shape_padded() if dealias == '3/2-rule' else self.real_shape()

    # ---- slices (line.py:93-103: the slices over whole axes carry no step)
    def real_local_slice(self, padsize=1):
        return (G.block(self.Np[0], self.rank, padsize), G.whole(self.N[1], padsize, None))

    def _ky_slice(self):
        lo = self.rank * self.Np[1] // 2
        return slice(lo, lo + self.Npf, 1)

    def complex_local_slice(self):
        return (G.whole(self.N[0], 1, None), self._ky_slice())

    # ---- meshes (line.py:105-136)
    def get_local_mesh(self):
        return G.dense_physical_mesh((G.block(self.Np[0], self.rank), G.whole(self.N[1])), self.N, self.L, self.float)

    def get_local_wavenumbermesh(self, scaled=True, broadcast=False, eliminate_highest_freq=False):
        """Sparse wavenumber mesh of the local spectral block; scaled by 2 pi / L by DEFAULT here (unlike the 3D classes)."""
        ks = [G.frequencies(self.N[0]), G.frequencies(self.N[1], half=True)]
        if eliminate_highest_freq:
            G.drop_nyquist(ks, self.N)
        K = G.sparse_spectral_mesh((ks[0], ks[1][self._ky_slice()]))
        if scaled is True:
            for k, f in zip(K, 2 * np.pi / self.L):
                k *= f
        return G.Vectors([np.broadcast_to(k, self.complex_shape()) for k in K] if broadcast is True else K)

    def get_dealias_filter(self):
        """``line.py:131-136``: the mask tests the default (scaled) wavenumbers, i.e. assumes the 2 pi-periodic box; the
        engine's fused mask uses the mode numbers, the same thing there."""
        return G.two_thirds_mask(self.get_local_wavenumbermesh(), self.N)

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
