"""Index and mesh arithmetic shared by the slab, pencil and line classes.

The three decompositions of the reference differ in which axis is cut by which communicator, but every shape,
slice, mesh and mask they expose is assembled from the same few pieces: an equal block partition of an axis
(optionally stretched by the 3/2-rule pad factor), the FFT frequency vector of an axis, a sparse outer mesh, and
the 2/3-rule band test.  They are written once here; the classes only say which block of which axis they own.

Results are value- and dtype-identical to the reference's methods (``slab.py:98-197``, ``pencil.py:248-349,
945-969``, ``line.py:77-136``) -- tests/test_host_api.py compares against the unmodified reference -- including its
quirks: which slices carry ``step=1`` and which ``None``, integer versus float wavenumbers, and the two different
roundings of the physical mesh (``index * L / N`` for the sparse meshes, ``index * (L / N)`` for the dense ones).
"""
import numpy as np
from numpy.fft import fftfreq, rfftfreq


def block(width, index, pad=1, step=1):
    """Slice of block ``index`` of an axis cut into blocks of ``width`` points, stretched by ``pad``."""
    return slice(int(pad * index * width), int(pad * (index + 1) * width), step)


def whole(n, pad=1, step=1):
    """Slice of a whole axis of ``n`` points (``step`` is 1 or None exactly where the reference has it)."""
    return slice(0, int(pad * n), step) if step is not None else slice(0, int(pad * n))


def padded(shape, pad):
    return tuple(int(pad * s) for s in shape)


def frequencies(n, half=False):
    """Wavenumbers 0, 1, ..., -1 of an axis of n points (float64); ``half``: the non-negative half of a real axis."""
    return rfftfreq(n, 1. / n) if half else fftfreq(n, 1. / n)


def drop_nyquist(ks, N):
    """Zero the highest frequency of every even axis, in place (``eliminate_highest_freq``)."""
    for k, n in zip(ks, N):
        if n % 2 == 0:
            k[n // 2] = 0


def sparse_physical_mesh(slices, N, L, dtype, shape):
    """Broadcast views of the local grid coordinates, one per axis: ``(index * L / N)`` rounded to ``dtype``."""
    axes = list(np.ogrid[tuple(slices)])
    axes = [(a * L[i] / N[i]).astype(dtype) for i, a in enumerate(axes)]
    return [np.broadcast_to(a, shape) for a in axes]


def dense_physical_mesh(slices, N, L, dtype):
    """Dense ``(dims, ...)`` array of the local grid coordinates: indices cast to ``dtype``, times ``L / N``."""
    X = np.mgrid[tuple(slices)].astype(dtype)
    for i in range(X.shape[0]):
        X[i] *= L[i] / N[i]
    return X


class Vectors(list):
    """The list of per-axis arrays the mesh methods return, usable as ONE operand of ``*`` as well: ``a * K`` is the
    stack ``[a * K[0], a * K[1], ...]``.  The reference's callers write exactly that (``P_hat*K`` with the sparse
    wavenumber list, ``demo/spectral_dns_solver.py:75``); numpy used to build it from the ragged list on its own and
    refuses to since 1.24, which is why the upstream demo no longer runs on a current numpy -- with this class it
    does, unedited.  Everything else (indexing, iteration, ``len``, ``isinstance(K, list)``) is a plain list."""

    __array_ufunc__ = None  # ndarray.__mul__(K) then defers to K.__rmul__ instead of trying np.asarray(K)

    def __mul__(self, other):
        return np.array([k * other for k in self])

    def __rmul__(self, other):
        return np.array([other * k for k in self])


def sparse_spectral_mesh(ks, shape=None):
    """Outer (sparse) mesh of the per-axis wavenumber vectors; broadcast to ``shape`` when given."""
    K = list(np.meshgrid(*ks, indexing='ij', sparse=True))
    return K if shape is None else [np.broadcast_to(k, shape) for k in K]


def two_thirds_mask(K, N):
    """uint8 mask of the modes the 2/3-rule keeps: |k_i| < 2/3 (N_i // 2 + 1) on every axis."""
    kmax = 2. / 3. * (np.asarray(N) // 2 + 1)
    keep = abs(K[0]) < kmax[0]
    for i in range(1, len(K)):
        keep = keep * (abs(K[i]) < kmax[i])
    return np.array(keep, dtype=np.uint8)


def balanced_grid(P):
    """Process grid of ``MPI.Compute_dims(P, 2)``: the most balanced factor pair, larger factor first."""
    a = int(P ** 0.5)
    while P % a:
        a -= 1
    return P // a, a
