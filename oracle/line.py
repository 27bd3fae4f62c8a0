"""Oracle for ``mpiFFT4py.line.R2C`` (reference ``mpiFFT4py/line.py:14-340``), 2D row decomposition.

TEST INFRASTRUCTURE.  ``u[r]``: ``(Np0, N1)`` real rows of rank r; ``fu[r]``: ``(N0, Npf)`` spectral
columns, ``Npf = Np1//2`` (+1 on the last rank, which owns the Nyquist column) -- ``line.py:69-71``.

``fft2``/``ifft2`` restate the reference step by step INCLUDING its Nyquist pack trick
(``line.py:206,217-223`` and ``swap_Nq`` ``:26-39``): the real Nyquist column rides in the
imaginary part of the k=0 column through the exchange and is separated afterwards by Hermitian
symmetry.  For the unpadded transform this is exact.  For the 3/2-rule forward transform the
column ``ky = N1/2`` of the *padded* spectrum is in general not the transform of a real
sequence, so the reference's separation is only exact for inputs whose padded spectrum came
from a real unpadded field (what ``tests/test_FFT.py:112-156`` feeds it).  ``exact=True``
selects the intended semantics (plain truncation of that column), which is what the CUDA
engine implements; both agree on reference-test-style inputs (deviation D3).
"""
import numpy as np

from .common import F, alltoall, dealias_mask, dtypes, pad_copy, trunc_fold


class Geometry(object):
    """``line.py:55-105``."""

    def __init__(self, N, P, padsize=1.5):
        self.N = np.asarray(N, dtype=int)
        assert len(self.N) == 2
        self.P = int(P)
        self.padsize = padsize
        self.Np = self.N // self.P
        self.Nf = int(self.N[1]) // 2 + 1
        self.Nfp = int(padsize * self.N[1] / 2 + 1)
        self.ks = (np.fft.fftfreq(int(self.N[0])) * int(self.N[0])).astype(int)

    def Npf(self, rank):
        return int(self.Np[1]) // 2 + 1 if rank + 1 == self.P else int(self.Np[1]) // 2

    def real_shape(self):
        return (int(self.Np[0]), int(self.N[1]))

    def complex_shape(self, rank):
        return (int(self.N[0]), self.Npf(rank))

    def real_shape_padded(self):
        return (int(self.padsize * self.Np[0]), int(self.padsize * self.N[1]))

    def global_complex_shape(self):
        return (int(self.N[0]), self.Nf)

    def real_local_slice(self, rank, padsize=1):
        return (slice(int(padsize * rank * self.Np[0]), int(padsize * (rank + 1) * self.Np[0]), 1),
                slice(0, int(padsize * self.N[1])))

    def complex_local_slice(self, rank):
        return (slice(0, int(self.N[0])),
                slice(rank * int(self.Np[1]) // 2, rank * int(self.Np[1]) // 2 + self.Npf(rank), 1))

    def mask(self, rank):
        """``line.py:114-136`` (unscaled wavenumbers: scaling by 2pi/L = 1 for L = 2pi)."""
        s = self.complex_local_slice(rank)
        kx = np.fft.fftfreq(self.N[0], 1. / self.N[0])
        ky = np.fft.rfftfreq(self.N[1], 1. / self.N[1])[s[1]]
        return dealias_mask(np.meshgrid(kx, ky, indexing="ij", sparse=True), self.N)


def _separate(f, M):
    """``swap_Nq`` (``line.py:26-39``): from ``f = FFT(a + i b)`` of two real sequences return
    ``(FFT(a), FFT(b))`` using Hermitian symmetry; ends take ``.real``/``.imag``."""
    h = M // 2
    fa = np.empty_like(f)
    fb = np.empty_like(f)
    fa[0] = f[0].real
    fa[1:h] = 0.5 * (f[1:h] + np.conj(f[:h:-1]))
    fa[h] = f[h].real
    fa[h + 1:] = np.conj(fa[h - 1:0:-1])
    fb[0] = f[0].imag
    fb[1:h] = -0.5j * (f[1:h] - np.conj(f[:h:-1]))
    fb[h] = f[h].imag
    fb[h + 1:] = np.conj(fb[h - 1:0:-1])
    return fa, fb


def fft2(u, N, P, dealias=None, padsize=1.5, precision="double", exact=False):
    """Forward 2D transform of all ranks (``line.py:179-260``)."""
    assert dealias in ("3/2-rule", "2/3-rule", "None", None)
    g = Geometry(N, P, padsize)
    rt, ct = dtypes(precision)
    N0, N1 = int(g.N[0]), int(g.N[1])
    Nf = g.Nf
    padded = dealias == "3/2-rule"

    if P == 1:  # line.py:182-191
        a = np.asarray(u[0], dtype=rt)
        if not padded:
            return [F.fft(F.rfft(a, 1), 0).astype(ct)]
        fp = F.fft(F.rfft((a / padsize ** 2).astype(rt), 1), 0)
        return [fp[g.ks, :Nf].astype(ct)]  # fancy index: NO Nyquist fold at P == 1

    h = int(g.Np[1]) // 2
    pNp0 = int(padsize * g.Np[0]) if padded else int(g.Np[0])
    send, nyq = [], []
    for r in range(P):
        a = np.asarray(u[r], dtype=rt)
        assert a.shape == (g.real_shape_padded() if padded else g.real_shape())
        if padded:
            t = F.rfft((a / padsize).astype(rt), 1)[:, :Nf].copy()  # line.py:236-237
        else:
            t = F.rfft(a, 1)  # line.py:205
        if exact:
            nyq.append(t[:, -1].copy())
        else:
            t[:, 0] += 1j * t[:, -1]  # line.py:206,238
        send.append([t[:, j * h:(j + 1) * h] for j in range(P)])  # transpose_x, line.py:14-18
    recv = alltoall([list(range(P))], send)
    out = []
    for r in range(P):
        U = np.concatenate(recv[r], axis=0)
        if padded:  # line.py:245-249
            U = F.fft((U / padsize).astype(ct), 0)
            V = np.zeros((N0, h), dtype=ct)
            U = trunc_fold(U, V, N0, 0)
        else:
            U = F.fft(U, 0)  # line.py:213
        fu = np.zeros(g.complex_shape(r), dtype=ct)
        fu[:, :h] = U
        out.append(fu)
    if exact:
        col = np.concatenate(nyq, axis=0)
        if padded:
            col = F.fft((col / padsize).astype(ct), 0)
            V = np.zeros((N0,), dtype=ct)
            col = trunc_fold(col, V, N0, 0)
        else:
            col = F.fft(col, 0)
        out[P - 1][:, -1] = col
    else:  # rank 0 separates, sends the Nyquist column to the last rank: line.py:217-223
        fa, fb = _separate(out[0][:, 0].copy(), N0)
        out[0][:, 0] = fa
        out[P - 1][:, -1] = fb
    return out


def ifft2(fu, N, P, dealias=None, padsize=1.5, precision="double"):
    """Inverse 2D transform of all ranks (``line.py:262-340``)."""
    assert dealias in ("3/2-rule", "2/3-rule", "None", None)
    g = Geometry(N, P, padsize)
    rt, ct = dtypes(precision)
    N0, N1 = int(g.N[0]), int(g.N[1])
    Nf = g.Nf
    padded = dealias == "3/2-rule"
    fu = [np.asarray(f, dtype=ct) for f in fu]
    if dealias == "2/3-rule":  # line.py:268-272
        fu = [f * g.mask(r) for r, f in enumerate(fu)]

    if P == 1:  # line.py:274-283
        if not padded:
            return [F.irfft(F.ifft(fu[0], 0), 1).astype(rt)]
        fp = np.zeros((int(padsize * N0), g.Nfp), dtype=ct)
        fp[g.ks, :Nf] = fu[0]
        return [F.irfft(F.ifft((fp * padsize ** 2).astype(ct), 0), 1).astype(rt)]

    h = int(g.Np[1]) // 2
    p = padsize if padded else 1
    pN0, pNp0 = int(p * N0), int(p * g.Np[0])
    send = []
    nyq = None
    for r in range(P):
        f = fu[r]
        if padded:  # copy_to_padded_x line.py:156-159,322-323
            V = np.zeros((pN0, f.shape[1]), dtype=ct)
            f = pad_copy(f, V, N0, 0)
        U = F.ifft(f, 0)  # line.py:295,323
        if r == P - 1:
            nyq = U[:, -1].copy()  # line.py:302-303
        send.append([U[j * pNp0:(j + 1) * pNp0, :h] for j in range(P)])
    recv = alltoall([list(range(P))], send)
    out = []
    for r in range(P):
        T = np.zeros((pNp0, Nf), dtype=ct)
        T[:, :-1] = np.concatenate(recv[r], axis=1)  # transpose_y, line.py:20-24
        T[:, -1] = nyq[r * pNp0:(r + 1) * pNp0]  # Scatter from the last rank, line.py:305-306
        if padded:  # line.py:336-338
            V = np.zeros((pNp0, g.Nfp), dtype=ct)
            V[:, :Nf] = T
            T = (V * padsize ** 2).astype(ct)
        out.append(F.irfft(T, 1).astype(rt))
    return out
