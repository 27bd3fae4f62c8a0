"""Oracle for ``mpiFFT4py.slab.R2C`` (reference ``mpiFFT4py/slab.py:49-536``).

TEST INFRASTRUCTURE.  Ranks are list entries: ``u[r]`` is rank r's local real block of shape
``(Np0, N1, N2)``; ``fu[r]`` its spectral block ``(N0, Np1, Nf)``.
"""
import numpy as np

from .common import F, alltoall, dealias_mask, dtypes, pad_copy, trunc_fold


class Geometry(object):
    """Integer bookkeeping of ``slab.py:67-144,487-514``."""

    def __init__(self, N, P, padsize=1.5):
        self.N = np.asarray(N, dtype=int)
        assert len(self.N) == 3
        self.P = int(P)
        self.padsize = padsize
        N0 = int(self.N[0])
        if self.P not in [2 ** i for i in range(int(np.log2(N0)) + 1)]:  # slab.py:89-91
            raise IOError("Number of cpus must be a power of two <= N[0]")
        self.Np = self.N // self.P
        self.Nf = int(self.N[2]) // 2 + 1
        self.Nfp = int(padsize * self.N[2] // 2 + 1)

    def real_shape(self):
        return (int(self.Np[0]), int(self.N[1]), int(self.N[2]))

    def complex_shape(self):
        return (int(self.N[0]), int(self.Np[1]), self.Nf)

    def real_shape_padded(self):
        p = self.padsize
        return (int(p * self.Np[0]), int(p * self.N[1]), int(p * self.N[2]))

    def global_complex_shape(self, padsize=1.):
        return (int(padsize * self.N[0]), int(padsize * self.N[1]), int(padsize * self.N[2] // 2 + 1))

    def real_local_slice(self, rank, padsize=1):
        return (slice(int(padsize * rank * self.Np[0]), int(padsize * (rank + 1) * self.Np[0]), 1),
                slice(0, int(padsize * self.N[1]), 1),
                slice(0, int(padsize * self.N[2]), 1))

    def complex_local_slice(self, rank):
        return (slice(0, int(self.N[0]), 1),
                slice(rank * int(self.Np[1]), (rank + 1) * int(self.Np[1]), 1),
                slice(0, self.Nf, 1))

    def wavenumbers(self, rank, dtype=np.float64):
        """``slab.py:140-144``."""
        s = self.complex_local_slice(rank)
        return (np.fft.fftfreq(self.N[0], 1. / self.N[0]).astype(dtype),
                np.fft.fftfreq(self.N[1], 1. / self.N[1])[s[1]].astype(dtype),
                np.fft.rfftfreq(self.N[2], 1. / self.N[2]).astype(dtype))

    def mask(self, rank):
        kx, ky, kz = self.wavenumbers(rank)
        return dealias_mask(np.meshgrid(kx, ky, kz, indexing="ij", sparse=True), self.N)


def fftn(u, N, P, dealias=None, padsize=1.5, precision="double"):
    """Forward transform of all ranks (``slab.py:349-485``).  ``u``: list of P local arrays."""
    assert dealias in ("3/2-rule", "2/3-rule", "None", None)
    g = Geometry(N, P, padsize)
    rt, ct = dtypes(precision)
    N0, N1, N2 = (int(n) for n in g.N)
    Np0, Np1 = int(g.Np[0]), int(g.Np[1])
    Nf = g.Nf
    padded = dealias == "3/2-rule"

    if P == 1:  # slab.py:366-387
        a = np.asarray(u[0], dtype=rt)
        if not padded:
            assert a.shape == g.real_shape()
            c = F.fft(F.fft(F.rfft(a, 2), 1), 0)
            return [c.astype(ct)]
        assert a.shape == g.real_shape_padded()
        fp = F.fft(F.fft(F.rfft(a, 2), 1), 0)
        fu = np.zeros(g.complex_shape(), dtype=ct)
        h0, h1 = N0 // 2, N1 // 2
        fu[:h0 + 1, :h1 + 1] = fp[:h0 + 1, :h1 + 1, :Nf]
        fu[:h0 + 1, h1:] += fp[:h0 + 1, -h1:, :Nf]
        fu[h0:, :h1 + 1] += fp[-h0:, :h1 + 1, :Nf]
        fu[h0:, h1:] += fp[-h0:, -h1:, :Nf]
        fu /= padsize ** 3
        return [fu]

    if padded:
        assert P <= N0 // 2  # slab.py:446
    pNp0 = int(padsize * Np0) if padded else Np0
    send = []
    for r in range(P):
        a = np.asarray(u[r], dtype=rt)
        assert a.shape == (g.real_shape_padded() if padded else g.real_shape())
        t = F.fft(F.rfft(a, 2), 1)  # rfft2 over axes (1, 2): slab.py:434,456
        if padded:  # copy_from_padded(axis=1): slab.py:459,529-533
            tt = np.zeros((pNp0, N1, Nf), dtype=ct)
            t = trunc_fold(t[:, :, :Nf], tt, N1, 1)
        send.append([t[:, j * Np1:(j + 1) * Np1, :] for j in range(P)])  # subarraysB slab.py:206-209
    recv = alltoall([list(range(P))], send)
    out = []
    for r in range(P):
        U = np.concatenate(recv[r], axis=0)  # subarraysA: rows i*Np0.. from rank i, slab.py:202-205
        U = F.fft(U, 0)  # slab.py:442,476
        if padded:  # slab.py:479-483
            fu = np.zeros(g.complex_shape(), dtype=ct)
            fu = trunc_fold(U, fu, N0, 0)
            fu /= padsize ** 3
        else:
            fu = U
        out.append(fu.astype(ct))
    return out


def ifftn(fu, N, P, dealias=None, padsize=1.5, precision="double"):
    """Inverse transform of all ranks (``slab.py:214-346``).  ``fu``: list of P spectral blocks."""
    assert dealias in ("3/2-rule", "2/3-rule", "None", None)
    g = Geometry(N, P, padsize)
    rt, ct = dtypes(precision)
    N0, N1, N2 = (int(n) for n in g.N)
    Np0, Np1 = int(g.Np[0]), int(g.Np[1])
    Nf = g.Nf
    padded = dealias == "3/2-rule"

    fu = [np.asarray(f, dtype=ct) for f in fu]
    if dealias == "2/3-rule":  # slab.py:237-245 (works on a copy; input is never modified)
        fu = [f * g.mask(r) for r, f in enumerate(fu)]

    if P == 1:  # slab.py:247-268
        if not padded:
            a = F.irfft(F.ifft(F.ifft(fu[0], 0), 1), 2)
            return [a.astype(rt)]
        f = fu[0] * padsize ** 3
        fp = np.zeros(g.global_complex_shape(padsize), dtype=ct)
        h0, h1 = N0 // 2, N1 // 2
        fp[:h0, :h1, :Nf] = f[:h0, :h1]
        fp[:h0, -h1:, :Nf] = f[:h0, h1:]
        fp[-h0:, :h1, :Nf] = f[h0:, :h1]
        fp[-h0:, -h1:, :Nf] = f[h0:, -h1:]
        a = F.irfft(F.ifft(F.ifft(fp, 0), 1), 2)
        return [a.astype(rt)]

    if padded:
        assert P <= N0 // 2  # slab.py:311
    p = padsize if padded else 1
    pN0, pNp0, pN1 = int(p * N0), int(p * Np0), int(p * N1)
    send = []
    for r in range(P):
        f = fu[r]
        if padded:  # slab.py:320
            fp = np.zeros((pN0, Np1, Nf), dtype=ct)
            f = pad_copy((f * padsize ** 3).astype(ct), fp, N0, 0)
        U = F.ifft(f, 0)  # slab.py:275,321
        send.append([U[j * pNp0:(j + 1) * pNp0] for j in range(P)])  # subarraysA
    recv = alltoall([list(range(P))], send)
    out = []
    for r in range(P):
        U = np.concatenate(recv[r], axis=1)  # subarraysB: y columns i*Np1.. from rank i
        if padded:  # slab.py:337-343
            U2 = np.zeros((pNp0, pN1, Nf), dtype=ct)
            U2 = pad_copy(U, U2, N1, 1)
            U2 = F.ifft(U2, 1)
            U3 = np.zeros((pNp0, pN1, g.Nfp), dtype=ct)
            U3[:, :, :Nf] = U2
            a = F.irfft(U3, 2)
        else:  # irfft2 over axes (1, 2): slab.py:306
            a = F.irfft(F.ifft(U, 1), 2)
        out.append(a.astype(rt))
    return out


# ------------------------------------------------------------------------------------------------
# slab.C2C (``slab.py:538-825``): the same decomposition with a complex-to-complex z transform.
# ------------------------------------------------------------------------------------------------
class GeometryC2C(Geometry):
    """``slab.py:556-586``: R2C's shapes with the z extent ``Nf = N[2]``."""

    def __init__(self, N, P, padsize=1.5):
        Geometry.__init__(self, N, P, padsize)
        self.Nf = int(self.N[2])
        self.Nfp = int(padsize * self.N[2])

    original_shape = Geometry.real_shape
    original_shape_padded = Geometry.real_shape_padded
    transformed_shape = Geometry.complex_shape
    original_local_slice = Geometry.real_local_slice
    transformed_local_slice = Geometry.complex_local_slice

    def mask(self, rank):
        """Intended 2/3-rule mask.  (The reference's own ``get_dealias_filter`` is inherited from R2C
        and builds an rfft-sized z extent, so ``C2C.ifftn(dealias='2/3-rule')`` raises a broadcast
        ``ValueError`` upstream -- untested there; here the mask simply spans the full kz range.)"""
        s = self.complex_local_slice(rank)
        kx = np.fft.fftfreq(self.N[0], 1. / self.N[0])
        ky = np.fft.fftfreq(self.N[1], 1. / self.N[1])[s[1]]
        kz = np.fft.fftfreq(self.N[2], 1. / self.N[2])
        return dealias_mask(np.meshgrid(kx, ky, kz, indexing="ij", sparse=True), self.N)


def _trunc_keep_neg(fp, fu, N, axis):
    """Plain truncation that keeps mode -N/2 (``slab.py:735-738,796-797`` and the ``ks`` gather)."""
    h = int(N) // 2
    a = [slice(None)] * fu.ndim
    b = [slice(None)] * fu.ndim
    a[axis] = slice(0, h)
    fu[tuple(a)] = fp[tuple(a)]
    a[axis] = slice(h, None)
    b[axis] = slice(fp.shape[axis] - h, None)
    fu[tuple(a)] = fp[tuple(b)]
    return fu


def c2c_fftn(u, N, P, dealias=None, padsize=1.5, precision="double"):
    """``C2C.fftn`` of all ranks (``slab.py:700-800``)."""
    assert dealias in ("3/2-rule", "2/3-rule", "None", None)
    g = GeometryC2C(N, P, padsize)
    rt, ct = dtypes(precision)
    N0, N1, N2 = (int(n) for n in g.N)
    Np0, Np1 = int(g.Np[0]), int(g.Np[1])
    padded = dealias == "3/2-rule"
    if P == 1:  # slab.py:721-740
        a = np.asarray(u[0], dtype=ct)
        if not padded:
            return [F.fft(F.fft(F.fft(a, 2), 1), 0).astype(ct)]
        fp = F.fft(F.fft(F.fft(a, 2), 1), 0)
        t2 = _trunc_keep_neg(fp, np.zeros(fp.shape[:2] + (N2,), dtype=ct), N2, 2)
        t1 = _trunc_keep_neg(t2, np.zeros((fp.shape[0], N1, N2), dtype=ct), N1, 1)
        fu = _trunc_keep_neg(t1, np.zeros((N0, N1, N2), dtype=ct), N0, 0)
        return [(fu / padsize ** 3).astype(ct)]
    pNp0 = int(padsize * Np0) if padded else Np0
    send = []
    for r in range(P):
        a = np.asarray(u[r], dtype=ct)
        t = F.fft(F.fft(a, 2), 1)  # fft2 over axes (1, 2): slab.py:748,778
        if padded:  # copy_from_padded(axis=1): fold in y and z, slab.py:816-823
            t = trunc_fold(t, np.zeros(t.shape[:2] + (N2,), dtype=ct), N2, 2)
            t = trunc_fold(t, np.zeros((pNp0, N1, N2), dtype=ct), N1, 1)
        send.append([t[:, j * Np1:(j + 1) * Np1, :] for j in range(P)])
    recv = alltoall([list(range(P))], send)
    out = []
    for r in range(P):
        U = F.fft(np.concatenate(recv[r], axis=0), 0)  # slab.py:771,791
        if padded:  # slab.py:794-798: plain truncation in x
            U = _trunc_keep_neg(U, np.zeros(g.complex_shape(), dtype=ct), N0, 0) / padsize ** 3
        out.append(U.astype(ct))
    return out


def c2c_ifftn(fu, N, P, dealias=None, padsize=1.5, precision="double"):
    """``C2C.ifftn`` of all ranks (``slab.py:587-698``)."""
    assert dealias in ("3/2-rule", "2/3-rule", "None", None)
    g = GeometryC2C(N, P, padsize)
    rt, ct = dtypes(precision)
    N0, N1, N2 = (int(n) for n in g.N)
    Np0, Np1 = int(g.Np[0]), int(g.Np[1])
    padded = dealias == "3/2-rule"
    fu = [np.asarray(f, dtype=ct) for f in fu]
    if dealias == "2/3-rule":
        fu = [f * g.mask(r) for r, f in enumerate(fu)]
    p = padsize if padded else 1
    pN0, pNp0, pN1, pN2 = int(p * N0), int(p * Np0), int(p * N1), int(p * N2)
    if P == 1:  # slab.py:609-634
        f = fu[0]
        if padded:
            f = pad_copy(f * padsize ** 3, np.zeros((pN0, N1, N2), dtype=ct), N0, 0)
            f = pad_copy(f, np.zeros((pN0, pN1, N2), dtype=ct), N1, 1)
            f = pad_copy(f, np.zeros((pN0, pN1, pN2), dtype=ct), N2, 2)
        return [F.ifft(F.ifft(F.ifft(f, 0), 1), 2).astype(ct)]
    send = []
    for r in range(P):
        f = fu[r]
        if padded:  # slab.py:675
            f = pad_copy((f * padsize ** 3).astype(ct), np.zeros((pN0, Np1, N2), dtype=ct), N0, 0)
        U = F.ifft(f, 0)  # slab.py:646,676
        send.append([U[j * pNp0:(j + 1) * pNp0] for j in range(P)])
    recv = alltoall([list(range(P))], send)
    out = []
    for r in range(P):
        U = np.concatenate(recv[r], axis=1)
        if padded:  # slab.py:683-692
            U = F.ifft(pad_copy(U, np.zeros((pNp0, pN1, N2), dtype=ct), N1, 1), 1)
            U = F.ifft(pad_copy(U, np.zeros((pNp0, pN1, pN2), dtype=ct), N2, 2), 2)
        else:  # ifft2 over axes (1, 2): slab.py:662
            U = F.ifft(F.ifft(U, 1), 2)
        out.append(U.astype(ct))
    return out
