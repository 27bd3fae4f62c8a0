"""Oracle for ``mpiFFT4py.pencil.R2CX`` / ``R2CY`` (reference ``mpiFFT4py/pencil.py:76-1484``).

TEST INFRASTRUCTURE.  World rank r sits at ``c0 = r % P1`` (rank inside ``comm0``, the group of
P1 consecutive ranks) and ``c1 = r // P1`` (rank inside ``comm1``, the group with equal
``r % P1``) -- ``pencil.py:192-195``.  Real blocks are ``(N1[0], N2[1], N[2])`` with
``N1 = N // P1``, ``N2 = N // P2``.

The data path restated here is the reference's ``'Alltoallw'`` one (``pencil.py:483-507,
730-754, 1082-1105, 1312-1337`` and the padded twins).  ``'Alltoall'`` (pack the Nyquist plane
into k=0, equal-size exchanges, Send/Recv/Scatter of one plane) produces the same arrays in the
same layout to rounding -- the golden vectors of both variants pin that.  ``'AlltoallN'`` drops
the Nyquist plane (``pencil.py:197-199,908-910``; inverse zeroes it ``:428,1045``).
"""
import numpy as np

from .common import F, alltoall, dealias_mask, dtypes, pad_copy, trunc_fold


def compute_dims(P):
    """MPI_Dims_create(P, 2): balanced, non-increasing (8 -> [4, 2], 4 -> [2, 2])."""
    best = (P, 1)
    for a in range(1, int(P ** 0.5) + 1):
        if P % a == 0:
            best = (P // a, a)
    return best


def _chunks(total, parts, drop_remainder=False):
    """``pencil.py:80-90`` _distribution: equal chunks, the remainder (1) goes to the LAST."""
    q, r = total // parts, total % parts
    out = []
    for i in range(parts):
        n = q + (r if (i + 1 == parts and not drop_remainder) else 0)
        out.append((n, q * i))
    return out


class Geometry(object):
    """``pencil.py:167-216`` (Y) and ``:903-969`` (X) bookkeeping."""

    def __init__(self, N, P, alignment="X", P1=None, communication="Alltoall", padsize=1.5):
        self.N = np.asarray(N, dtype=int)
        assert len(self.N) == 3
        self.P = int(P)
        assert self.P > 1  # pencil.py:176
        self.alignment = alignment
        self.communication = communication
        self.padsize = padsize
        if P1 is None:
            P1, P2 = compute_dims(self.P)
        else:
            P2 = self.P // P1
        self.P1, self.P2 = int(P1), int(P2)
        if self.P % 2 != 0:
            raise IOError("Number of cpus must be even")
        if (self.P1 % 2 != 0) or (self.P2 % 2 != 0):  # pencil.py:204-205
            raise IOError("Number of cpus in each direction must be even power of 2")
        self.N1 = self.N // self.P1
        self.N2 = self.N // self.P2
        self.Nf = int(self.N[2]) // 2 + 1
        self.dropN = communication == "AlltoallN"

    def coords(self, rank):
        return rank % self.P1, rank // self.P1

    def comm0_groups(self):
        return [[c1 * self.P1 + c0 for c0 in range(self.P1)] for c1 in range(self.P2)]

    def comm1_groups(self):
        return [[c1 * self.P1 + c0 for c1 in range(self.P2)] for c0 in range(self.P1)]

    def zparts(self):
        """(number of z-chunks, per-chunk (len, start)) for this alignment."""
        parts = self.P1 if self.alignment == "Y" else self.P2
        return _chunks(self.Nf, parts, drop_remainder=self.dropN)

    def kzlen(self, rank):
        c0, c1 = self.coords(rank)
        return self.zparts()[c0 if self.alignment == "Y" else c1][0]

    def real_shape(self):
        return (int(self.N1[0]), int(self.N2[1]), int(self.N[2]))

    def real_shape_padded(self):
        p = self.padsize
        return (int(p * self.N1[0]), int(p * self.N2[1]), int(p * self.N[2]))

    def complex_shape(self, rank):
        if self.alignment == "Y":
            return (int(self.N2[0]), int(self.N[1]), self.kzlen(rank))
        return (int(self.N[0]), int(self.N1[1]), self.kzlen(rank))

    def real_local_slice(self, rank, padsize=1):
        c0, c1 = self.coords(rank)
        return (slice(int(padsize * c0 * self.N1[0]), int(padsize * (c0 + 1) * self.N1[0]), 1),
                slice(int(padsize * c1 * self.N2[1]), int(padsize * (c1 + 1) * self.N2[1]), 1),
                slice(0, int(padsize * self.N[2])))

    def complex_local_slice(self, rank):
        c0, c1 = self.coords(rank)
        if self.alignment == "Y":  # pencil.py:271-276
            return (slice(c1 * int(self.N2[0]), (c1 + 1) * int(self.N2[0]), 1),
                    slice(0, int(self.N[1])),
                    slice(c0 * int(self.N1[2]) // 2, c0 * int(self.N1[2]) // 2 + self.kzlen(rank), 1))
        return (slice(0, int(self.N[0])),  # pencil.py:937-943
                slice(c0 * int(self.N1[1]), (c0 + 1) * int(self.N1[1]), 1),
                slice(c1 * int(self.N2[2]) // 2, c1 * int(self.N2[2]) // 2 + self.kzlen(rank), 1))

    def mask(self, rank):
        """Intended 2/3-rule mask on the local spectral block (R2CY ``pencil.py:321-349``; for
        R2CX the reference's own mask/ifftn is broken -- SURVEY.md 8a-Q1 -- so the slab/R2CY
        semantics are used)."""
        s = self.complex_local_slice(rank)
        kx = np.fft.fftfreq(self.N[0], 1. / self.N[0])[s[0]]
        ky = np.fft.fftfreq(self.N[1], 1. / self.N[1])[s[1]]
        kz = np.fft.rfftfreq(self.N[2], 1. / self.N[2])[s[2]]
        return dealias_mask(np.meshgrid(kx, ky, kz, indexing="ij", sparse=True), self.N)


def _split(a, axis, chunks):
    idx = [slice(None)] * a.ndim
    out = []
    for n, s in chunks:
        idx[axis] = slice(s, s + n)
        out.append(a[tuple(idx)])
    return out


def fftn(u, N, P, alignment="X", P1=None, communication="Alltoall", dealias=None, padsize=1.5,
         precision="double"):
    """Forward transform of all ranks (Y: ``pencil.py:634-883``; X: ``:1228-1477``)."""
    assert dealias in ("3/2-rule", "2/3-rule", "None", None)
    g = Geometry(N, P, alignment, P1, communication, padsize)
    rt, ct = dtypes(precision)
    padded = dealias == "3/2-rule"
    p = padsize if padded else 1
    N0, N1_, N2_ = (int(n) for n in g.N)
    Nf = g.Nf
    zc = g.zparts()
    # First the z transform on every rank (pencil.py:738,1321; padded :859-861,1452-1454)
    stage = []
    for r in range(g.P):
        a = np.asarray(u[r], dtype=rt)
        assert a.shape == (g.real_shape_padded() if padded else g.real_shape())
        t = F.rfft(a, 2)[:, :, :Nf]  # copy_from_padded_z: plain truncation, pencil.py:365-367
        stage.append(t)

    if alignment == "Y":
        # comm0: split z over P1, gather x  (subarrays2B -> 2A, pencil.py:226-245,741-743)
        send = [_split(t, 2, zc) for t in stage]
        recv = alltoall(g.comm0_groups(), send)
        stage = []
        for r in range(g.P):
            U = F.fft(np.concatenate(recv[r], axis=0), 0)  # pencil.py:745,868
            if padded:  # copy_from_padded_x pencil.py:369-373,871
                V = np.zeros((N0,) + U.shape[1:], dtype=ct)
                U = trunc_fold(U, V, N0, 0)
            stage.append(U)
        # comm1: split x over P2, gather y (subarrays1B -> 1A, pencil.py:748-750)
        xc = _chunks(N0, g.P2)
        send = [_split(t, 0, xc) for t in stage]
        recv = alltoall(g.comm1_groups(), send)
        out = []
        for r in range(g.P):
            U = F.fft(np.concatenate(recv[r], axis=1), 1)  # pencil.py:753,878
            if padded:
                V = np.zeros(g.complex_shape(r), dtype=ct)
                U = trunc_fold(U, V, N1_, 1)
                U /= padsize ** 3
            out.append(U.astype(ct))
        return out

    # alignment X
    # comm1: split z over P2, gather y (subarrays2B -> 2A, pencil.py:1324-1326)
    send = [_split(t, 2, zc) for t in stage]
    recv = alltoall(g.comm1_groups(), send)
    stage = []
    for r in range(g.P):
        U = F.fft(np.concatenate(recv[r], axis=1), 1)  # pencil.py:1327,1461
        if padded:  # copy_from_padded_y pencil.py:375-379,1464
            V = np.zeros((U.shape[0], N1_, U.shape[2]), dtype=ct)
            U = trunc_fold(U, V, N1_, 1)
        stage.append(U)
    # comm0: split y over P1, gather x (subarrays1B -> 1A, pencil.py:1331-1333)
    yc = _chunks(N1_, g.P1)
    send = [_split(t, 1, yc) for t in stage]
    recv = alltoall(g.comm0_groups(), send)
    out = []
    for r in range(g.P):
        U = F.fft(np.concatenate(recv[r], axis=0), 0)  # pencil.py:1336,1472
        if padded:
            V = np.zeros(g.complex_shape(r), dtype=ct)
            U = trunc_fold(U, V, N0, 0)
            U /= padsize ** 3
        out.append(U.astype(ct))
    return out


def ifftn(fu, N, P, alignment="X", P1=None, communication="Alltoall", dealias=None, padsize=1.5,
          precision="double"):
    """Inverse transform of all ranks (Y: ``pencil.py:386-632``; X: ``:1001-1226``)."""
    assert dealias in ("3/2-rule", "2/3-rule", "None", None)
    g = Geometry(N, P, alignment, P1, communication, padsize)
    rt, ct = dtypes(precision)
    padded = dealias == "3/2-rule"
    p = padsize if padded else 1
    N0, N1_, N2_ = (int(n) for n in g.N)
    Nf = g.Nf
    zc = g.zparts()
    fu = [np.asarray(f, dtype=ct) for f in fu]
    for r in range(g.P):
        assert fu[r].shape == g.complex_shape(r), (fu[r].shape, g.complex_shape(r))
    if dealias == "2/3-rule":
        fu = [f * g.mask(r) for r, f in enumerate(fu)]
    if padded:
        fu = [(f * padsize ** 3).astype(ct) for f in fu]  # pencil.py:560,1159

    def zfinish(U):
        """Assemble Nf planes (AlltoallN: Nyquist plane zero), pad z, C2R (pencil.py:428-431,
        503-506, 624-629)."""
        full = np.zeros(U.shape[:2] + (int(p * N2_) // 2 + 1,), dtype=ct)
        full[:, :, :U.shape[2]] = U
        return F.irfft(full, 2).astype(rt)

    if alignment == "Y":
        stage = []
        for r in range(g.P):
            U = fu[r]
            if padded:  # copy_to_padded_y, pencil.py:356-359,600
                V = np.zeros((U.shape[0], int(p * N1_), U.shape[2]), dtype=ct)
                U = pad_copy(U, V, N1_, 1)
            stage.append(F.ifft(U, 1))  # pencil.py:491,603
        yc = _chunks(int(p * N1_), g.P2)
        send = [_split(t, 1, yc) for t in stage]  # comm1: split y, gather x (1A -> 1B)
        recv = alltoall(g.comm1_groups(), send)
        stage = []
        for r in range(g.P):
            U = np.concatenate(recv[r], axis=0)
            if padded:  # copy_to_padded_x, pencil.py:351-354,611
                V = np.zeros((int(p * N0),) + U.shape[1:], dtype=ct)
                U = pad_copy(U, V, N0, 0)
            stage.append(F.ifft(U, 0))  # pencil.py:498,612
        xc = _chunks(int(p * N0), g.P1)
        send = [_split(t, 0, xc) for t in stage]  # comm0: split x, gather z (2A -> 2B)
        recv = alltoall(g.comm0_groups(), send)
        return [zfinish(np.concatenate(recv[r], axis=2)) for r in range(g.P)]

    stage = []
    for r in range(g.P):
        U = fu[r]
        if padded:  # copy_to_padded_x, pencil.py:1196
            V = np.zeros((int(p * N0),) + U.shape[1:], dtype=ct)
            U = pad_copy(U, V, N0, 0)
        stage.append(F.ifft(U, 0))  # pencil.py:1090,1199
    xc = _chunks(int(p * N0), g.P1)
    send = [_split(t, 0, xc) for t in stage]  # comm0: split x, gather y (1A -> 1B)
    recv = alltoall(g.comm0_groups(), send)
    stage = []
    for r in range(g.P):
        U = np.concatenate(recv[r], axis=1)
        if padded:  # copy_to_padded_y, pencil.py:1207
            V = np.zeros((U.shape[0], int(p * N1_), U.shape[2]), dtype=ct)
            U = pad_copy(U, V, N1_, 1)
        stage.append(F.ifft(U, 1))  # pencil.py:1097,1209
    yc = _chunks(int(p * N1_), g.P2)
    send = [_split(t, 1, yc) for t in stage]  # comm1: split y, gather z (2A -> 2B)
    recv = alltoall(g.comm1_groups(), send)
    return [zfinish(np.concatenate(recv[r], axis=2)) for r in range(g.P)]
