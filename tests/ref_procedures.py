"""The reference's own test procedures, restated for this engine (TEST INFRASTRUCTURE; no test collects from here).

/root/reference/tests/test_FFT.py checks the distributed classes the same way in every case: rank 0 draws the data and
computes the expected result with SERIAL transforms of the same backend (the module-level ``rfftn`` / ``fftn``, or a
``COMM_SELF`` slab / line object for the padded cases), ``Bcast`` hands both to every rank, each rank transforms its
block and compares it with its slice of the expected array, with tolerances relative to the largest entry
(``:76-91``).  The functions below follow that data flow, the fixtures' parameter lists (``:24-56``) and the
tolerances (``:75``, ``:146``, ``:217``), with mpifft4py_b200's classes, serial functions and communicators in the
places of mpiFFT4py's -- arrays are numpy arrays throughout, as in every reference caller.  Run at P = 1 by
tests/test_gpu_reference_procedures.py and on every rank count by tests/gpu_dist_worker.py.

Deviations, on purpose: ``abs(x).max()`` where upstream divides by ``x.max()`` of a complex array; the ``COMM_SELF``
object has the precision of the object under test (upstream's ``isinstance(FFT.float, np.float32)`` is always false,
SURVEY.md 8a-Q5, so its serial side is always double).
"""
import numpy as np

import mpifft4py_b200 as m
from mpifft4py_b200.comm import COMM_SELF

L3 = np.array([2 * np.pi] * 3)
N0 = 2 ** 5  # tests/test_FFT.py:19


def tolerances(F):
    """(atol, rtol) of tests/test_FFT.py:75."""
    return (1e-10, 1e-8) if F.float is np.float64 else (5e-7, 1e-4)


def precision_of(F):
    return "double" if F.float is np.float64 else "single"


def make(param, comm):
    """Object for one fixture parameter of tests/test_FFT.py:24-56: 'slab' + a|w + s|d, 'pencil' + s|n|a + x|y + s|d,
    'line' + s|d, 'c2c' + s|d."""
    prec = {"s": "single", "d": "double"}[param[-1]]
    N = np.array([N0, 2 * N0, 4 * N0])
    if param.startswith("pencil"):
        communication = {"s": "Alltoall", "n": "AlltoallN", "a": "Alltoallw"}[param[-3]]
        return m.Pencil_R2C(N, L3, comm, prec, communication=communication, alignment=param[-2].upper())
    if param.startswith("slab"):
        return m.Slab_R2C(N, L3, comm, prec, communication="Alltoall" if param[-2] == "a" else "Alltoallw")
    if param.startswith("line"):
        return m.Line_R2C(N[:2], L3[:2], comm, prec)
    return m.Slab_C2C(N, L3, comm, prec)


def params(P):
    """The parameter lists of the three fixtures for a communicator of P ranks."""
    slab = ["slabas", "slabad", "slabws", "slabwd"]
    pencil = ["pencil" + c + a + p for c in "sna" for a in "xy" for p in "sd"] if P >= 4 else []
    return slab + pencil, ["lines", "lined"], ["c2cd", "c2cs"]


def _close(got, want, F, what):
    atol, rtol = tolerances(F)
    assert np.allclose(got, want, rtol, atol), "%s: max |difference| %.3e" % (what, np.abs(got - want).max())


def _close_rel_max(got, want, F, what):
    _, rtol = tolerances(F)
    err = np.abs(got - want).max() / np.abs(got).max()
    assert err < rtol, "%s: %.3e of the largest entry (limit %.1e)" % (what, err, rtol)


def forward_backward(F, rng):
    """tests/test_FFT.py:60-91 (3D classes) and :93-112 (line): distributed transform == serial rfftn / rfft2 of
    the same global array, and back."""
    is2d = len(F.N) == 2
    axes = (0, 1) if is2d else (0, 1, 2)
    serial_fwd, serial_inv = (m.rfft2, m.irfft2) if is2d else (m.rfftn, m.irfftn)
    gshape = tuple(int(s) for s in F.global_complex_shape())
    A = np.zeros(tuple(int(n) for n in F.N), dtype=F.float)
    B2 = np.zeros(gshape, dtype=F.complex)
    if F.rank == 0:
        A[...] = rng.random(A.shape)
        if getattr(F, "communication", None) == "AlltoallN":  # that layout has no Nyquist plane
            C = serial_fwd(A, np.empty(gshape, dtype=F.complex), axes=axes)
            C[:, :, -1] = 0
            A = serial_inv(C, A, axes=axes)
        B2 = serial_fwd(A, B2, axes=axes)
    F.comm.Bcast(A, root=0)
    F.comm.Bcast(B2, root=0)
    fwd, inv = (F.fft2, F.ifft2) if is2d else (F.fftn, F.ifftn)
    a = np.zeros(F.real_shape(), dtype=F.float)
    a[:] = A[F.real_local_slice()]
    c = fwd(a, np.zeros(F.complex_shape(), dtype=F.complex))
    _close_rel_max(c, B2[F.complex_local_slice()], F, "forward")
    a = inv(c, a)
    _close_rel_max(a, A[F.real_local_slice()], F, "backward")


def padded(F, rng):
    """tests/test_FFT.py:159-211 (3D classes) and :114-156 (line): the 3/2-rule transforms against a COMM_SELF slab /
    line object on the whole mesh -- the padded inverse of the same spectrum, then the truncating forward transform
    gives the spectrum back."""
    is2d = len(F.N) == 2
    prec = precision_of(F)
    if is2d:
        S = m.Line_R2C(F.N, F.L, COMM_SELF, prec)
        s_fwd, s_inv, fwd, inv = S.fft2, S.ifft2, F.fft2, F.ifft2
    else:
        S = m.Slab_R2C(F.N, L3, COMM_SELF, prec, communication=F.communication)
        s_fwd, s_inv, fwd, inv = S.fftn, S.ifftn, F.fftn, F.ifftn
    C = np.zeros(tuple(int(s) for s in F.global_complex_shape()), dtype=F.complex)
    A_pad = np.zeros(S.real_shape_padded(), dtype=F.float)
    if F.rank == 0:
        A = rng.random(tuple(int(n) for n in F.N)).astype(F.float)
        C = s_fwd(A, C)
        if is2d:
            C[-int(F.N[0]) // 2] = 0          # "Eliminate Nyquist, otherwise test will fail" (:128)
        elif F.communication == "AlltoallN":
            C[:, :, -1] = 0
        A_pad = s_inv(C, A_pad, dealias="3/2-rule")
    F.comm.Bcast(C, root=0)
    F.comm.Bcast(A_pad, root=0)
    c = np.zeros(F.complex_shape(), dtype=F.complex)
    c[:] = C[F.complex_local_slice()]
    ae = np.zeros(F.real_shape_padded(), dtype=F.float)
    ae[:] = A_pad[F.real_local_slice(padsize=1.5)]
    ap = inv(c, np.zeros(F.real_shape_padded(), dtype=F.float), dealias="3/2-rule")
    _close(ap, ae, F, "padded inverse")
    cp = fwd(ap, np.zeros(F.complex_shape(), dtype=F.complex), dealias="3/2-rule")
    _close_rel_max(cp, c, F, "truncating forward")


def c2c(F, rng):
    """tests/test_FFT.py:213-273: slab.C2C padded and plain, the expected arrays from serial fftn / ifftn with the
    spectrum copied into the 3/2-sized one by hand."""
    N = [int(n) for n in F.N]
    Np = [3 * n // 2 for n in N]
    A = np.zeros(N, dtype=F.complex)
    C = np.zeros(F.global_shape(), dtype=F.complex)
    Ap = np.zeros(Np, dtype=F.complex)
    if F.rank == 0:
        A = (rng.random(N) + rng.random(N) * 1j).astype(F.complex)
        C = m.fftn(A, C, axes=(0, 1, 2))
        Cp = np.zeros(Np, dtype=F.complex)
        ks = (np.fft.fftfreq(N[2]) * N[2]).astype(int)
        h0, h1 = N[0] // 2, N[1] // 2
        Cp[:h0, :h1, ks] = C[:h0, :h1]
        Cp[:h0, -h1:, ks] = C[:h0, h1:]
        Cp[-h0:, :h1, ks] = C[h0:, :h1]
        Cp[-h0:, -h1:, ks] = C[h0:, h1:]
        Ap = m.ifftn(Cp * 1.5 ** 3, Ap, axes=(0, 1, 2))
    for arr in (C, Ap, A):
        F.comm.Bcast(arr, root=0)
    ae = np.zeros(F.original_shape_padded(), dtype=F.complex)
    ae[:] = Ap[F.original_local_slice(padsize=1.5)]
    c = np.zeros(F.transformed_shape(), dtype=F.complex)
    c[:] = C[F.transformed_local_slice()]
    atol, rtol = (1e-8, 1e-8) if F.float is np.float64 else (5e-7, 1e-4)   # :217
    ap = F.ifftn(c, np.zeros(F.original_shape_padded(), dtype=F.complex), dealias="3/2-rule")
    assert np.allclose(ap, ae, rtol, atol), "c2c padded inverse: %.3e" % np.abs(ap - ae).max()
    cp = F.fftn(ap, np.zeros(F.transformed_shape(), dtype=F.complex), dealias="3/2-rule")
    _close_rel_max(cp, c, F, "c2c truncating forward")
    aa = F.ifftn(c, np.zeros(F.original_shape(), dtype=F.complex))
    assert np.allclose(aa, A[F.original_local_slice()], rtol, atol), "c2c inverse"
    c2 = F.fftn(aa, np.zeros(F.transformed_shape(), dtype=F.complex))
    _close_rel_max(c2, c, F, "c2c forward")


def run_all(comm, note=None, only=None):
    """Every fixture parameter x every procedure that takes it, as pytest would run the reference's module on this
    communicator.  Returns the number of (parameter, procedure) cases run."""
    P = comm.Get_size()
    three_d, lines, c2cs = params(P)
    rng = np.random.default_rng(2024)  # (only rank 0's draws are used: the data travels by Bcast)
    done = 0
    for p in three_d + lines:
        if only and p not in only:
            continue
        if note:
            note("reference procedure: %s" % p)
        F = make(p, comm)
        forward_backward(F, rng)
        padded(F, rng)
        done += 2
    for p in c2cs:
        if only and p not in only:
            continue
        if note:
            note("reference procedure: %s" % p)
        c2c(make(p, comm), rng)
        done += 1
    return done
