"""Generate ``tests/golden/*.npz`` by running the UNMODIFIED reference under the refshim.

TEST INFRASTRUCTURE.  Run in the build container only (needs ``/root/reference``):

    python oracle/refshim/make_golden.py

Each file holds, for one (class, N, P, communication, alignment, precision) configuration:
the seeded global input ``A``; the assembled global results of the reference's ``fftn`` (``C``),
``ifftn(C)`` (``A2``), ``ifftn(C, '3/2-rule')`` (``Ap``), ``fftn(Ap, '3/2-rule')`` (``Cp``),
``ifftn(C, '2/3-rule')`` (``A23``, where the reference's implementation works); and a JSON
``meta`` with every rank's shapes and slices (steps included: 1 vs None matters).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(HERE)), "tests", "golden")
SEED = 1234


def _sl(s):
    return [[int(x.start), int(x.stop), (None if x.step is None else int(x.step))] for x in s]


def _shape(t):
    return [int(x) for x in t]


def slab_or_pencil(kind, N, P, precision, communication, alignment=None, P1=None):
    from mpi4py import MPI
    N = np.array(N, dtype=int)
    L = np.array([2 * np.pi] * 3)
    rt = np.float32 if precision == "single" else np.float64
    A = np.random.default_rng(SEED).random(tuple(N)).astype(rt)
    if communication == "AlltoallN":  # tests/test_FFT.py:64-68
        C0 = np.fft.rfftn(A.astype(np.float64), axes=(0, 1, 2))
        C0[:, :, -1] = 0
        A = np.fft.irfftn(C0, s=tuple(N), axes=(0, 1, 2)).astype(rt)
    do23 = not (kind == "pencil" and alignment == "X")  # reference R2CX 2/3-rule is broken (Q1)

    def body():
        if kind == "slab":
            from mpiFFT4py.slab import R2C
            FFT = R2C(N, L, MPI.COMM_WORLD, precision, communication=communication)
        else:
            from mpiFFT4py.pencil import R2C
            FFT = R2C(N, L, MPI.COMM_WORLD, precision, P1=P1, communication=communication,
                      alignment=alignment)
        info = dict(rank=int(FFT.rank), real_shape=_shape(FFT.real_shape()),
                    complex_shape=_shape(FFT.complex_shape()),
                    real_shape_padded=_shape(FFT.real_shape_padded()),
                    real_local_slice=_sl(FFT.real_local_slice()),
                    real_local_slice_padded=_sl(FFT.real_local_slice(padsize=1.5)),
                    complex_local_slice=_sl(FFT.complex_local_slice()),
                    work_shape_32=_shape(FFT.work_shape("3/2-rule")),
                    work_shape_none=_shape(FFT.work_shape(None)))
        if kind == "pencil":
            info.update(P1=int(FFT.P1), P2=int(FFT.P2), comm0_rank=int(FFT.comm0_rank),
                        comm1_rank=int(FFT.comm1_rank))
        a = np.zeros(FFT.real_shape(), dtype=FFT.float)
        a[:] = A[FFT.real_local_slice()]
        c = np.zeros(FFT.complex_shape(), dtype=FFT.complex)
        c = FFT.fftn(a, c).copy()
        a2 = np.zeros(FFT.real_shape(), dtype=FFT.float)
        a2 = FFT.ifftn(c.copy(), a2).copy()
        ap = np.zeros(FFT.real_shape_padded(), dtype=FFT.float)
        ap = FFT.ifftn(c.copy(), ap, dealias="3/2-rule").copy()
        cp = np.zeros(FFT.complex_shape(), dtype=FFT.complex)
        cp = FFT.fftn(ap.copy(), cp, dealias="3/2-rule").copy()
        a23 = None
        if do23:
            a23 = np.zeros(FFT.real_shape(), dtype=FFT.float)
            a23 = FFT.ifftn(c.copy(), a23, dealias="2/3-rule").copy()
        return info, c, a2, ap, cp, a23

    res = load_reference.run_ranks(P, body)
    ct = np.complex64 if precision == "single" else np.complex128
    Nf = int(N[2]) // 2 + 1
    C = np.zeros((int(N[0]), int(N[1]), Nf), dtype=ct)
    Cp = np.zeros_like(C)
    A2 = np.zeros(tuple(N), dtype=rt)
    A23 = np.zeros(tuple(N), dtype=rt)
    Ap = np.zeros(tuple(int(1.5 * n) for n in N), dtype=rt)
    metas = []
    for info, c, a2, ap, cp, a23 in res:
        cs = tuple(slice(*s) for s in info["complex_local_slice"])
        rs = tuple(slice(*s) for s in info["real_local_slice"])
        rps = tuple(slice(*s) for s in info["real_local_slice_padded"])
        C[cs] = c
        Cp[cs] = cp
        A2[rs] = a2
        Ap[rps] = ap
        if a23 is not None:
            A23[rs] = a23
        metas.append(info)
    meta = dict(kind=kind, N=_shape(N), P=P, precision=precision, communication=communication,
                alignment=alignment, P1=P1, seed=SEED, has23=bool(do23), ranks=metas,
                reference_version="1.1.2", maths=load_reference.load()._refshim_maths)
    out = dict(A=A, C=C, A2=A2, Ap=Ap, Cp=Cp, meta=np.array(json.dumps(meta)))
    if do23:
        out["A23"] = A23
    return out


def line(N, P, precision):
    from mpi4py import MPI
    N = np.array(N, dtype=int)
    L = np.array([2 * np.pi] * 2)
    rt = np.float32 if precision == "single" else np.float64
    A = np.random.default_rng(SEED).random(tuple(N)).astype(rt)

    def body():
        from mpiFFT4py.line import R2C
        FFT = R2C(N, L, MPI.COMM_WORLD, precision)
        info = dict(rank=int(FFT.rank), real_shape=_shape(FFT.real_shape()),
                    complex_shape=_shape(FFT.complex_shape()),
                    real_shape_padded=_shape(FFT.real_shape_padded()),
                    real_local_slice=_sl((FFT.real_local_slice()[0],)) + [[0, int(N[1]), None]],
                    real_local_slice_padded=_sl((FFT.real_local_slice(padsize=1.5)[0],)) +
                    [[0, int(1.5 * N[1]), None]],
                    complex_local_slice=[[0, int(N[0]), None]] + _sl((FFT.complex_local_slice()[1],)))
        a = np.zeros(FFT.real_shape(), dtype=FFT.float)
        a[:] = A[FFT.real_local_slice()]
        c = np.zeros(FFT.complex_shape(), dtype=FFT.complex)
        c = FFT.fft2(a, c).copy()
        a2 = np.zeros(FFT.real_shape(), dtype=FFT.float)
        a2 = FFT.ifft2(c.copy(), a2).copy()
        # tests/test_FFT.py:124-125: drop the x-Nyquist row before the padded transforms
        c0 = c.copy()
        c0[-int(N[0]) // 2] = 0
        ap = np.zeros(FFT.real_shape_padded(), dtype=FFT.float)
        ap = FFT.ifft2(c0.copy(), ap, dealias="3/2-rule").copy()
        cp = np.zeros(FFT.complex_shape(), dtype=FFT.complex)
        cp = FFT.fft2(ap.copy(), cp, dealias="3/2-rule").copy()
        a23 = np.zeros(FFT.real_shape(), dtype=FFT.float)
        a23 = FFT.ifft2(c.copy(), a23, dealias="2/3-rule").copy()
        return info, c, a2, ap, cp, a23

    res = load_reference.run_ranks(P, body)
    ct = np.complex64 if precision == "single" else np.complex128
    Nf = int(N[1]) // 2 + 1
    C = np.zeros((int(N[0]), Nf), dtype=ct)
    Cp = np.zeros_like(C)
    A2 = np.zeros(tuple(N), dtype=rt)
    A23 = np.zeros(tuple(N), dtype=rt)
    Ap = np.zeros(tuple(int(1.5 * n) for n in N), dtype=rt)
    metas = []
    for info, c, a2, ap, cp, a23 in res:
        cs = tuple(slice(*s) for s in info["complex_local_slice"])
        rs = tuple(slice(*s) for s in info["real_local_slice"])
        rps = tuple(slice(*s) for s in info["real_local_slice_padded"])
        C[cs] = c
        Cp[cs] = cp
        A2[rs] = a2
        Ap[rps] = ap
        A23[rs] = a23
        metas.append(info)
    meta = dict(kind="line", N=_shape(N), P=P, precision=precision, communication=None,
                alignment=None, P1=None, seed=SEED, has23=bool(P == 1), ranks=metas,
                reference_version="1.1.2", maths=load_reference.load()._refshim_maths)
    # P > 1: the reference's line 2/3-rule inverse returns zeros (its masked copy and the
    # zero-filled work array Uc_hat share one work_arrays key, line.py:270,287 -- same defect
    # as R2CX), so A23 is only pinned at P == 1.
    out = dict(A=A, C=C, A2=A2, Ap=Ap, Cp=Cp, meta=np.array(json.dumps(meta)))
    if P == 1:
        out["A23"] = A23
    return out


def slab_c2c(N, P, precision):
    """slab.C2C (slab.py:538-825): tests/golden_c2c/*.npz.  A: seeded complex input; C = fftn(A);
    A2 = ifftn(C); Ap = ifftn(C, '3/2-rule'); Cp = fftn(Ap, '3/2-rule').  (Its '2/3-rule' raises.)"""
    from mpi4py import MPI
    N = np.array(N, dtype=int)
    L = np.array([2 * np.pi] * 3)
    ct = np.complex64 if precision == "single" else np.complex128
    rng = np.random.default_rng(SEED)
    A = (rng.random(tuple(N)) + 1j * rng.random(tuple(N))).astype(ct)

    def body():
        from mpiFFT4py.slab import C2C
        FFT = C2C(N, L, MPI.COMM_WORLD, precision)
        info = dict(rank=int(FFT.rank), real_shape=_shape(FFT.original_shape()),
                    complex_shape=_shape(FFT.transformed_shape()),
                    real_shape_padded=_shape(FFT.original_shape_padded()),
                    real_local_slice=_sl(FFT.original_local_slice()),
                    real_local_slice_padded=_sl(FFT.original_local_slice(padsize=1.5)),
                    complex_local_slice=_sl(FFT.transformed_local_slice()),
                    global_shape=_shape(FFT.global_shape()), global_shape_padded=_shape(FFT.global_shape(1.5)))
        a = np.zeros(FFT.original_shape(), dtype=FFT.complex)
        a[:] = A[FFT.original_local_slice()]
        c = FFT.fftn(a, np.zeros(FFT.transformed_shape(), dtype=FFT.complex)).copy()
        a2 = FFT.ifftn(c.copy(), np.zeros(FFT.original_shape(), dtype=FFT.complex)).copy()
        ap = FFT.ifftn(c.copy(), np.zeros(FFT.original_shape_padded(), dtype=FFT.complex), dealias="3/2-rule").copy()
        cp = FFT.fftn(ap.copy(), np.zeros(FFT.transformed_shape(), dtype=FFT.complex), dealias="3/2-rule").copy()
        return info, c, a2, ap, cp

    res = load_reference.run_ranks(P, body)
    C = np.zeros(tuple(N), dtype=ct)
    Cp = np.zeros_like(C)
    A2 = np.zeros_like(C)
    Ap = np.zeros(tuple(int(1.5 * n) for n in N), dtype=ct)
    metas = []
    for info, c, a2, ap, cp in res:
        cs = tuple(slice(*s) for s in info["complex_local_slice"])
        rs = tuple(slice(*s) for s in info["real_local_slice"])
        rps = tuple(slice(*s) for s in info["real_local_slice_padded"])
        C[cs] = c
        Cp[cs] = cp
        A2[rs] = a2
        Ap[rps] = ap
        metas.append(info)
    meta = dict(kind="slabc2c", N=_shape(N), P=P, precision=precision, seed=SEED, ranks=metas,
                reference_version="1.1.2", maths=load_reference.load()._refshim_maths)
    return dict(A=A, C=C, A2=A2, Ap=Ap, Cp=Cp, meta=np.array(json.dumps(meta)))


CONFIGS = []
N3 = (8, 16, 32)
for P, comm, prec in [(1, "Alltoallw", "double"), (2, "Alltoall", "double"), (4, "Alltoallw", "double"),
                      (4, "Alltoall", "single")]:
    CONFIGS.append(("slab_P%d_%s_%s" % (P, comm, prec[0]), ("slab", N3, P, prec, comm, None, None)))
for al in "XY":
    for comm in ("Alltoall", "Alltoallw", "AlltoallN"):
        CONFIGS.append(("pencil%s_P4_%s_d" % (al, comm), ("pencil", N3, 4, "double", comm, al, None)))
CONFIGS.append(("pencilX_P8p1_2_Alltoallw_d", ("pencil", N3, 8, "double", "Alltoallw", "X", 2)))
CONFIGS.append(("pencilY_P8_Alltoall_d", ("pencil", N3, 8, "double", "Alltoall", "Y", None)))
CONFIGS.append(("pencilX_P4_Alltoall_s", ("pencil", N3, 4, "single", "Alltoall", "X", None)))
# eight ranks: the other grid of each alignment, and a slab whose first axis allows the 3/2-rule there (P <= N[0] // 2)
N8 = (16, 16, 32)
CONFIGS.append(("slab_P8_Alltoallw_d", ("slab", N8, 8, "double", "Alltoallw", None, None)))
CONFIGS.append(("pencilX_P8_Alltoall_s", ("pencil", N3, 8, "single", "Alltoall", "X", None)))
CONFIGS.append(("pencilY_P8p1_2_AlltoallN_d", ("pencil", N3, 8, "double", "AlltoallN", "Y", 2)))
LINES = [("line_P%d_%s" % (P, prec[0]), ((16, 32), P, prec))
         for P, prec in [(1, "double"), (2, "double"), (4, "double"), (2, "single"), (8, "double")]]


def main():
    """``--missing``: write only the files that do not exist yet (npz archives carry timestamps, so rewriting the
    others would change their bytes without changing their content)."""
    missing_only = "--missing" in sys.argv[1:]
    os.makedirs(OUT, exist_ok=True)
    load_reference.load()

    def write(path, make):
        if missing_only and os.path.exists(path):
            return
        np.savez_compressed(path, **make())
        print("wrote", os.path.basename(path))

    for name, args in CONFIGS:
        write(os.path.join(OUT, name + ".npz"), lambda: slab_or_pencil(*args))
    for name, args in LINES:
        write(os.path.join(OUT, name + ".npz"), lambda: line(*args))
    out_c2c = os.path.join(os.path.dirname(OUT), "golden_c2c")
    os.makedirs(out_c2c, exist_ok=True)
    for P, prec, N in [(1, "double", N3), (2, "double", N3), (4, "double", N3), (2, "single", N3), (8, "double", N8)]:
        write(os.path.join(out_c2c, "c2c_P%d_%s.npz" % (P, prec[0])), lambda: slab_c2c(N, P, prec))


if __name__ == "__main__":
    main()
