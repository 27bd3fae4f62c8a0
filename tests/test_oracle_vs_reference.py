"""Live check of the oracle against the UNMODIFIED reference at the reference's own test size
N = (32, 64, 128) (tests/test_FFT.py:22-46) on generic random inputs.  Runs only where
/root/reference exists (the build container); the GPU box relies on tests/golden instead."""
import os
import sys
import warnings

import numpy as np
import pytest

import oracle

SHIM = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "refshim")
pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/mpiFFT4py"),
                                reason="reference tree not present")
N = (32, 64, 128)


def _ref():
    if SHIM not in sys.path:
        sys.path.insert(0, SHIM)
    import load_reference
    warnings.simplefilter("ignore")
    load_reference.load()
    return load_reference


def _run(kind, P, comm, alignment=None, P1=None, prec="double"):
    lr = _ref()
    from mpi4py import MPI
    rng = np.random.default_rng(7)
    rt = np.float64 if prec == "double" else np.float32
    A = rng.random(N).astype(rt)
    Apad = rng.random(tuple(int(1.5 * n) for n in N)).astype(rt)
    Nn = np.array(N, dtype=int)
    L = np.array([2 * np.pi] * 3)

    def body():
        if kind == "slab":
            from mpiFFT4py.slab import R2C
            F = R2C(Nn, L, MPI.COMM_WORLD, prec, communication=comm)
        else:
            from mpiFFT4py.pencil import R2C
            F = R2C(Nn, L, MPI.COMM_WORLD, prec, P1=P1, communication=comm, alignment=alignment)
        a = A[F.real_local_slice()].copy()
        c = F.fftn(a, np.zeros(F.complex_shape(), dtype=F.complex)).copy()
        ap = Apad[F.real_local_slice(padsize=1.5)].copy()
        cp = F.fftn(ap, np.zeros(F.complex_shape(), dtype=F.complex), dealias="3/2-rule").copy()
        up = F.ifftn(cp.copy(), np.zeros(F.real_shape_padded(), dtype=F.float), dealias="3/2-rule").copy()
        return (F.real_local_slice(), F.real_local_slice(padsize=1.5)), c, cp, up

    res = lr.run_ranks(P, body)
    sl = [r[0] for r in res]
    if kind == "slab":
        fwd = lambda u, d=None: oracle.slab.fftn(u, N, P, dealias=d, precision=prec)
        inv = lambda f, d=None: oracle.slab.ifftn(f, N, P, dealias=d, precision=prec)
    else:
        kw = dict(alignment=alignment, P1=P1, communication=comm, precision=prec)
        fwd = lambda u, d=None: oracle.pencil.fftn(u, N, P, dealias=d, **kw)
        inv = lambda f, d=None: oracle.pencil.ifftn(f, N, P, dealias=d, **kw)
    tol = 1e-13 if prec == "double" else 1e-5
    c = fwd([A[s[0]] for s in sl])
    cp = fwd([Apad[s[1]] for s in sl], "3/2-rule")
    up = inv([r[2] for r in res], "3/2-rule")
    for r in range(P):
        assert c[r].shape == res[r][1].shape and c[r].dtype == res[r][1].dtype
        assert oracle.rel_l2(c[r], res[r][1]) <= tol
        assert oracle.rel_l2(cp[r], res[r][2]) <= tol
        assert oracle.rel_l2(up[r], res[r][3]) <= tol


@pytest.mark.parametrize("P,comm", [(1, "Alltoallw"), (4, "Alltoallw"), (8, "Alltoall")])
def test_slab(P, comm):
    _run("slab", P, comm)


@pytest.mark.parametrize("alignment", ["X", "Y"])
@pytest.mark.parametrize("P,P1", [(4, None), (8, None), (8, 2)])
def test_pencil_alltoallw(alignment, P, P1):
    _run("pencil", P, "Alltoallw", alignment, P1)


def test_slab_single():
    _run("slab", 4, "Alltoallw", prec="single")
