"""The reference's golden vectors (tests/golden/*.npz: outputs of the UNMODIFIED mpiFFT4py, every rank's
slices included) through the product's C-ABI layer on the CPU -- host build of csrc/b200fft.cu, ranks as
threads, kernels in the emulator (tests/host_shim_util.py).  Same comparisons as the multi-GPU worker
(tests/gpu_dist_worker.py: run_golden), for every transport the plan kind supports."""
import glob
import json
import os

import numpy as np
import pytest

import host_shim_util
import oracle
from conftest import GOLDEN
from mpifft4py_b200 import _cdefs as D
from test_host_shim_multi import Ranks, _exec, _make_plan

FILES = sorted(glob.glob(os.path.join(GOLDEN, "*.npz")))
TOL = {"double": 1e-12, "single": 1e-5}
KIND = {("slab", None): D.SLAB, ("pencil", "X"): D.PENCIL_X, ("pencil", "Y"): D.PENCIL_Y, ("line", None): D.LINE}


@pytest.mark.parametrize("transport", [D.TRANSPORT_NCCL, D.TRANSPORT_P2P, D.TRANSPORT_STORE])
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_reference_golden_through_the_c_abi(path, transport):
    z = dict(np.load(path))  # (every array read here: NpzFile decompresses lazily and is not thread-safe)
    meta = json.loads(str(z["meta"]))
    P, prec, N = meta["P"], meta["precision"], meta["N"]
    if P == 1 and transport != D.TRANSPORT_NCCL:
        pytest.skip("single rank: no exchange")
    tol = TOL[prec]
    kind = KIND[(meta["kind"], meta["alignment"])]
    P1 = P2 = 1
    if meta["kind"] == "pencil":
        g = oracle.pencil.Geometry(N, P, meta["alignment"], meta["P1"], meta["communication"])
        P1, P2 = g.P1, g.P2
    drop = int(meta["communication"] == "AlltoallN")
    A, Cg, Ap = z["A"], z["C"], z["Ap"]
    Cin = Cg.copy()
    if meta["kind"] == "line":
        Cin[-N[0] // 2] = 0
    L = host_shim_util.load()
    R = Ranks(P)

    def rank(r):
        info = meta["ranks"][r]
        rs = tuple(slice(*s) for s in info["real_local_slice"])
        rps = tuple(slice(*s) for s in info["real_local_slice_padded"])
        cs = tuple(slice(*s) for s in info["complex_local_slice"])
        h, comms = _make_plan(L, R, r, kind, N, P, transport if P > 1 else D.TRANSPORT_NCCL, P1=P1, P2=P2, drop=drop,
                              precision=D.DOUBLE if prec == "double" else D.SINGLE)
        c = _exec(L, h, 0, D.DEALIAS_NONE, np.ascontiguousarray(A[rs]), np.full(Cg[cs].shape, np.nan, dtype=Cg.dtype))
        assert oracle.rel_l2(c, Cg[cs]) <= tol, "C"
        a2 = _exec(L, h, 1, D.DEALIAS_NONE, np.ascontiguousarray(Cg[cs]), np.full(A[rs].shape, np.nan, dtype=A.dtype))
        assert oracle.rel_l2(a2, z["A2"][rs]) <= tol, "A2"
        ap = _exec(L, h, 1, D.DEALIAS_3_2, np.ascontiguousarray(Cin[cs]), np.full(Ap[rps].shape, np.nan, dtype=Ap.dtype))
        assert oracle.rel_l2(ap, Ap[rps]) <= tol, "Ap"
        cp = _exec(L, h, 0, D.DEALIAS_3_2, np.ascontiguousarray(Ap[rps]), np.full(Cg[cs].shape, np.nan, dtype=Cg.dtype))
        assert oracle.rel_l2(cp, z["Cp"][cs]) <= 10 * tol, "Cp"
        if meta["has23"]:
            a23 = _exec(L, h, 1, D.DEALIAS_2_3, np.ascontiguousarray(Cg[cs]), np.full(A[rs].shape, np.nan, dtype=A.dtype))
            assert oracle.rel_l2(a23, z["A23"][rs]) <= tol, "A23"
        R.bar.wait()
        assert L.b200fft_plan_destroy(h) == 0
        for cm in comms:
            L.b200fft_comm_destroy(cm)

    R.run(rank)


C2C_FILES = sorted(glob.glob(os.path.join(os.path.dirname(GOLDEN), "golden_c2c", "*.npz")))


@pytest.mark.parametrize("transport", [D.TRANSPORT_NCCL, D.TRANSPORT_P2P, D.TRANSPORT_STORE])
@pytest.mark.parametrize("path", C2C_FILES, ids=[os.path.basename(f)[:-4] for f in C2C_FILES])
def test_reference_c2c_golden_through_the_c_abi(path, transport):
    """slab.C2C goldens (outputs of the unmodified reference, slab.py:538-825)."""
    z = dict(np.load(path))
    meta = json.loads(str(z["meta"]))
    P, prec, N = meta["P"], meta["precision"], meta["N"]
    if P == 1 and transport != D.TRANSPORT_NCCL:
        pytest.skip("single rank: no exchange")
    tol = TOL[prec]
    A, Cg, Ap = z["A"], z["C"], z["Ap"]
    L = host_shim_util.load()
    R = Ranks(P)

    def rank(r):
        info = meta["ranks"][r]
        rs = tuple(slice(*s) for s in info["real_local_slice"])
        rps = tuple(slice(*s) for s in info["real_local_slice_padded"])
        cs = tuple(slice(*s) for s in info["complex_local_slice"])
        h, comms = _make_plan(L, R, r, D.SLAB_C2C, N, P, transport if P > 1 else D.TRANSPORT_NCCL,
                              precision=D.DOUBLE if prec == "double" else D.SINGLE)
        c = _exec(L, h, 0, D.DEALIAS_NONE, np.ascontiguousarray(A[rs]), np.full(Cg[cs].shape, np.nan, dtype=Cg.dtype))
        assert oracle.rel_l2(c, Cg[cs]) <= tol, "C"
        a2 = _exec(L, h, 1, D.DEALIAS_NONE, np.ascontiguousarray(Cg[cs]), np.full(A[rs].shape, np.nan, dtype=A.dtype))
        assert oracle.rel_l2(a2, z["A2"][rs]) <= tol, "A2"
        ap = _exec(L, h, 1, D.DEALIAS_3_2, np.ascontiguousarray(Cg[cs]), np.full(Ap[rps].shape, np.nan, dtype=Ap.dtype))
        assert oracle.rel_l2(ap, Ap[rps]) <= tol, "Ap"
        cp = _exec(L, h, 0, D.DEALIAS_3_2, np.ascontiguousarray(Ap[rps]), np.full(Cg[cs].shape, np.nan, dtype=Cg.dtype))
        assert oracle.rel_l2(cp, z["Cp"][cs]) <= 10 * tol, "Cp"
        R.bar.wait()
        assert L.b200fft_plan_destroy(h) == 0
        for cm in comms:
            L.b200fft_comm_destroy(cm)

    R.run(rank)
