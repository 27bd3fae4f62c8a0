"""slab.C2C (reference ``mpiFFT4py/slab.py:538-825``; its own test: ``tests/test_FFT.py:213-273``).

 * the oracle restatement against golden outputs of the UNMODIFIED reference (tests/golden_c2c);
 * the plan programs + kernels in the CPU emulator against the oracle, every rank count;
 * the host class's shapes / slices against the goldens;
 * (gpu) the class on the device against the oracle and the goldens."""
import glob
import json
import os

import numpy as np
import pytest

import oracle
from conftest import ROOT
from mpifft4py_b200 import _cdefs as D
from test_emu_plans import _check, _desc, _rand_c, run_plan

GOLDEN = os.path.join(ROOT, "tests", "golden_c2c")
FILES = sorted(glob.glob(os.path.join(GOLDEN, "*.npz")))
assert FILES, "C2C golden fixtures missing"
TOL = {"double": 5e-14, "single": 5e-6}


def _split(G, slices):
    return [np.ascontiguousarray(G[tuple(slice(*s) for s in sl)]) for sl in slices]


def _assemble(shape, dtype, parts, slices):
    G = np.zeros(shape, dtype=dtype)
    for p, sl in zip(parts, slices):
        G[tuple(slice(*s) for s in sl)] = p
    return G


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_oracle_c2c_matches_reference_golden(path):
    z = np.load(path)
    meta = json.loads(str(z["meta"]))
    N, P, prec = meta["N"], meta["P"], meta["precision"]
    ranks = meta["ranks"]
    rs = [r["real_local_slice"] for r in ranks]
    rps = [r["real_local_slice_padded"] for r in ranks]
    cs = [r["complex_local_slice"] for r in ranks]
    tol = 2e-14 if prec == "double" else 2e-6
    g = oracle.slab.GeometryC2C(N, P)
    for r, info in enumerate(ranks):
        assert list(g.original_shape()) == info["real_shape"]
        assert list(g.transformed_shape()) == info["complex_shape"]
        assert list(g.original_shape_padded()) == info["real_shape_padded"]
        assert [[s.start, s.stop, s.step] for s in g.transformed_local_slice(r)] == info["complex_local_slice"]
    A, Cg = z["A"], z["C"]
    c = oracle.slab.c2c_fftn(_split(A, rs), N, P, precision=prec)
    assert oracle.rel_l2(_assemble(Cg.shape, Cg.dtype, c, cs), Cg) <= tol
    assert oracle.rel_l2(Cg, np.fft.fftn(A.astype(np.complex128))) <= tol  # test_FFT.py:222-223
    a2 = oracle.slab.c2c_ifftn(_split(Cg, cs), N, P, precision=prec)
    assert oracle.rel_l2(_assemble(A.shape, A.dtype, a2, rs), z["A2"]) <= tol
    ap = oracle.slab.c2c_ifftn(_split(Cg, cs), N, P, dealias="3/2-rule", precision=prec)
    Ap = z["Ap"]
    assert oracle.rel_l2(_assemble(Ap.shape, Ap.dtype, ap, rps), Ap) <= tol
    cp = oracle.slab.c2c_fftn(_split(Ap, rps), N, P, dealias="3/2-rule", precision=prec)
    assert oracle.rel_l2(_assemble(Cg.shape, Cg.dtype, cp, cs), z["Cp"]) <= 10 * tol


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_host_class_matches_reference_golden(path):
    """Bit-exact shapes and slices of mpifft4py_b200.slab.C2C (no GPU touched)."""
    import mpifft4py_b200 as m
    z = np.load(path)
    meta = json.loads(str(z["meta"]))
    N, P, prec = np.array(meta["N"]), meta["P"], meta["precision"]

    class FakeComm(object):
        def __init__(self, r):
            self.r = r

        def Get_size(self):
            return P

        def Get_rank(self):
            return self.r

    for r, info in enumerate(meta["ranks"]):
        F = m.Slab_C2C(N, np.array([2 * np.pi] * 3), FakeComm(r), prec)
        assert [int(x) for x in F.original_shape()] == info["real_shape"]
        assert [int(x) for x in F.transformed_shape()] == info["complex_shape"]
        assert [int(x) for x in F.original_shape_padded()] == info["real_shape_padded"]
        assert [[int(s.start), int(s.stop), s.step] for s in F.original_local_slice()] == info["real_local_slice"]
        assert [[int(s.start), int(s.stop), s.step] for s in F.original_local_slice(padsize=1.5)] == info["real_local_slice_padded"]
        assert [[int(s.start), int(s.stop), s.step] for s in F.transformed_local_slice()] == info["complex_local_slice"]
        assert [int(x) for x in F.global_shape()] == info["global_shape"]
        assert [int(x) for x in F.global_shape(1.5)] == info["global_shape_padded"]


@pytest.mark.parametrize("chunks", [0, 2])
@pytest.mark.parametrize("prec", ["double", "single"])
@pytest.mark.parametrize("P", [1, 2, 4])
@pytest.mark.parametrize("N", [(8, 16, 32), (16, 16, 16), (32, 8, 8)])
def test_c2c_plans_in_emulator(N, P, prec, chunks):
    if P > N[0] // 2 or N[1] % P:
        pytest.skip("illegal decomposition")
    if chunks and (P == 1 or prec == "single"):
        pytest.skip("chunking only changes multi-rank programs")
    rt, ct = oracle.common.dtypes(prec)
    g = oracle.slab.GeometryC2C(N, P)
    rng = np.random.default_rng(sum(N) + P)
    d = _desc(D.SLAB_C2C, N, P, prec, chunks=chunks)
    tol = TOL[prec]
    A = _rand_c(rng, N, ct)
    u = [np.ascontiguousarray(A[g.real_local_slice(r)]) for r in range(P)]
    _check(run_plan(d, 0, D.DEALIAS_NONE, u, [g.complex_shape()] * P, ct), oracle.slab.c2c_fftn(u, N, P, precision=prec), tol)
    fu = [_rand_c(rng, g.complex_shape(), ct) for _ in range(P)]
    for mode, name in ((D.DEALIAS_NONE, None), (D.DEALIAS_2_3, "2/3-rule"), (D.DEALIAS_3_2, "3/2-rule")):
        shp = g.real_shape_padded() if name == "3/2-rule" else g.real_shape()
        _check(run_plan(d, 1, mode, fu, [shp] * P, ct), oracle.slab.c2c_ifftn(fu, N, P, dealias=name, precision=prec), tol)
    up = [_rand_c(rng, g.real_shape_padded(), ct) for _ in range(P)]
    _check(run_plan(d, 0, D.DEALIAS_3_2, up, [g.complex_shape()] * P, ct),
           oracle.slab.c2c_fftn(up, N, P, dealias="3/2-rule", precision=prec), tol)


def test_c2c_23_rule_is_the_serial_masked_ifftn():
    """The intended 2/3-rule (the reference's own raises, see oracle.slab.GeometryC2C.mask)."""
    N, P = (16, 16, 16), 2
    g = oracle.slab.GeometryC2C(N, P)
    rng = np.random.default_rng(0)
    Cg = _rand_c(rng, N, np.complex128)
    k = [np.fft.fftfreq(n, 1. / n) for n in N]
    kmax = 2. / 3. * (np.array(N) // 2 + 1)
    keep = ((abs(k[0]) < kmax[0])[:, None, None] * (abs(k[1]) < kmax[1])[None, :, None] * (abs(k[2]) < kmax[2])[None, None, :])
    ref = np.fft.ifftn(Cg * keep)
    fu = [np.ascontiguousarray(Cg[g.complex_local_slice(r)]) for r in range(P)]
    got = oracle.slab.c2c_ifftn(fu, N, P, dealias="2/3-rule")
    for r in range(P):
        assert oracle.rel_l2(got[r], ref[g.real_local_slice(r)]) < 1e-14


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["double", "single"])
@pytest.mark.parametrize("N", [(32, 64, 128), (32, 32, 32)])
def test_c2c_on_gpu_against_oracle(N, prec):
    import torch
    import mpifft4py_b200 as m
    from mpifft4py_b200.comm import SelfComm
    torch.cuda.set_device(0)
    rt, ct = oracle.common.dtypes(prec)
    tol = 1e-12 if prec == "double" else 1e-5
    F = m.Slab_C2C(np.array(N), np.array([2 * np.pi] * 3), SelfComm(), prec)
    rng = np.random.default_rng(7)
    A = _rand_c(rng, N, ct)
    c = F.fftn(A, np.zeros(F.transformed_shape(), dtype=ct))
    assert oracle.rel_l2(c, oracle.slab.c2c_fftn([A], N, 1, precision=prec)[0]) <= tol
    assert oracle.rel_l2(F.ifftn(c, np.zeros(F.original_shape(), dtype=ct)), A) <= tol
    fu = _rand_c(rng, F.transformed_shape(), ct)
    for name in (None, "2/3-rule", "3/2-rule"):
        shp = F.original_shape_padded() if name == "3/2-rule" else F.original_shape()
        got = F.ifftn(fu, np.zeros(shp, dtype=ct), dealias=name)
        assert oracle.rel_l2(got, oracle.slab.c2c_ifftn([fu], N, 1, dealias=name, precision=prec)[0]) <= tol
    up = _rand_c(rng, F.original_shape_padded(), ct)
    got = F.fftn(up, np.zeros(F.transformed_shape(), dtype=ct), dealias="3/2-rule")
    assert oracle.rel_l2(got, oracle.slab.c2c_fftn([up], N, 1, dealias="3/2-rule", precision=prec)[0]) <= tol
    # CUDA tensors in place of numpy arrays
    tdt = torch.complex128 if prec == "double" else torch.complex64
    tf = torch.zeros(tuple(int(s) for s in F.transformed_shape()), dtype=tdt, device="cuda")
    F.fftn(torch.from_numpy(A).cuda(), tf)
    assert oracle.rel_l2(tf.cpu().numpy(), c) <= 1e-15


@pytest.mark.gpu
def test_c2c_on_gpu_against_reference_golden():
    import torch
    import mpifft4py_b200 as m
    from mpifft4py_b200.comm import SelfComm
    torch.cuda.set_device(0)
    z = np.load(os.path.join(GOLDEN, "c2c_P1_d.npz"))
    meta = json.loads(str(z["meta"]))
    F = m.Slab_C2C(np.array(meta["N"]), np.array([2 * np.pi] * 3), SelfComm(), "double")
    A, Cg, Ap = z["A"], z["C"], z["Ap"]
    assert oracle.rel_l2(F.fftn(A, np.zeros_like(Cg)), Cg) <= 1e-12
    assert oracle.rel_l2(F.ifftn(Cg, np.zeros_like(A)), z["A2"]) <= 1e-12
    assert oracle.rel_l2(F.ifftn(Cg, np.zeros_like(Ap), dealias="3/2-rule"), Ap) <= 1e-12
    assert oracle.rel_l2(F.fftn(Ap, np.zeros_like(Cg), dealias="3/2-rule"), z["Cp"]) <= 1e-11
