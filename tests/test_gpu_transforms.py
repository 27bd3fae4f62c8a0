"""GPU parity tests proper: the drop-in classes (through the C ABI, real kernels) against the
oracle, the reference's golden vectors and size-independent properties.

Tolerances are the north-star's: relative L2 <= 1e-12 (double), <= 1e-5 (single)."""
import json
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

TOL = {"double": 1e-12, "single": 1e-5}
L3 = np.array([2 * np.pi] * 3)


def _mod():
    import mpifft4py_b200 as m
    return m


def _self():
    from mpifft4py_b200.comm import SelfComm
    return SelfComm()


def _rand_c(rng, shape, ct):
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(ct)


@pytest.mark.parametrize("prec", ["double", "single"])
@pytest.mark.parametrize("N", [(32, 64, 128), (8, 16, 32), (64, 64, 64)])
def test_slab_single_rank_vs_oracle(N, prec):
    m = _mod()
    rt, ct = oracle.common.dtypes(prec)
    tol = TOL[prec]
    F = m.Slab_R2C(np.array(N), L3, _self(), prec)
    rng = np.random.default_rng(1234)
    A = rng.random(N).astype(rt)
    c = F.fftn(A, np.zeros(F.complex_shape(), dtype=ct))
    assert c.dtype == ct
    assert oracle.rel_l2(c, oracle.slab.fftn([A], N, 1, precision=prec)[0]) <= tol
    assert oracle.rel_l2(c, np.fft.rfftn(A.astype(np.float64), axes=(0, 1, 2))) <= tol
    a = F.ifftn(c, np.zeros(F.real_shape(), dtype=rt))
    assert oracle.rel_l2(a, A) <= tol  # round trip
    fu = _rand_c(rng, F.complex_shape(), ct)
    keep = fu.copy()
    for d in (None, "2/3-rule", "3/2-rule"):
        shp = F.real_shape_padded() if d == "3/2-rule" else F.real_shape()
        u = F.ifftn(fu, np.zeros(shp, dtype=rt), dealias=d)
        assert oracle.rel_l2(u, oracle.slab.ifftn([fu], N, 1, dealias=d, precision=prec)[0]) <= tol
        assert np.array_equal(fu, keep), "ifftn must not modify its input"
    up = rng.random(F.real_shape_padded()).astype(rt)
    cp = F.fftn(up, np.zeros(F.complex_shape(), dtype=ct), dealias="3/2-rule")
    assert oracle.rel_l2(cp, oracle.slab.fftn([up], N, 1, dealias="3/2-rule", precision=prec)[0]) <= tol
    # reference's padded test (tests/test_FFT.py:159-205): pad, then truncate, returns the spectrum
    ap = F.ifftn(c, np.zeros(F.real_shape_padded(), dtype=rt), dealias="3/2-rule")
    cp = F.fftn(ap, np.zeros(F.complex_shape(), dtype=ct), dealias="3/2-rule")
    assert np.all(np.abs((cp - c) / cp.max()) < (1e-8 if prec == "double" else 1e-4))


def test_slab_cuda_tensors_zero_copy():
    import torch
    m = _mod()
    N = (32, 64, 128)
    F = m.Slab_R2C(np.array(N), L3, _self(), "double")
    A = np.random.default_rng(3).random(N)
    u = torch.from_numpy(A).cuda()
    fu = torch.zeros(tuple(int(s) for s in F.complex_shape()), dtype=torch.complex128, device="cuda")
    out = F.fftn(u, fu)
    assert out is fu
    assert oracle.rel_l2(fu.cpu().numpy(), np.fft.rfftn(A, axes=(0, 1, 2))) <= 1e-12
    assert np.array_equal(u.cpu().numpy(), A), "fftn must not modify its input"
    u2 = torch.zeros_like(u)
    F.ifftn(fu, u2)
    assert oracle.rel_l2(u2.cpu().numpy(), A) <= 1e-12


@pytest.mark.parametrize("prec", ["double", "single"])
def test_line_single_rank_vs_oracle(prec):
    m = _mod()
    rt, ct = oracle.common.dtypes(prec)
    tol = TOL[prec]
    N = (64, 128)
    F = m.Line_R2C(np.array(N), L3[:2], _self(), prec)
    rng = np.random.default_rng(5)
    A = rng.random(N).astype(rt)
    c = F.fft2(A, np.zeros(F.complex_shape(), dtype=ct))
    assert oracle.rel_l2(c, np.fft.rfft2(A.astype(np.float64))) <= tol
    assert oracle.rel_l2(F.ifft2(c, np.zeros(F.real_shape(), dtype=rt)), A) <= tol
    fu = _rand_c(rng, F.complex_shape(), ct)
    for d in (None, "2/3-rule", "3/2-rule"):
        shp = F.real_shape_padded() if d == "3/2-rule" else F.real_shape()
        u = F.ifft2(fu, np.zeros(shp, dtype=rt), dealias=d)
        assert oracle.rel_l2(u, oracle.line.ifft2([fu], N, 1, dealias=d, precision=prec)[0]) <= tol
    up = rng.random(F.real_shape_padded()).astype(rt)
    cp = F.fft2(up, np.zeros(F.complex_shape(), dtype=ct), dealias="3/2-rule")
    assert oracle.rel_l2(cp, oracle.line.fft2([up], N, 1, dealias="3/2-rule", precision=prec)[0]) <= tol


@pytest.mark.parametrize("name", ["slab_P1_Alltoallw_d", "line_P1_d"])
def test_single_rank_against_reference_golden(name):
    """Outputs of the UNMODIFIED reference (tests/golden) reproduced on the GPU."""
    m = _mod()
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    N = meta["N"]
    if meta["kind"] == "slab":
        F = m.Slab_R2C(np.array(N), L3, _self(), "double")
        fwd, inv = F.fftn, F.ifftn
    else:
        F = m.Line_R2C(np.array(N), L3[:2], _self(), "double")
        fwd, inv = F.fft2, F.ifft2
    A, C = z["A"], z["C"]
    assert oracle.rel_l2(fwd(A, np.zeros_like(C)), C) <= 1e-12
    assert oracle.rel_l2(inv(C, np.zeros_like(A)), z["A2"]) <= 1e-12
    Cin = C.copy()
    if meta["kind"] == "line":
        Cin[-N[0] // 2] = 0
    Ap = z["Ap"]
    assert oracle.rel_l2(inv(Cin, np.zeros_like(Ap), dealias="3/2-rule"), Ap) <= 1e-12
    assert oracle.rel_l2(fwd(Ap, np.zeros_like(C), dealias="3/2-rule"), z["Cp"]) <= 1e-12
    assert oracle.rel_l2(inv(C, np.zeros_like(A), dealias="2/3-rule"), z["A23"]) <= 1e-12


def test_serial_functions_vs_numpy():
    m = _mod()
    rng = np.random.default_rng(11)
    a = rng.standard_normal((16, 32, 64))
    c = _rand_c(rng, (16, 32, 33), np.complex128)
    assert oracle.rel_l2(m.rfftn(a, axes=(0, 1, 2)), np.fft.rfftn(a, axes=(0, 1, 2))) <= 1e-12
    assert oracle.rel_l2(m.irfftn(c, axes=(0, 1, 2)), np.fft.irfftn(c, s=(16, 32, 64), axes=(0, 1, 2))) <= 1e-12
    assert oracle.rel_l2(m.rfft2(a, axes=(1, 2)), np.fft.rfft2(a, axes=(1, 2))) <= 1e-12
    assert oracle.rel_l2(m.irfft2(c, axes=(1, 2)), np.fft.irfft2(c, s=(32, 64), axes=(1, 2))) <= 1e-12
    assert oracle.rel_l2(m.rfft(a, axis=2), np.fft.rfft(a, axis=2)) <= 1e-12
    assert oracle.rel_l2(m.irfft(c, axis=2), np.fft.irfft(c, n=64, axis=2)) <= 1e-12
    c2 = _rand_c(rng, (16, 48, 64), np.complex128)  # complex lengths must be 2^k or 3*2^k
    for ax in (0, 1, 2):
        assert oracle.rel_l2(m.fft(c2, axis=ax), np.fft.fft(c2, axis=ax)) <= 1e-12
        assert oracle.rel_l2(m.ifft(c2, axis=ax), np.fft.ifft(c2, axis=ax)) <= 1e-12
    for ax in (0, 1):  # odd inner extent J = 33 next to the transformed axis
        assert oracle.rel_l2(m.fft(c, axis=ax), np.fft.fft(c, axis=ax)) <= 1e-12
    b = np.zeros_like(c)
    assert m.fft(c, b, axis=0) is b and oracle.rel_l2(b, np.fft.fft(c, axis=0)) <= 1e-12
    a32 = a.astype(np.float32)
    out = m.rfftn(a32, axes=(0, 1, 2))
    assert out.dtype == np.complex64 and oracle.rel_l2(out, np.fft.rfftn(a, axes=(0, 1, 2))) <= 1e-5
    with pytest.raises(NotImplementedError):
        m.fft(_rand_c(rng, (10, 4), np.complex128), axis=0)  # length 10 has no radix plan


@pytest.mark.parametrize("prec,N", [("single", (256, 256, 256)), ("double", (512, 512, 512)),
                                    ("double", (1024, 1024, 1024))])
def test_full_size_properties(prec, N):
    """BASELINE sizes on one GPU: round trip, and an exact check through separability -- for
    u = a(x) b(y) c(z) the 3D transform is the outer product of three 1D transforms."""
    import torch
    m = _mod()
    rt, ct = oracle.common.dtypes(prec)
    tol = TOL[prec]
    free, _ = torch.cuda.mem_get_info()
    need = 5 * np.prod(N) * np.dtype(rt).itemsize
    if free < need:
        pytest.skip("not enough device memory")
    F = m.Slab_R2C(np.array(N), L3, _self(), prec)
    tdt = torch.float64 if prec == "double" else torch.float32
    g = torch.Generator(device="cuda").manual_seed(7)
    u = torch.rand(tuple(N), dtype=tdt, device="cuda", generator=g)
    fu = torch.empty(tuple(int(s) for s in F.complex_shape()), dtype=torch.complex128 if prec == "double" else torch.complex64,
                     device="cuda")
    F.fftn(u, fu)
    u2 = torch.empty_like(u)
    F.ifftn(fu, u2)
    err = (torch.linalg.vector_norm(u2 - u) / torch.linalg.vector_norm(u)).item()
    assert err <= tol, err
    # Parseval (Hermitian half-spectrum weights)
    w = torch.full((fu.shape[2],), 2.0, dtype=tdt, device="cuda")
    w[0] = 1.0
    w[-1] = 1.0
    lhs = (u.double() ** 2).sum().item()
    rhs = ((fu.abs().double() ** 2) * w.double()).sum().item() / float(np.prod(N))
    assert abs(lhs - rhs) / lhs <= 10 * tol
    del u2
    rng = np.random.default_rng(1)
    a, b, c = (rng.random(n).astype(rt) for n in N)
    u.copy_(torch.from_numpy(a).cuda()[:, None, None] * torch.from_numpy(b).cuda()[None, :, None] *
            torch.from_numpy(c).cuda()[None, None, :])
    F.fftn(u, fu)
    fa, fb, fc = np.fft.fft(a.astype(np.float64)), np.fft.fft(b.astype(np.float64)), np.fft.rfft(c.astype(np.float64))
    ref = (torch.from_numpy(fa).cuda()[:, None, None] * torch.from_numpy(fb).cuda()[None, :, None] *
           torch.from_numpy(fc).cuda()[None, None, :])
    err = (torch.linalg.vector_norm(fu.to(torch.complex128) - ref) / torch.linalg.vector_norm(ref)).item()
    assert err <= tol, err


def test_3_2_rule_full_size_roundtrip():
    """1024^3-class padded sizes exercise the 3*2^k plans (768-point here: 512^3 padded)."""
    import torch
    m = _mod()
    N = (512, 512, 512)
    F = m.Slab_R2C(np.array(N), L3, _self(), "double")
    g = torch.Generator(device="cuda").manual_seed(9)
    u = torch.rand(tuple(N), dtype=torch.float64, device="cuda", generator=g)
    fu = torch.empty(tuple(int(s) for s in F.complex_shape()), dtype=torch.complex128, device="cuda")
    F.fftn(u, fu)
    up = torch.empty(tuple(int(s) for s in F.real_shape_padded()), dtype=torch.float64, device="cuda")
    F.ifftn(fu, up, dealias="3/2-rule")
    fu2 = torch.empty_like(fu)
    F.fftn(up, fu2, dealias="3/2-rule")
    err = (torch.linalg.vector_norm(fu2 - fu) / torch.linalg.vector_norm(fu)).item()
    assert err <= 1e-12, err
    # padded field sampled on the coarse grid points that coincide (every 3rd of every 2nd)
    sub = up[::3, ::3, ::3]
    ref = u[::2, ::2, ::2]
    # interpolation is exact only without the Nyquist modes; compare through the spectrum instead
    assert sub.shape == ref.shape


@pytest.mark.parametrize("dealias", [None, "2/3-rule", "3/2-rule"])
@pytest.mark.parametrize("prec", ["double", "single"])
@pytest.mark.parametrize("N", [(32, 64, 128), (64, 16, 32)])
def test_slab_single_rank_natural_layout_vs_oracle(N, prec, dealias):
    """layout="natural": z, y, x on [x][y][kz] like slab.py:366-370 (the y-blocked default is what every other test
    runs); both directions of every dealias mode against the oracle, and bit-for-bit the same shapes."""
    m = _mod()
    rt, ct = oracle.common.dtypes(prec)
    tol = TOL[prec]
    F = m.Slab_R2C(np.array(N), L3, _self(), prec)
    F.layout = "natural"
    rng = np.random.default_rng(77)
    fu = _rand_c(rng, F.complex_shape(), ct)
    shp = F.real_shape_padded() if dealias == "3/2-rule" else F.real_shape()
    u = F.ifftn(fu, np.zeros(shp, dtype=rt), dealias=dealias)
    assert oracle.rel_l2(u, oracle.slab.ifftn([fu], N, 1, dealias=dealias, precision=prec)[0]) <= tol
    fdeal = dealias if dealias == "3/2-rule" else None
    a = rng.random(shp).astype(rt)
    c = F.fftn(a, np.zeros(F.complex_shape(), dtype=ct), dealias=fdeal)
    assert oracle.rel_l2(c, oracle.slab.fftn([a], N, 1, dealias=fdeal, precision=prec)[0]) <= tol
    assert F.last_launches()[0] == 3


@pytest.mark.parametrize("layout", ["yblock", "natural"])
@pytest.mark.parametrize("N,prec", [((16384, 64), "single"), ((8192, 32), "double"), ((4096, 32), "double")])
def test_line_long_columns_two_launches(N, prec, layout):
    """line.R2C with columns of >= 64 KB (BASELINE config 5a shape class): the x pass runs as two launches (four-step)
    by default, as one with layout="natural"; fft2 / ifft2 against numpy either way."""
    m = _mod()
    rt, ct = oracle.common.dtypes(prec)
    tol = TOL[prec]
    F = m.Line_R2C(np.array(N), L3[:2], _self(), prec)
    F.layout = layout
    rng = np.random.default_rng(N[0])
    A = rng.random(N).astype(rt)
    c = F.fft2(A, np.zeros(F.complex_shape(), dtype=ct))
    assert oracle.rel_l2(c, np.fft.rfft2(A.astype(np.float64))) <= tol
    split = layout == "yblock" and N[0] * np.dtype(ct).itemsize >= 64 * 1024
    assert F.last_launches()[0] == (3 if split else 2)
    a = F.ifft2(c, np.zeros(F.real_shape(), dtype=rt))
    assert oracle.rel_l2(a, A) <= tol
    fu = _rand_c(rng, F.complex_shape(), ct)
    u = F.ifft2(fu, np.zeros(F.real_shape(), dtype=rt), dealias="2/3-rule")
    assert oracle.rel_l2(u, oracle.line.ifft2([fu], N, 1, dealias="2/3-rule", precision=prec)[0]) <= tol
