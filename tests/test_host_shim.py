"""The product's C-ABI layer (csrc/b200fft.cu: plan objects, the program execution loop with its
descriptor construction, event bookkeeping, timing / step records, error codes)
executed on the CPU: built with g++ against an inert CUDA runtime stand-in, kernels bound to the emulator
(tests/host_shim_util.py).  The emulator tests run the plan PROGRAMS; these run the code that executes
them on the device, through the same entry points the Python classes call."""
import ctypes as C

import numpy as np
import pytest

import host_shim_util
import oracle
from mpifft4py_b200 import _cdefs as D

TOL = {"double": 5e-14, "single": 5e-6}


def _plan(L, kind, N, prec, **kw):
    d = D.PlanDesc()
    d.kind = kind
    d.precision = D.DOUBLE if prec == "double" else D.SINGLE
    for i, n in enumerate(N):
        d.N[i] = n
    d.nranks, d.rank, d.P1, d.P2 = 1, 0, 1, 1
    d.padsize = 1.5
    for k, v in kw.items():
        setattr(d, k, v)
    h = C.c_void_p()
    rc = L.b200fft_plan_create(C.byref(h), C.byref(d))
    return rc, h


def _run(L, h, inverse, mode, src, dst):
    fn = L.b200fft_exec_inverse if inverse else L.b200fft_exec_forward
    L.shim_next_epoch()  # an event recorded by an earlier call must not satisfy a wait of this one
    rc = fn(h, C.c_void_p(src.ctypes.data), C.c_void_p(dst.ctypes.data), mode, None)
    assert rc == 0, (rc, L.b200fft_last_error())
    return dst


def _steps(L, h):
    M = 4096
    n = C.c_int()
    ty, ms, by, ln, ps = (C.c_int * M)(), (C.c_float * M)(), (C.c_double * M)(), (C.c_int * M)(), (C.c_int * M)()
    assert L.b200fft_plan_last_steps(h, M, C.byref(n), ty, ms, by, ln, ps) == 0
    return [(ty[i], ms[i], by[i], ln[i], ps[i]) for i in range(n.value)]


@pytest.mark.parametrize("layout", [D.LAYOUT_YBLOCK, D.LAYOUT_NATURAL])
@pytest.mark.parametrize("prec", ["double", "single"])
def test_slab_single_rank_through_the_c_abi(prec, layout):
    """fftn / ifftn of every dealias mode through b200fft_exec_forward / _inverse, y-blocked (default) and natural
    intermediate layout, with timing records on."""
    L = host_shim_util.load()
    N = (8, 16, 32)
    rt, ct = oracle.common.dtypes(prec)
    rc, h = _plan(L, D.SLAB, N, prec, layout=layout)
    assert rc == 0, L.b200fft_last_error()
    assert L.b200fft_plan_set_timing(h, 1) == 0
    rng = np.random.default_rng(3)
    g = oracle.slab.Geometry(N, 1)
    A = rng.random(N).astype(rt)
    c = _run(L, h, 0, D.DEALIAS_NONE, A, np.full(g.complex_shape(), np.nan, dtype=ct))
    assert oracle.rel_l2(c, oracle.slab.fftn([A], N, 1, precision=prec)[0]) <= TOL[prec]
    k, x = C.c_int(), C.c_int()
    assert L.b200fft_plan_last_launches(h, C.byref(k), C.byref(x)) == 0
    assert (k.value, x.value) == (3, 0)
    st = _steps(L, h)
    assert len(st) == 3 and all(s[1] >= 0 for s in st)
    # the records carry the algorithmic bytes of the whole transform whatever the layout
    csz, rsz = np.dtype(ct).itemsize, np.dtype(rt).itemsize
    Nf = N[2] // 2 + 1
    total = N[0] * N[1] * (N[2] * rsz + Nf * csz) + 2 * 2 * N[0] * N[1] * Nf * csz
    assert sum(s[2] for s in st) == pytest.approx(total)
    back = _run(L, h, 1, D.DEALIAS_NONE, c, np.full(N, np.nan, dtype=rt))
    assert oracle.rel_l2(back, A) <= TOL[prec]
    fu = (rng.standard_normal(g.complex_shape()) + 1j * rng.standard_normal(g.complex_shape())).astype(ct)
    for mode, name in ((D.DEALIAS_2_3, "2/3-rule"), (D.DEALIAS_3_2, "3/2-rule")):
        shp = g.real_shape_padded() if name == "3/2-rule" else g.real_shape()
        got = _run(L, h, 1, mode, fu, np.full(shp, np.nan, dtype=rt))
        assert oracle.rel_l2(got, oracle.slab.ifftn([fu], N, 1, dealias=name, precision=prec)[0]) <= TOL[prec]
    up = rng.random(g.real_shape_padded()).astype(rt)
    got = _run(L, h, 0, D.DEALIAS_3_2, up, np.full(g.complex_shape(), np.nan, dtype=ct))
    assert oracle.rel_l2(got, oracle.slab.fftn([up], N, 1, dealias="3/2-rule", precision=prec)[0]) <= TOL[prec]
    assert L.b200fft_plan_workspace_bytes(h) > 0
    f, xms = C.c_float(), C.c_float()
    assert L.b200fft_plan_last_phase_ms(h, C.byref(f), C.byref(xms)) == 0 and f.value >= 0
    assert L.b200fft_plan_destroy(h) == 0


def test_c2c_and_line_plans_and_error_codes():
    L = host_shim_util.load()
    rng = np.random.default_rng(2)
    N = (8, 16, 32)
    rc, h = _plan(L, D.SLAB_C2C, N, "double")
    assert rc == 0
    A = (rng.standard_normal(N) + 1j * rng.standard_normal(N))
    c = _run(L, h, 0, D.DEALIAS_NONE, A, np.zeros(N, dtype=np.complex128))
    assert oracle.rel_l2(c, np.fft.fftn(A)) <= 5e-14
    assert oracle.rel_l2(_run(L, h, 1, D.DEALIAS_NONE, c, np.zeros(N, dtype=np.complex128)), A) <= 5e-14
    L.b200fft_plan_destroy(h)
    rc, h = _plan(L, D.LINE, (32, 64), "double")
    assert rc == 0
    B = rng.random((32, 64))
    c = _run(L, h, 0, D.DEALIAS_NONE, B, np.zeros((32, 33), dtype=np.complex128))
    assert oracle.rel_l2(c, np.fft.rfft2(B)) <= 5e-14
    L.b200fft_plan_destroy(h)
    # error codes (mapped to the reference's exceptions by the Python layer)
    rc, h = _plan(L, D.SLAB, (10, 16, 32), "double")            # 10 has no radix plan
    assert rc == D.ERR_UNSUPPORTED and b"length" in L.b200fft_last_error()
    rc, h = _plan(L, D.SLAB, (8, 16, 32), "double", nranks=3)   # slab.py:89-91
    assert rc == D.ERR_RANKS
    rc, h = _plan(L, D.PENCIL_X, (8, 16, 32), "double")         # pencil needs more than one rank
    assert rc == D.ERR_ARG
    rc, h = _plan(L, D.SLAB, (8, 16, 32), "double", pipeline=7)
    assert rc == D.ERR_ARG
    rc, h = _plan(L, D.SLAB, (8, 16, 32), "double")
    assert rc == 0
    A = rng.random((8, 16, 32))
    assert L.b200fft_exec_forward(h, C.c_void_p(A.ctypes.data), None, 0, None) == D.ERR_ARG
    assert L.b200fft_exec_forward(h, C.c_void_p(A.ctypes.data), C.c_void_p(A.ctypes.data), 9, None) == D.ERR_ARG
    L.b200fft_plan_destroy(h)
