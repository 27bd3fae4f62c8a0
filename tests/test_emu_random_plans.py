"""Seeded random sweep over the slab plan space -- mesh x ranks x transport x pipeline x chunk count x
L2 group size x kind x dealias mode -- in the CPU emulator against the oracle.  The fixed cases of test_emu_plans.py pin
the common configurations; this one looks for index-map mistakes in the corners (uneven kz ranges, chunk
counts that do not divide the local planes, non-cubic meshes, 3*2^k sizes)."""
import ctypes as C

import numpy as np
import pytest

import emu_util
import oracle
from mpifft4py_b200 import _cdefs as D
from test_emu_plans import TOL, _check, _desc, _rand_c, run_plan

SIZES = [8, 12, 16, 24, 32, 48]


def _cases(count, seed):
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < count:
        N = tuple(int(rng.choice(SIZES)) for _ in range(3))
        P = int(rng.choice([2, 4, 8]))
        if N[0] % P or N[1] % P or P > N[0] // 2:
            continue
        transport = int(rng.choice([D.TRANSPORT_NCCL, D.TRANSPORT_P2P, D.TRANSPORT_STORE]))
        pipeline = int(rng.choice([D.PIPELINE_X, D.PIPELINE_KZ]))
        chunks = int(rng.choice([0, 1, 2, 3, 4, 5]))
        kind = str(rng.choice(["r2c", "c2c"]))
        prec = "double" if rng.random() < 0.75 else "single"
        out.append((N, P, transport, pipeline, chunks, kind, prec))
    return out


def _supported(n):
    while n % 2 == 0:
        n //= 2
    return n in (1, 3)


@pytest.mark.parametrize("N,P,transport,pipeline,chunks,kind,prec", _cases(60, 2026),
                         ids=lambda v: "x".join(map(str, v)) if isinstance(v, tuple) else str(v))
def test_random_slab_plan(N, P, transport, pipeline, chunks, kind, prec):
    rt, ct = oracle.common.dtypes(prec)
    c2c = kind == "c2c"
    g = oracle.slab.Geometry(N, P)
    rng = np.random.default_rng(sum(N) * P + chunks)
    d = _desc(D.SLAB_C2C if c2c else D.SLAB, N, P, prec, chunks=chunks, pipeline=pipeline, transport=transport)
    tol = TOL[prec]
    if c2c:
        cs, it = (N[0], N[1] // P, N[2]), ct
        new_in = lambda shape: _rand_c(rng, shape, ct)
        fwd = lambda u, **k: oracle.slab.c2c_fftn(u, N, P, precision=prec, **k)
        inv = lambda fu, **k: oracle.slab.c2c_ifftn(fu, N, P, precision=prec, **k)
    else:
        cs, it = g.complex_shape(), rt
        new_in = lambda shape: rng.random(shape).astype(rt)
        fwd = lambda u, **k: oracle.slab.fftn(u, N, P, precision=prec, **k)
        inv = lambda fu, **k: oracle.slab.ifftn(fu, N, P, precision=prec, **k)
    u = [new_in(g.real_shape()) for _ in range(P)]
    _check(run_plan(d, 0, D.DEALIAS_NONE, u, [cs] * P, ct), fwd(u), tol)
    fu = [_rand_c(rng, cs, ct) for _ in range(P)]
    _check(run_plan(d, 1, D.DEALIAS_NONE, fu, [g.real_shape()] * P, it), inv(fu), tol)
    _check(run_plan(d, 1, D.DEALIAS_2_3, fu, [g.real_shape()] * P, it), inv(fu, dealias="2/3-rule"), tol)
    modes = [D.DEALIAS_NONE, D.DEALIAS_2_3]
    if all(_supported(3 * n // 2) and n % 2 == 0 for n in N):  # padded lengths need a radix plan too
        up = [new_in(g.real_shape_padded()) for _ in range(P)]
        _check(run_plan(d, 0, D.DEALIAS_3_2, up, [cs] * P, ct), fwd(up, dealias="3/2-rule"), tol)
        _check(run_plan(d, 1, D.DEALIAS_3_2, fu, [g.real_shape_padded()] * P, it), inv(fu, dealias="3/2-rule"), tol)
        modes.append(D.DEALIAS_3_2)
    if transport != D.TRANSPORT_NCCL:
        lib = emu_util.load()
        for inverse in (0, 1):
            for m in modes:
                n = C.c_int()
                assert lib.emu_check_p2p(C.byref(d), inverse, m, C.byref(n)) == 0


def _pencil_cases(count, seed):
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < count:
        N = tuple(int(rng.choice([8, 16, 32, 48])) for _ in range(3))
        P1, P2 = [(2, 2), (4, 2), (2, 4), (4, 4), (2, 8)][int(rng.integers(0, 5))]
        alignment = str(rng.choice(["X", "Y"]))
        zparts = P2 if alignment == "X" else P1
        if any(n % P1 or n % P2 for n in N) or (N[2] // 2) % zparts:
            continue
        comm = str(rng.choice(["Alltoall", "Alltoallw", "AlltoallN"]))
        transport = int(rng.choice([D.TRANSPORT_NCCL, D.TRANSPORT_P2P, D.TRANSPORT_STORE]))
        prec = "double" if rng.random() < 0.75 else "single"
        chunks = int(rng.choice([0, 0, 2, 3, 4]))
        out.append((N, P1, P2, alignment, comm, transport, prec, chunks))
    return out


@pytest.mark.parametrize("N,P1,P2,alignment,comm,transport,prec,chunks", _pencil_cases(40, 77),
                         ids=lambda v: "x".join(map(str, v)) if isinstance(v, tuple) else str(v))
def test_random_pencil_plan(N, P1, P2, alignment, comm, transport, prec, chunks):
    P = P1 * P2
    rt, ct = oracle.common.dtypes(prec)
    g = oracle.pencil.Geometry(N, P, alignment, P1, comm)
    assert (g.P1, g.P2) == (P1, P2)
    rng = np.random.default_rng(sum(N) + P)
    d = _desc(D.PENCIL_X if alignment == "X" else D.PENCIL_Y, N, P, prec, P1, P2, int(comm == "AlltoallN"), transport=transport,
              chunks=chunks)
    kw = dict(alignment=alignment, P1=P1, communication=comm, precision=prec)
    tol = TOL[prec]
    A = rng.random(N).astype(rt)
    u = [A[g.real_local_slice(r)] for r in range(P)]
    cshape = [g.complex_shape(r) for r in range(P)]
    _check(run_plan(d, 0, D.DEALIAS_NONE, u, cshape, ct), oracle.pencil.fftn(u, N, P, **kw), tol)
    fu = [_rand_c(rng, s, ct) for s in cshape]
    modes = [(D.DEALIAS_NONE, None), (D.DEALIAS_2_3, "2/3-rule")]
    # the padded blocks must tile the padded mesh: 1.5 * N / P1 and 1.5 * N / P2 integral
    if all(_supported(3 * n // 2) for n in N) and all((3 * n) % (2 * q) == 0 for n in N[:2] for q in (P1, P2)):
        modes.append((D.DEALIAS_3_2, "3/2-rule"))
    else:
        lib0 = emu_util.load()
        if all(_supported(3 * n // 2) for n in N):  # supported lengths, untileable blocks: refused, not overrun
            ins = [np.zeros(s_, dtype=ct) for s_ in cshape]
            outs = [np.zeros(g.real_shape_padded(), dtype=rt) for _ in range(P)]
            ip = (C.c_void_p * P)(*[a.ctypes.data for a in ins])
            op = (C.c_void_p * P)(*[a.ctypes.data for a in outs])
            assert lib0.emu_plan_run(C.byref(d), 1, D.DEALIAS_3_2, ip, op) == D.ERR_ARG
    for mode, name in modes:
        shp = g.real_shape_padded() if name == "3/2-rule" else g.real_shape()
        _check(run_plan(d, 1, mode, fu, [shp] * P, rt), oracle.pencil.ifftn(fu, N, P, dealias=name, **kw), tol)
    if len(modes) == 3:
        up = [rng.random(g.real_shape_padded()).astype(rt) for _ in range(P)]
        _check(run_plan(d, 0, D.DEALIAS_3_2, up, cshape, ct), oracle.pencil.fftn(up, N, P, dealias="3/2-rule", **kw), tol)
    lib = emu_util.load()
    for inverse in (0, 1):
        for mode, _ in modes:
            assert lib.emu_check_schedule(C.byref(d), inverse, mode) == 0
            if transport != D.TRANSPORT_NCCL:
                n = C.c_int()
                assert lib.emu_check_p2p(C.byref(d), inverse, mode, C.byref(n)) == 0
