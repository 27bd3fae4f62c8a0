"""The distributed plan programs of libb200fft.so (csrc/plan_program.h), executed for ALL ranks in
lockstep by the CPU emulator (kernels emulated, exchanges by memcpy), against the oracle.

This pins every step list -- index maps, peer offsets, uneven Nyquist chunks, masks, scales --
without a GPU; tests/test_gpu_*.py then run the same programs on the device over NCCL."""
import ctypes as C

import numpy as np
import pytest

import emu_util
import oracle
from mpifft4py_b200 import _cdefs as D

TOL = {"double": 5e-14, "single": 5e-6}


def _desc(kind, N, P, prec, P1=1, P2=1, drop=0, chunks=0, pipeline=0, transport=0, layout=0):
    d = D.PlanDesc()
    d.kind = kind
    d.precision = D.DOUBLE if prec == "double" else D.SINGLE
    for i, n in enumerate(N):
        d.N[i] = n
    d.nranks = P
    d.rank = 0
    d.P1, d.P2 = P1, P2
    d.padsize = 1.5
    d.drop_nyquist = drop
    d.transport = transport
    d.chunks = chunks
    d.pipeline = pipeline
    d.layout = layout
    return d


def run_plan(d, inverse, dealias, ins, out_shapes, out_dtype):
    lib = emu_util.load()
    P = d.nranks
    ins = [np.ascontiguousarray(a) for a in ins]
    outs = [np.full(s, np.nan, dtype=out_dtype) for s in out_shapes]
    ip = (C.c_void_p * P)(*[a.ctypes.data for a in ins])
    op = (C.c_void_p * P)(*[a.ctypes.data for a in outs])
    rc = lib.emu_plan_run(C.byref(d), inverse, dealias, ip, op)
    assert rc == 0, rc
    return outs


def _check(got, ref, tol):
    for r, (g, e) in enumerate(zip(got, ref)):
        assert g.shape == e.shape, (r, g.shape, e.shape)
        assert np.isfinite(g).all(), "rank %d: unwritten output" % r
        assert oracle.rel_l2(g, e) <= tol, (r, oracle.rel_l2(g, e))


def _rand_c(rng, shape, ct):
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(ct)


@pytest.mark.parametrize("chunks", [0, 2, 4])
@pytest.mark.parametrize("prec", ["double", "single"])
@pytest.mark.parametrize("P", [1, 2, 4])
@pytest.mark.parametrize("N", [(8, 16, 32), (16, 16, 16), (32, 8, 8)])
def test_slab(N, P, prec, chunks):
    """chunks > 1: the exchange is cut into pieces pipelined against the FFT passes (two streams on
    the device; the emulator runs the same step list in program order)."""
    if P > N[0] // 2 or N[1] % P:
        pytest.skip("illegal decomposition")
    if chunks and (P == 1 or prec == "single"):
        pytest.skip("chunking only changes multi-rank programs; one precision is enough")
    rt, ct = oracle.common.dtypes(prec)
    g = oracle.slab.Geometry(N, P)
    rng = np.random.default_rng(sum(N) + P)
    d = _desc(D.SLAB, N, P, prec, chunks=chunks)
    tol = TOL[prec]
    A = rng.random(N).astype(rt)
    u = [A[g.real_local_slice(r)] for r in range(P)]
    ref = oracle.slab.fftn(u, N, P, precision=prec)
    got = run_plan(d, 0, D.DEALIAS_NONE, u, [g.complex_shape()] * P, ct)
    _check(got, ref, tol)
    # inverse of an arbitrary complex spectrum (non-Hermitian content included), plain and 2/3
    fu = [_rand_c(rng, g.complex_shape(), ct) for _ in range(P)]
    for mode, name in ((D.DEALIAS_NONE, None), (D.DEALIAS_2_3, "2/3-rule")):
        ref = oracle.slab.ifftn(fu, N, P, dealias=name, precision=prec)
        got = run_plan(d, 1, mode, fu, [g.real_shape()] * P, rt)
        _check(got, ref, tol)
    # 3/2-rule both ways on generic inputs
    ref = oracle.slab.ifftn(fu, N, P, dealias="3/2-rule", precision=prec)
    got = run_plan(d, 1, D.DEALIAS_3_2, fu, [g.real_shape_padded()] * P, rt)
    _check(got, ref, tol)
    up = [rng.random(g.real_shape_padded()).astype(rt) for _ in range(P)]
    ref = oracle.slab.fftn(up, N, P, dealias="3/2-rule", precision=prec)
    got = run_plan(d, 0, D.DEALIAS_3_2, up, [g.complex_shape()] * P, ct)
    _check(got, ref, tol)


@pytest.mark.parametrize("comm", ["Alltoallw", "AlltoallN"])
@pytest.mark.parametrize("P,P1", [(4, None), (8, None), (8, 2), (16, 4)])
@pytest.mark.parametrize("alignment", ["X", "Y"])
@pytest.mark.parametrize("N", [(8, 16, 32), (16, 16, 16)])
def test_pencil(N, alignment, P, P1, comm):
    _pencil_body(N, alignment, P, P1, comm, D.TRANSPORT_NCCL)


def _check_peer_mapped(d, modes):
    lib = emu_util.load()
    for inverse in (0, 1):
        for dealias in modes:
            n = C.c_int()
            rc = lib.emu_check_p2p(C.byref(d), inverse, dealias, C.byref(n))
            assert rc == 0, (rc, inverse, dealias)


@pytest.mark.parametrize("transport", [D.TRANSPORT_P2P, D.TRANSPORT_STORE])
@pytest.mark.parametrize("comm", ["Alltoallw", "AlltoallN"])
@pytest.mark.parametrize("P,P1", [(4, None), (8, None), (8, 2)])
@pytest.mark.parametrize("alignment", ["X", "Y"])
def test_pencil_peer_mapped_transports(alignment, P, P1, comm, transport):
    """Copy-engine and fused-store transports over the pencil grid's sub-communicators: buffers and flags
    are addressed by world rank; with fused stores the z pass (kz chunks) and the strided passes write
    the peers' receive buffers directly."""
    _pencil_body((8, 16, 32), alignment, P, P1, comm, transport)


def _pencil_body(N, alignment, P, P1, comm, transport):
    prec = "double"
    rt, ct = oracle.common.dtypes(prec)
    g = oracle.pencil.Geometry(N, P, alignment, P1, comm)
    zparts = g.P2 if alignment == "X" else g.P1
    if (N[2] // 2) % zparts or any(n % g.P1 or n % g.P2 for n in N):
        pytest.skip("illegal decomposition")
    rng = np.random.default_rng(sum(N) + P + (P1 or 0))
    d = _desc(D.PENCIL_X if alignment == "X" else D.PENCIL_Y, N, P, prec, g.P1, g.P2, int(comm == "AlltoallN"),
              transport=transport)
    if transport != D.TRANSPORT_NCCL:
        _check_peer_mapped(d, (D.DEALIAS_NONE, D.DEALIAS_3_2, D.DEALIAS_2_3))
    kw = dict(alignment=alignment, P1=P1, communication=comm, precision=prec)
    tol = TOL[prec]
    A = rng.random(N).astype(rt)
    u = [A[g.real_local_slice(r)] for r in range(P)]
    cshape = [g.complex_shape(r) for r in range(P)]
    _check(run_plan(d, 0, D.DEALIAS_NONE, u, cshape, ct), oracle.pencil.fftn(u, N, P, **kw), tol)
    fu = [_rand_c(rng, s, ct) for s in cshape]
    for mode, name in ((D.DEALIAS_NONE, None), (D.DEALIAS_2_3, "2/3-rule"), (D.DEALIAS_3_2, "3/2-rule")):
        shp = g.real_shape_padded() if name == "3/2-rule" else g.real_shape()
        _check(run_plan(d, 1, mode, fu, [shp] * P, rt), oracle.pencil.ifftn(fu, N, P, dealias=name, **kw), tol)
    up = [rng.random(g.real_shape_padded()).astype(rt) for _ in range(P)]
    _check(run_plan(d, 0, D.DEALIAS_3_2, up, cshape, ct), oracle.pencil.fftn(up, N, P, dealias="3/2-rule", **kw), tol)


def test_pencil_single_precision():
    N, P, prec = (8, 16, 32), 4, "single"
    rt, ct = oracle.common.dtypes(prec)
    g = oracle.pencil.Geometry(N, P, "X", None, "Alltoall")
    rng = np.random.default_rng(0)
    d = _desc(D.PENCIL_X, N, P, prec, g.P1, g.P2)
    A = rng.random(N).astype(rt)
    u = [A[g.real_local_slice(r)] for r in range(P)]
    cshape = [g.complex_shape(r) for r in range(P)]
    ref = oracle.pencil.fftn(u, N, P, alignment="X", precision=prec)
    got = run_plan(d, 0, D.DEALIAS_NONE, u, cshape, ct)
    _check(got, ref, TOL[prec])
    _check(run_plan(d, 1, D.DEALIAS_NONE, got, [g.real_shape()] * P, rt), u, TOL[prec])


@pytest.mark.parametrize("prec", ["double", "single"])
@pytest.mark.parametrize("P", [1, 2, 4, 8])
@pytest.mark.parametrize("N", [(16, 32), (64, 32)])
def test_line(N, P, prec):
    _line_body(N, P, prec, D.TRANSPORT_NCCL)


@pytest.mark.parametrize("transport", [D.TRANSPORT_P2P, D.TRANSPORT_STORE])
@pytest.mark.parametrize("P", [2, 4, 8])
def test_line_peer_mapped_transports(P, transport):
    _line_body((64, 32), P, "double", transport)


@pytest.mark.parametrize("N,prec,P,transport,chunks", [
    ((16384, 16), "single", 1, D.TRANSPORT_NCCL, 0), ((16384, 16), "single", 2, D.TRANSPORT_P2P, 0),
    ((16384, 32), "single", 4, D.TRANSPORT_NCCL, 2), ((16384, 32), "single", 4, D.TRANSPORT_STORE, 0),
    ((8192, 16), "double", 2, D.TRANSPORT_P2P, 2), ((8192, 8), "double", 1, D.TRANSPORT_NCCL, 0)])
def test_line_long_axis_four_step(N, prec, P, transport, chunks):
    """BASELINE config 5a shape class: columns of 16384 single / 8192 double points run as two launches (four-step:
    n1-point transforms over interleaved sub-columns with cross twiddles, then n2-point transforms with a scattering
    store -- into the peers' send blocks for the inverse).  fft2 / ifft2 against the oracle; the 2/3-rule inverse
    (masked first pass) keeps the single launch.  Also: the program really has the extra step, peer invariants hold."""
    rt, ct = oracle.common.dtypes(prec)
    g = oracle.line.Geometry(N, P)
    rng = np.random.default_rng(N[0] + P)
    d = _desc(D.LINE, N, P, prec, transport=transport if P > 1 else 0, chunks=chunks)
    lib = emu_util.load()
    nsteps = lib.emu_plan_steps(C.byref(d), 0, D.DEALIAS_NONE, None, 0)
    d1 = _desc(D.LINE, (N[0] // 4, N[1]), P, prec, transport=transport if P > 1 else 0, chunks=chunks)
    assert nsteps == lib.emu_plan_steps(C.byref(d1), 0, D.DEALIAS_NONE, None, 0) + 1
    assert lib.emu_plan_steps(C.byref(d), 1, D.DEALIAS_2_3, None, 0) == lib.emu_plan_steps(C.byref(d1), 1, D.DEALIAS_2_3, None, 0)
    if transport != D.TRANSPORT_NCCL and P > 1:
        _check_peer_mapped(d, (D.DEALIAS_NONE, D.DEALIAS_2_3))
    for inverse in (0, 1):
        assert lib.emu_check_schedule(C.byref(d), inverse, D.DEALIAS_NONE) == 0
    tol = 2 * TOL[prec]
    A = rng.random(N).astype(rt)
    u = [A[g.real_local_slice(r)] for r in range(P)]
    cshape = [g.complex_shape(r) for r in range(P)]
    _check(run_plan(d, 0, D.DEALIAS_NONE, u, cshape, ct), oracle.line.fft2(u, N, P, precision=prec), tol)
    fu = [_rand_c(rng, s_, ct) for s_ in cshape]
    for mode, name in ((D.DEALIAS_NONE, None), (D.DEALIAS_2_3, "2/3-rule")):
        _check(run_plan(d, 1, mode, fu, [g.real_shape()] * P, rt), oracle.line.ifft2(fu, N, P, dealias=name, precision=prec), tol)


def _line_body(N, P, prec, transport):
    if N[1] % (2 * P) or N[0] % P:
        pytest.skip("illegal decomposition")
    rt, ct = oracle.common.dtypes(prec)
    g = oracle.line.Geometry(N, P)
    rng = np.random.default_rng(sum(N) + P)
    d = _desc(D.LINE, N, P, prec, transport=transport if P > 1 else 0)
    if transport != D.TRANSPORT_NCCL and P > 1:
        _check_peer_mapped(d, (D.DEALIAS_NONE, D.DEALIAS_3_2, D.DEALIAS_2_3))
    tol = TOL[prec]
    A = rng.random(N).astype(rt)
    u = [A[g.real_local_slice(r)] for r in range(P)]
    cshape = [g.complex_shape(r) for r in range(P)]
    got = run_plan(d, 0, D.DEALIAS_NONE, u, cshape, ct)
    _check(got, oracle.line.fft2(u, N, P, precision=prec), tol)  # reference pack trick: exact here
    fu = [_rand_c(rng, s, ct) for s in cshape]
    modes = [(D.DEALIAS_NONE, None), (D.DEALIAS_3_2, "3/2-rule"), (D.DEALIAS_2_3, "2/3-rule")]
    for mode, name in modes:
        shp = g.real_shape_padded() if name == "3/2-rule" else g.real_shape()
        _check(run_plan(d, 1, mode, fu, [shp] * P, rt), oracle.line.ifft2(fu, N, P, dealias=name, precision=prec), tol)
    # 3/2 forward on a reference-test-style input (tests/test_FFT.py:112-156): both semantics agree
    C0 = np.fft.rfft2(A.astype(np.float64)).astype(ct)
    C0[-N[0] // 2] = 0
    c0 = [np.ascontiguousarray(C0[g.complex_local_slice(r)]) for r in range(P)]
    ap = oracle.line.ifft2(c0, N, P, dealias="3/2-rule", precision=prec)
    ref = oracle.line.fft2(ap, N, P, dealias="3/2-rule", precision=prec)
    _check(run_plan(d, 0, D.DEALIAS_3_2, ap, cshape, ct), ref, 4 * tol)
    # generic padded input: the engine implements plain truncation of the Nyquist column (exact=True)
    up = [rng.random(g.real_shape_padded()).astype(rt) for _ in range(P)]
    ref = oracle.line.fft2(up, N, P, dealias="3/2-rule", precision=prec, exact=(P > 1))
    _check(run_plan(d, 0, D.DEALIAS_3_2, up, cshape, ct), ref, tol)


@pytest.mark.parametrize("chunks", [0, 1, 2, 4])
@pytest.mark.parametrize("P", [2, 4, 8])
@pytest.mark.parametrize("N", [(32, 32, 32), (64, 32, 16), (1024, 1024, 1024)])
def test_slab_p2p_program_invariants(N, P, chunks):
    """Copy-engine transport: every rank's push target is exactly the peer's receive slot, peers only
    write plan-owned buffers, one credit wait / one credit return per transform, equal step counts."""
    lib = emu_util.load()
    d = _desc(D.SLAB, N, P, "double", chunks=chunks)
    d.transport = D.TRANSPORT_P2P
    for inverse in (0, 1):
        for dealias in (D.DEALIAS_NONE, D.DEALIAS_3_2, D.DEALIAS_2_3):
            if dealias == D.DEALIAS_3_2 and P > N[0] // 2:
                continue
            n = C.c_int()
            rc = lib.emu_check_p2p(C.byref(d), inverse, dealias, C.byref(n))
            assert rc == 0, (rc, inverse, dealias)
            if chunks:
                assert n.value == min(chunks, max(1, n.value))
    # the automatic pipeline follows the measurements (plan_program.h): 1024^3 double, plain transform -> kz ranges, 4
    # chunks at P = 2, 4 and 2 at P = 8; 3/2-rule -> x planes, 8 chunks at P = 2, 4 and 4 at P = 8; an explicit "x" -> 8 / 8 / 2
    if N[0] == 1024 and chunks == 0:
        n = C.c_int()
        assert lib.emu_check_p2p(C.byref(d), 0, D.DEALIAS_NONE, C.byref(n)) == 0
        assert n.value == {2: 4, 4: 4, 8: 2}[P]
        assert lib.emu_check_p2p(C.byref(d), 0, D.DEALIAS_3_2, C.byref(n)) == 0
        assert n.value == {2: 8, 4: 8, 8: 4}[P]
        d.pipeline = D.PIPELINE_X
        assert lib.emu_check_p2p(C.byref(d), 0, D.DEALIAS_NONE, C.byref(n)) == 0
        assert n.value == {2: 8, 4: 8, 8: 2}[P]


@pytest.mark.parametrize("chunks", [2, 4])
def test_slab_p2p_layout_runs_in_emulator(chunks):
    """The P2P programs receive into a plan buffer instead of the caller's output: same results."""
    N, P, prec = (16, 16, 16), 4, "double"
    rt, ct = oracle.common.dtypes(prec)
    g = oracle.slab.Geometry(N, P)
    rng = np.random.default_rng(3)
    d = _desc(D.SLAB, N, P, prec, chunks=chunks)
    d.transport = D.TRANSPORT_P2P
    A = rng.random(N).astype(rt)
    u = [A[g.real_local_slice(r)] for r in range(P)]
    ref = oracle.slab.fftn(u, N, P, precision=prec)
    got = run_plan(d, 0, D.DEALIAS_NONE, u, [g.complex_shape()] * P, ct)
    _check(got, ref, TOL[prec])
    _check(run_plan(d, 1, D.DEALIAS_NONE, got, [g.real_shape()] * P, rt), u, TOL[prec])
    up = [rng.random(g.real_shape_padded()).astype(rt) for _ in range(P)]
    _check(run_plan(d, 0, D.DEALIAS_3_2, up, [g.complex_shape()] * P, ct),
           oracle.slab.fftn(up, N, P, dealias="3/2-rule", precision=prec), TOL[prec])


@pytest.mark.parametrize("chunks", [0, 1, 2, 4])
@pytest.mark.parametrize("P", [2, 4, 8])
@pytest.mark.parametrize("N", [(32, 32, 32), (64, 32, 16), (1024, 1024, 1024)])
def test_slab_fused_store_program_invariants(N, P, chunks):
    """Fused transport: the y (forward) / x (inverse) pass stores block `me` of every peer's receive
    buffer exactly where that peer's program expects it, inside the peer's plan-owned buffer, after
    one credit wait; exchange steps move no data; no pass loads from a peer."""
    lib = emu_util.load()
    d = _desc(D.SLAB, N, P, "double", chunks=chunks)
    d.transport = D.TRANSPORT_STORE
    for inverse in (0, 1):
        for dealias in (D.DEALIAS_NONE, D.DEALIAS_3_2, D.DEALIAS_2_3):
            if dealias == D.DEALIAS_3_2 and P > N[0] // 2:
                continue
            n = C.c_int()
            rc = lib.emu_check_p2p(C.byref(d), inverse, dealias, C.byref(n))
            assert rc == 0, (rc, inverse, dealias)
            # one flag step per forward chunk (default: one); the inverse x pass moves everything at once
            assert n.value == 1 if inverse else 1 <= n.value <= max(1, chunks)  # (a divisor of the local planes)


@pytest.mark.parametrize("kind", ["r2c", "c2c"])
@pytest.mark.parametrize("chunks", [0, 2])
@pytest.mark.parametrize("P", [2, 4])
def test_slab_fused_store_runs_in_emulator(P, chunks, kind):
    """Peer stores land in the other ranks' buffers (the emulator resolves `peer` references to that
    rank's work space, as the IPC mapping does on the device): same results as the oracle, every mode."""
    N, prec = (16, 16, 16), "double"
    rt, ct = oracle.common.dtypes(prec)
    g = oracle.slab.Geometry(N, P)
    rng = np.random.default_rng(7 + P)
    c2c = kind == "c2c"
    d = _desc(D.SLAB_C2C if c2c else D.SLAB, N, P, prec, chunks=chunks)
    d.transport = D.TRANSPORT_STORE
    if c2c:
        cs = (N[0], N[1] // P, N[2])
        A = _rand_c(rng, N, ct)
        fwd = lambda u, **k: oracle.slab.c2c_fftn(u, N, P, precision=prec, **k)
        inv = lambda fu, **k: oracle.slab.c2c_ifftn(fu, N, P, precision=prec, **k)
        padded = [_rand_c(rng, g.real_shape_padded(), ct) for _ in range(P)]
        it = ct
    else:
        cs = g.complex_shape()
        A = rng.random(N).astype(rt)
        fwd = lambda u, **k: oracle.slab.fftn(u, N, P, precision=prec, **k)
        inv = lambda fu, **k: oracle.slab.ifftn(fu, N, P, precision=prec, **k)
        padded = [rng.random(g.real_shape_padded()).astype(rt) for _ in range(P)]
        it = rt
    u = [A[g.real_local_slice(r)] for r in range(P)]
    got = run_plan(d, 0, D.DEALIAS_NONE, u, [cs] * P, ct)
    _check(got, fwd(u), TOL[prec])
    _check(run_plan(d, 1, D.DEALIAS_NONE, got, [g.real_shape()] * P, it), u, TOL[prec])
    fu = [_rand_c(rng, cs, ct) for _ in range(P)]
    _check(run_plan(d, 1, D.DEALIAS_2_3, fu, [g.real_shape()] * P, it), inv(fu, dealias="2/3-rule"), TOL[prec])
    _check(run_plan(d, 1, D.DEALIAS_3_2, fu, [g.real_shape_padded()] * P, it), inv(fu, dealias="3/2-rule"), TOL[prec])
    _check(run_plan(d, 0, D.DEALIAS_3_2, padded, [cs] * P, ct), fwd(padded, dealias="3/2-rule"), TOL[prec])


@pytest.mark.parametrize("transport", [D.TRANSPORT_NCCL, D.TRANSPORT_P2P, D.TRANSPORT_STORE])
@pytest.mark.parametrize("chunks", [0, 1, 3])
@pytest.mark.parametrize("P", [2, 4])
@pytest.mark.parametrize("kind,N", [("r2c", (16, 16, 64)), ("r2c", (8, 32, 128)), ("c2c", (16, 16, 32))])
def test_slab_kz_pipeline_runs_in_emulator(kind, N, P, chunks, transport):
    """Three-stage pipeline (B200FFT_PIPELINE_KZ): one z pass, then y(c) | exchange(c) | x(c) per kz range
    with chunk-major exchange buffers (uneven last range), every transport, every dealias mode."""
    prec = "double"
    rt, ct = oracle.common.dtypes(prec)
    g = oracle.slab.Geometry(N, P)
    rng = np.random.default_rng(11 + P + chunks)
    c2c = kind == "c2c"
    d = _desc(D.SLAB_C2C if c2c else D.SLAB, N, P, prec, chunks=chunks, pipeline=D.PIPELINE_KZ, transport=transport)
    if c2c:
        cs = (N[0], N[1] // P, N[2])
        A = _rand_c(rng, N, ct)
        fwd = lambda u, **k: oracle.slab.c2c_fftn(u, N, P, precision=prec, **k)
        inv = lambda fu, **k: oracle.slab.c2c_ifftn(fu, N, P, precision=prec, **k)
        padded = [_rand_c(rng, g.real_shape_padded(), ct) for _ in range(P)]
        it = ct
    else:
        cs = g.complex_shape()
        A = rng.random(N).astype(rt)
        fwd = lambda u, **k: oracle.slab.fftn(u, N, P, precision=prec, **k)
        inv = lambda fu, **k: oracle.slab.ifftn(fu, N, P, precision=prec, **k)
        padded = [rng.random(g.real_shape_padded()).astype(rt) for _ in range(P)]
        it = rt
    u = [A[g.real_local_slice(r)] for r in range(P)]
    got = run_plan(d, 0, D.DEALIAS_NONE, u, [cs] * P, ct)
    _check(got, fwd(u), TOL[prec])
    _check(run_plan(d, 1, D.DEALIAS_NONE, got, [g.real_shape()] * P, it), u, TOL[prec])
    fu = [_rand_c(rng, cs, ct) for _ in range(P)]
    _check(run_plan(d, 1, D.DEALIAS_2_3, fu, [g.real_shape()] * P, it), inv(fu, dealias="2/3-rule"), TOL[prec])
    _check(run_plan(d, 1, D.DEALIAS_3_2, fu, [g.real_shape_padded()] * P, it), inv(fu, dealias="3/2-rule"), TOL[prec])
    _check(run_plan(d, 0, D.DEALIAS_3_2, padded, [cs] * P, ct), fwd(padded, dealias="3/2-rule"), TOL[prec])
    if transport != D.TRANSPORT_NCCL:
        lib = emu_util.load()
        for inverse in (0, 1):
            for dealias in (D.DEALIAS_NONE, D.DEALIAS_3_2, D.DEALIAS_2_3):
                n = C.c_int()
                assert lib.emu_check_p2p(C.byref(d), inverse, dealias, C.byref(n)) == 0
                assert 1 <= n.value <= max(chunks, 4)


@pytest.mark.parametrize("layout", [D.LAYOUT_YBLOCK, D.LAYOUT_NATURAL])
@pytest.mark.parametrize("N,prec", [((16, 8, 32), "double"), ((8, 32, 16), "single"), ((4, 64, 8), "double"), ((32, 4, 16), "double"),
                                    ((8, 6, 16), "double"), ((6, 48, 8), "double"), ((2, 2, 4), "double")])
def test_slab_single_rank_layouts(N, prec, layout):
    """P = 1, slab.R2C: the y-blocked intermediate (default: z, x, y over [y block][x][y in block][kz], the row
    kernels' row map, an in-place x pass with near rows, the y pass gathering from <= 16 blocks) and the natural
    layout (z, y, x on [x][y][kz], slab.py:366-370) against the oracle -- forward, inverse, 2/3-rule, 3/2-rule both
    ways; y extents that give 16, 8, 4, 2 blocks and a single block (N1 = 6 -> padded 9 rows: odd)."""
    rt, ct = oracle.common.dtypes(prec)
    g = oracle.slab.Geometry(N, 1)
    rng = np.random.default_rng(sum(N) + layout)
    d = _desc(D.SLAB, N, 1, prec, layout=layout)
    tol = TOL[prec]
    u = [rng.random(g.real_shape()).astype(rt)]
    _check(run_plan(d, 0, D.DEALIAS_NONE, u, [g.complex_shape()], ct), oracle.slab.fftn(u, N, 1, precision=prec), tol)
    fu = [_rand_c(rng, g.complex_shape(), ct)]
    for mode, name in ((D.DEALIAS_NONE, None), (D.DEALIAS_2_3, "2/3-rule")):
        _check(run_plan(d, 1, mode, fu, [g.real_shape()], rt), oracle.slab.ifftn(fu, N, 1, dealias=name, precision=prec), tol)
    def ok(m):  # lengths with a radix plan: 2^k and 3 * 2^k
        while m % 2 == 0:
            m //= 2
        return m in (1, 3)

    if all(n % 2 == 0 and ok(3 * n // 2) for n in N):
        _check(run_plan(d, 1, D.DEALIAS_3_2, fu, [g.real_shape_padded()], rt),
               oracle.slab.ifftn(fu, N, 1, dealias="3/2-rule", precision=prec), tol)
        up = [rng.random(g.real_shape_padded()).astype(rt)]
        _check(run_plan(d, 0, D.DEALIAS_3_2, up, [g.complex_shape()], ct),
               oracle.slab.fftn(up, N, 1, dealias="3/2-rule", precision=prec), tol)


def test_slab_single_rank_c2c_keeps_the_natural_layout():
    """slab.C2C at P = 1 (rows of the z pass are complex: no row map) runs z, y, x whatever `layout` says."""
    N, prec = (16, 8, 32), "double"
    rt, ct = oracle.common.dtypes(prec)
    rng = np.random.default_rng(5)
    u = [_rand_c(rng, N, ct)]
    ref = oracle.slab.c2c_fftn(u, N, 1, precision=prec)
    for layout in (D.LAYOUT_YBLOCK, D.LAYOUT_NATURAL):
        _check(run_plan(_desc(D.SLAB_C2C, N, 1, prec, layout=layout), 0, D.DEALIAS_NONE, u, [tuple(N)], ct), ref, TOL[prec])


def test_y_blocked_program_has_no_far_pass():
    """The point of the layout, checked on the step list of the 1024^3 plan: three passes, the x pass in place
    with rows 16 x fewer elements apart than N1*Nf, the y pass reading 16 chunks and writing rows Nf apart."""
    lib = emu_util.load()
    N = (1024, 1024, 1024)
    Nf = 513
    for inverse in (0, 1):
        d = _desc(D.SLAB, N, 1, "double")
        n = lib.emu_plan_steps(C.byref(d), inverse, D.DEALIAS_NONE, None, 0)
        buf = (C.c_longlong * (8 * n))()
        assert lib.emu_plan_steps(C.byref(d), inverse, D.DEALIAS_NONE, buf, n) == n == 3
        steps = [tuple(buf[8 * i:8 * i + 8]) for i in range(n)]  # (type, n, B, J, in stride, out stride, in chunks, out chunks)
        order = [st[0] for st in steps]
        assert order == ([1, 0, 0] if not inverse else [0, 0, 2])  # 0 strided, 1 R2C, 2 C2R
        x = steps[1]
        assert (x[1], x[2], x[3]) == (1024, 16, 64 * Nf) and x[4] == x[5] == 64 * Nf
        y = steps[2] if not inverse else steps[0]
        assert (y[1], y[2], y[3]) == (1024, 1024, Nf) and y[4] == y[5] == Nf
        assert (y[6], y[7]) == ((16, 1) if not inverse else (1, 16))


@pytest.mark.parametrize("transport", [D.TRANSPORT_NCCL, D.TRANSPORT_P2P])
@pytest.mark.parametrize("chunks", [2, 4])
@pytest.mark.parametrize("comm", ["Alltoallw", "AlltoallN"])
@pytest.mark.parametrize("P,P1", [(4, None), (8, None), (8, 2)])
@pytest.mark.parametrize("alignment", ["X", "Y"])
def test_pencil_pipelined(alignment, P, P1, comm, chunks, transport):
    """Pencil programs with both exchanges cut into chunks -- local x planes for alignment X, kz sub-ranges
    with sub-range-major buffers for alignment Y -- and overlapped with the FFT passes of the neighbouring
    chunks (second stream): same results, clean schedule, peer invariants."""
    N, prec = (16, 16, 64), "double"
    rt, ct = oracle.common.dtypes(prec)
    g = oracle.pencil.Geometry(N, P, alignment, P1, comm)
    rng = np.random.default_rng(P + chunks)
    d = _desc(D.PENCIL_X if alignment == "X" else D.PENCIL_Y, N, P, prec, g.P1, g.P2, int(comm == "AlltoallN"), chunks=chunks,
              transport=transport)
    kw = dict(alignment=alignment, P1=P1, communication=comm, precision=prec)
    tol = TOL[prec]
    A = rng.random(N).astype(rt)
    u = [A[g.real_local_slice(r)] for r in range(P)]
    cshape = [g.complex_shape(r) for r in range(P)]
    _check(run_plan(d, 0, D.DEALIAS_NONE, u, cshape, ct), oracle.pencil.fftn(u, N, P, **kw), tol)
    fu = [_rand_c(rng, s, ct) for s in cshape]
    for mode, name in ((D.DEALIAS_NONE, None), (D.DEALIAS_2_3, "2/3-rule"), (D.DEALIAS_3_2, "3/2-rule")):
        shp = g.real_shape_padded() if name == "3/2-rule" else g.real_shape()
        _check(run_plan(d, 1, mode, fu, [shp] * P, rt), oracle.pencil.ifftn(fu, N, P, dealias=name, **kw), tol)
    up = [rng.random(g.real_shape_padded()).astype(rt) for _ in range(P)]
    _check(run_plan(d, 0, D.DEALIAS_3_2, up, cshape, ct), oracle.pencil.fftn(up, N, P, dealias="3/2-rule", **kw), tol)
    lib = emu_util.load()
    for inverse in (0, 1):
        for mode in (D.DEALIAS_NONE, D.DEALIAS_3_2, D.DEALIAS_2_3):
            assert lib.emu_check_schedule(C.byref(d), inverse, mode) == 0, (inverse, mode)
            if transport != D.TRANSPORT_NCCL:
                n = C.c_int()
                assert lib.emu_check_p2p(C.byref(d), inverse, mode, C.byref(n)) == 0, (inverse, mode)
                assert n.value >= 4 and n.value % 2 == 0  # two exchanges per chunk; the chunk count divides the local planes


@pytest.mark.parametrize("transport", [D.TRANSPORT_NCCL, D.TRANSPORT_P2P])
@pytest.mark.parametrize("chunks", [2, 4])
@pytest.mark.parametrize("P", [2, 4, 8])
def test_line_pipelined(P, chunks, transport):
    """line programs with the exchange cut into chunks of local rows and overlapped with the z pass."""
    N, prec = (64, 32), "double"
    rt, ct = oracle.common.dtypes(prec)
    g = oracle.line.Geometry(N, P)
    rng = np.random.default_rng(P * chunks)
    d = _desc(D.LINE, N, P, prec, chunks=chunks, transport=transport)
    A = rng.random(N).astype(rt)
    u = [A[g.real_local_slice(r)] for r in range(P)]
    cshape = [g.complex_shape(r) for r in range(P)]
    _check(run_plan(d, 0, D.DEALIAS_NONE, u, cshape, ct), oracle.line.fft2(u, N, P, precision=prec), TOL[prec])
    fu = [_rand_c(rng, s_, ct) for s_ in cshape]
    for mode, name in ((D.DEALIAS_NONE, None), (D.DEALIAS_3_2, "3/2-rule"), (D.DEALIAS_2_3, "2/3-rule")):
        shp = g.real_shape_padded() if name == "3/2-rule" else g.real_shape()
        _check(run_plan(d, 1, mode, fu, [shp] * P, rt), oracle.line.ifft2(fu, N, P, dealias=name, precision=prec), TOL[prec])
    up = [rng.random(g.real_shape_padded()).astype(rt) for _ in range(P)]
    _check(run_plan(d, 0, D.DEALIAS_3_2, up, cshape, ct), oracle.line.fft2(up, N, P, dealias="3/2-rule", precision=prec, exact=True),
           TOL[prec])
    lib = emu_util.load()
    for inverse in (0, 1):
        for mode in (D.DEALIAS_NONE, D.DEALIAS_3_2, D.DEALIAS_2_3):
            assert lib.emu_check_schedule(C.byref(d), inverse, mode) == 0
            if transport != D.TRANSPORT_NCCL:
                n = C.c_int()
                assert lib.emu_check_p2p(C.byref(d), inverse, mode, C.byref(n)) == 0
                assert n.value >= 2


@pytest.mark.parametrize("P", [2, 4])
def test_transports_and_pipelines_do_not_change_a_single_bit(P):
    """NCCL, copy engines and fused stores, x-plane and kz pipelines, any chunk count: same kernels, same
    arithmetic -- identical bits on every rank."""
    N, prec = (16, 16, 64), "double"
    rt, ct = oracle.common.dtypes(prec)
    g = oracle.slab.Geometry(N, P)
    rng = np.random.default_rng(22)
    A = rng.random(N).astype(rt)
    u = [A[g.real_local_slice(r)] for r in range(P)]
    fu = [_rand_c(rng, g.complex_shape(), ct) for _ in range(P)]
    base = _desc(D.SLAB, N, P, prec, chunks=1)
    ref_f = run_plan(base, 0, D.DEALIAS_NONE, u, [g.complex_shape()] * P, ct)
    ref_p = run_plan(base, 1, D.DEALIAS_3_2, fu, [g.real_shape_padded()] * P, rt)
    for transport in (D.TRANSPORT_NCCL, D.TRANSPORT_P2P, D.TRANSPORT_STORE):
        for pipeline in (D.PIPELINE_X, D.PIPELINE_KZ):
            for chunks in (0, 2, 3):
                d = _desc(D.SLAB, N, P, prec, chunks=chunks, pipeline=pipeline, transport=transport)
                got_f = run_plan(d, 0, D.DEALIAS_NONE, u, [g.complex_shape()] * P, ct)
                got_p = run_plan(d, 1, D.DEALIAS_3_2, fu, [g.real_shape_padded()] * P, rt)
                for r in range(P):
                    assert np.array_equal(got_f[r], ref_f[r]), (transport, pipeline, chunks, r)
                    assert np.array_equal(got_p[r], ref_p[r]), (transport, pipeline, chunks, r)
