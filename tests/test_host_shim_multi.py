"""Multi-rank runs of the product's C-ABI layer on the CPU (tests/emu/host_shim.cpp): every rank is a thread
of this process that creates its own plan through the real entry points and executes transforms
concurrently with the others.  The stand-ins give the protocols something to run against -- CUDA IPC handles
are the pointers themselves, stream memory operations are (bounded) polls and stores on the flag words, NCCL
is an in-process rendezvous that checks counts -- so the code that cannot run without several GPUs does run:
the handle exchange, copy-engine pushes with sequence flags and credits, fused peer stores, sub-communicator
addressing by world rank, pipelined programs, NCCL send / recv groups.  Several transforms back to back per
plan exercise the credit hand-over between calls."""
import ctypes as C
import threading

import numpy as np
import pytest

import host_shim_util
import oracle
from mpifft4py_b200 import _cdefs as D

TOL = 5e-14


class Ranks(object):
    """P threads, one per rank, with the collectives the host layer needs (allgather, barrier)."""

    def __init__(self, P):
        self.P = P
        self.bar = threading.Barrier(P, timeout=60)
        self.slots = [None] * P
        self.errors = []

    def allgather(self, r, obj):
        self.slots[r] = obj
        self.bar.wait()
        out = list(self.slots)
        self.bar.wait()
        return out

    def run(self, fn):
        def body(r):
            try:
                fn(r)
            except BaseException as e:  # noqa: BLE001
                self.errors.append((r, repr(e)))
                self.bar.abort()
        th = [threading.Thread(target=body, args=(r,)) for r in range(self.P)]
        for t in th:
            t.start()
        for t in th:
            t.join(120)
        assert not self.errors, self.errors[:3]
        assert not any(t.is_alive() for t in th), "a rank is stuck"


def _make_plan(L, R, r, kind, N, P, transport, P1=1, P2=1, drop=0, **kw):
    d = D.PlanDesc()
    d.kind, d.precision = kind, D.DOUBLE
    for i, n in enumerate(N):
        d.N[i] = n
    d.nranks, d.rank, d.P1, d.P2 = P, r, P1, P2
    d.padsize, d.drop_nyquist, d.transport = 1.5, drop, transport
    for k, v in kw.items():
        setattr(d, k, v)
    comms = []
    if transport == D.TRANSPORT_NCCL:
        def make(members, me):  # the rank with the lowest world rank creates the id, the members share it
            ids = R.allgather(r, None)
            uid = C.create_string_buffer(128)
            if me == 0:
                assert L.b200fft_comm_unique_id(uid) == 0
            ids = R.allgather(r, uid.raw if me == 0 else None)
            h = C.c_void_p()
            assert L.b200fft_comm_create(C.byref(h), len(members), me, ids[members[0]]) == 0
            comms.append(h)
            return h
        if kind in (D.SLAB, D.SLAB_C2C, D.LINE):
            d.comm = make(list(range(P)), r)
        else:
            c0, c1 = r % P1, r // P1
            d.comm0 = make([c1 * P1 + q for q in range(P1)], c0)
            d.comm1 = make([q * P1 + c0 for q in range(P2)], c1)
    h = C.c_void_p()
    rc = L.b200fft_plan_create(C.byref(h), C.byref(d))
    assert rc == 0, L.b200fft_last_error()
    if transport != D.TRANSPORT_NCCL:
        mine = C.create_string_buffer(256)
        assert L.b200fft_plan_p2p_handles(h, mine) == 0, L.b200fft_last_error()
        everyone = b"".join(R.allgather(r, mine.raw))
        assert L.b200fft_plan_p2p_connect(h, everyone) == 0, L.b200fft_last_error()
    return h, comms


def _exec(L, h, inverse, mode, src, dst):
    fn = L.b200fft_exec_inverse if inverse else L.b200fft_exec_forward
    L.shim_next_epoch()  # an event recorded by an earlier call must not satisfy a wait of this one
    rc = fn(h, C.c_void_p(src.ctypes.data), C.c_void_p(dst.ctypes.data), mode, None)
    assert rc == 0, (rc, L.b200fft_last_error())
    return dst


SLAB_CASES = [(D.TRANSPORT_NCCL, D.PIPELINE_X, 0), (D.TRANSPORT_NCCL, D.PIPELINE_X, 2),
              (D.TRANSPORT_P2P, D.PIPELINE_X, 0), (D.TRANSPORT_P2P, D.PIPELINE_X, 4),
              (D.TRANSPORT_STORE, D.PIPELINE_X, 0), (D.TRANSPORT_STORE, D.PIPELINE_X, 2),
              (D.TRANSPORT_NCCL, D.PIPELINE_KZ, 3), (D.TRANSPORT_P2P, D.PIPELINE_KZ, 2),
              (D.TRANSPORT_STORE, D.PIPELINE_KZ, 2)]


@pytest.mark.parametrize("transport,pipeline,chunks", SLAB_CASES)
@pytest.mark.parametrize("P", [2, 4])
def test_slab_ranks_as_threads(P, transport, pipeline, chunks):
    L = host_shim_util.load()
    N = (16, 16, 32)
    g = oracle.slab.Geometry(N, P)
    rng = np.random.default_rng(5)
    A = rng.random(N)
    u = [np.ascontiguousarray(A[g.real_local_slice(r)]) for r in range(P)]
    ref = oracle.slab.fftn(u, N, P)
    fu = [(rng.standard_normal(g.complex_shape()) + 1j * rng.standard_normal(g.complex_shape())) for _ in range(P)]
    refs = {name: oracle.slab.ifftn(fu, N, P, dealias=name) for name in (None, "2/3-rule", "3/2-rule")}
    R = Ranks(P)

    def rank(r):
        h, comms = _make_plan(L, R, r, D.SLAB, N, P, transport, pipeline=pipeline, chunks=chunks)
        for rep in range(3):  # back to back: sequence numbers and credits carry over from call to call
            c = _exec(L, h, 0, D.DEALIAS_NONE, u[r], np.full(g.complex_shape(), np.nan, dtype=np.complex128))
            assert oracle.rel_l2(c, ref[r]) <= TOL, (r, rep)
            back = _exec(L, h, 1, D.DEALIAS_NONE, c, np.full(g.real_shape(), np.nan))
            assert oracle.rel_l2(back, u[r]) <= TOL, (r, rep)
        for mode, name in ((D.DEALIAS_2_3, "2/3-rule"), (D.DEALIAS_3_2, "3/2-rule")):
            shp = g.real_shape_padded() if name == "3/2-rule" else g.real_shape()
            got = _exec(L, h, 1, mode, fu[r], np.full(shp, np.nan))
            assert oracle.rel_l2(got, refs[name][r]) <= TOL, (r, name)
        R.bar.wait()
        assert L.b200fft_plan_destroy(h) == 0
        for c in comms:
            L.b200fft_comm_destroy(c)

    R.run(rank)


@pytest.mark.parametrize("chunks", [0, 2])
@pytest.mark.parametrize("transport", [D.TRANSPORT_NCCL, D.TRANSPORT_P2P, D.TRANSPORT_STORE])
@pytest.mark.parametrize("kind,P,P1,P2,drop", [(D.PENCIL_X, 4, 2, 2, 0), (D.PENCIL_Y, 4, 2, 2, 0), (D.PENCIL_X, 8, 4, 2, 1),
                                                (D.PENCIL_Y, 8, 2, 4, 0)])
def test_pencil_ranks_as_threads(kind, P, P1, P2, drop, transport, chunks):
    if chunks and transport == D.TRANSPORT_STORE:
        pytest.skip("the fused transport is not pipelined for pencil plans")
    L = host_shim_util.load()
    N = (16, 16, 32)
    al = "X" if kind == D.PENCIL_X else "Y"
    comm = "AlltoallN" if drop else "Alltoallw"
    g = oracle.pencil.Geometry(N, P, al, P1, comm)
    kw = dict(alignment=al, P1=P1, communication=comm)
    rng = np.random.default_rng(6)
    A = rng.random(N)
    u = [np.ascontiguousarray(A[g.real_local_slice(r)]) for r in range(P)]
    cshape = [g.complex_shape(r) for r in range(P)]
    ref = oracle.pencil.fftn(u, N, P, **kw)
    rt_ref = oracle.pencil.ifftn(ref, N, P, **kw)
    R = Ranks(P)

    def rank(r):
        h, comms = _make_plan(L, R, r, kind, N, P, transport, P1=P1, P2=P2, drop=drop, chunks=chunks)
        for rep in range(3):
            c = _exec(L, h, 0, D.DEALIAS_NONE, u[r], np.full(cshape[r], np.nan, dtype=np.complex128))
            assert oracle.rel_l2(c, ref[r]) <= TOL, (r, rep)
            back = _exec(L, h, 1, D.DEALIAS_NONE, c, np.full(g.real_shape(), np.nan))
            assert oracle.rel_l2(back, rt_ref[r]) <= TOL, (r, rep)
        up = _exec(L, h, 1, D.DEALIAS_3_2, ref[r], np.full(g.real_shape_padded(), np.nan))
        assert oracle.rel_l2(up, oracle_padded[r]) <= TOL
        R.bar.wait()
        assert L.b200fft_plan_destroy(h) == 0
        for c in comms:
            L.b200fft_comm_destroy(c)

    oracle_padded = oracle.pencil.ifftn(ref, N, P, dealias="3/2-rule", **kw)
    R.run(rank)


@pytest.mark.parametrize("chunks", [0, 2])
@pytest.mark.parametrize("transport", [D.TRANSPORT_NCCL, D.TRANSPORT_P2P, D.TRANSPORT_STORE])
def test_line_ranks_as_threads(transport, chunks):
    L = host_shim_util.load()
    N, P = (32, 64), 4
    g = oracle.line.Geometry(N, P)
    A = np.random.default_rng(7).random(N)
    u = [np.ascontiguousarray(A[g.real_local_slice(r)]) for r in range(P)]
    ref = oracle.line.fft2(u, N, P)
    R = Ranks(P)

    def rank(r):
        h, comms = _make_plan(L, R, r, D.LINE, N, P, transport, chunks=chunks)
        for rep in range(3):
            c = _exec(L, h, 0, D.DEALIAS_NONE, u[r], np.full(g.complex_shape(r), np.nan, dtype=np.complex128))
            assert oracle.rel_l2(c, ref[r]) <= TOL
            assert oracle.rel_l2(_exec(L, h, 1, D.DEALIAS_NONE, c, np.full(g.real_shape(), np.nan)), u[r]) <= TOL
        R.bar.wait()
        assert L.b200fft_plan_destroy(h) == 0
        for c in comms:
            L.b200fft_comm_destroy(c)

    R.run(rank)


@pytest.mark.parametrize("transport,pipeline", [(D.TRANSPORT_P2P, D.PIPELINE_X), (D.TRANSPORT_STORE, D.PIPELINE_X),
                                                (D.TRANSPORT_STORE, D.PIPELINE_KZ), (D.TRANSPORT_P2P, D.PIPELINE_KZ)])
def test_skewed_ranks_many_transforms(transport, pipeline):
    """30 transforms of mixed kinds per plan with the ranks deliberately out of step (random sleeps): a rank
    that runs ahead must be held by the credit / sequence flags, never overwrite a buffer a slower peer
    still reads, and nobody may deadlock (the stand-in waits are bounded)."""
    import time
    L = host_shim_util.load()
    N, P = (16, 16, 32), 4
    g = oracle.slab.Geometry(N, P)
    rng = np.random.default_rng(8)
    A = rng.random(N)
    u = [np.ascontiguousarray(A[g.real_local_slice(r)]) for r in range(P)]
    ref = oracle.slab.fftn(u, N, P)
    fu = [(rng.standard_normal(g.complex_shape()) + 1j * rng.standard_normal(g.complex_shape())) for _ in range(P)]
    inv = {m: oracle.slab.ifftn(fu, N, P, dealias=n) for m, n in ((D.DEALIAS_NONE, None), (D.DEALIAS_3_2, "3/2-rule"))}
    order = np.random.default_rng(9).integers(0, 3, size=30)  # the same sequence of calls on every rank
    R = Ranks(P)

    def rank(r):
        h, _ = _make_plan(L, R, r, D.SLAB, N, P, transport, pipeline=pipeline, chunks=2)
        nap = np.random.default_rng(100 + r)
        for what in order:
            time.sleep(float(nap.random()) * 0.004 * (r + 1))
            if what == 0:
                c = _exec(L, h, 0, D.DEALIAS_NONE, u[r], np.full(g.complex_shape(), np.nan, dtype=np.complex128))
                assert oracle.rel_l2(c, ref[r]) <= TOL
            else:
                mode = D.DEALIAS_NONE if what == 1 else D.DEALIAS_3_2
                shp = g.real_shape_padded() if what == 2 else g.real_shape()
                got = _exec(L, h, 1, mode, fu[r], np.full(shp, np.nan))
                assert oracle.rel_l2(got, inv[mode][r]) <= TOL
        R.bar.wait()
        assert L.b200fft_plan_destroy(h) == 0

    R.run(rank)


@pytest.mark.parametrize("transport,pipeline,chunks", [(D.TRANSPORT_NCCL, D.PIPELINE_X, 2), (D.TRANSPORT_P2P, D.PIPELINE_X, 2),
                                                       (D.TRANSPORT_STORE, D.PIPELINE_X, 0), (D.TRANSPORT_STORE, D.PIPELINE_KZ, 2)])
def test_c2c_single_precision_ranks_as_threads(transport, pipeline, chunks):
    """slab.C2C in single precision on 4 ranks: complex rows in the z pass, keep-mode truncations."""
    L = host_shim_util.load()
    N, P, prec = (16, 16, 16), 4, "single"
    rt, ct = oracle.common.dtypes(prec)
    g = oracle.slab.GeometryC2C(N, P)
    rng = np.random.default_rng(10)
    A = (rng.standard_normal(N) + 1j * rng.standard_normal(N)).astype(ct)
    u = [np.ascontiguousarray(A[g.real_local_slice(r)]) for r in range(P)]
    ref = oracle.slab.c2c_fftn(u, N, P, precision=prec)
    padded = oracle.slab.c2c_ifftn(ref, N, P, dealias="3/2-rule", precision=prec)
    R = Ranks(P)

    def rank(r):
        d_extra = dict(pipeline=pipeline, chunks=chunks, precision=D.SINGLE)
        h, comms = _make_plan(L, R, r, D.SLAB_C2C, N, P, transport, **d_extra)
        for rep in range(2):
            c = _exec(L, h, 0, D.DEALIAS_NONE, u[r], np.full(g.complex_shape(), np.nan, dtype=ct))
            assert oracle.rel_l2(c, ref[r]) <= 5e-6
            assert oracle.rel_l2(_exec(L, h, 1, D.DEALIAS_NONE, c, np.full(g.real_shape(), np.nan, dtype=ct)), u[r]) <= 5e-6
            up = _exec(L, h, 1, D.DEALIAS_3_2, ref[r], np.full(g.real_shape_padded(), np.nan, dtype=ct))
            assert oracle.rel_l2(up, padded[r]) <= 5e-6
        R.bar.wait()
        assert L.b200fft_plan_destroy(h) == 0
        for c in comms:
            L.b200fft_comm_destroy(c)

    R.run(rank)


@pytest.mark.parametrize("kind,P,P1,P2", [(D.SLAB, 4, 1, 1), (D.PENCIL_X, 8, 4, 2), (D.LINE, 4, 1, 1)])
def test_copy_engine_pipelined_for_every_class(kind, P, P1, P2):
    """Copy-engine transport with two chunks for slab, pencil X (sub-communicators, world-rank flag indexing) and
    line plans: three round trips back to back (flags posted by the flag kernel's host stand-in, credits carried over)."""
    L = host_shim_util.load()
    if kind == D.LINE:
        N = (32, 64)
        g = oracle.line.Geometry(N, P)
        fwd = lambda u: oracle.line.fft2(u, N, P)
        cshape = [g.complex_shape(r) for r in range(P)]
    elif kind == D.SLAB:
        N = (16, 16, 32)
        g = oracle.slab.Geometry(N, P)
        fwd = lambda u: oracle.slab.fftn(u, N, P)
        cshape = [g.complex_shape()] * P
    else:
        N = (16, 16, 32)
        g = oracle.pencil.Geometry(N, P, "X", P1, "Alltoallw")
        fwd = lambda u: oracle.pencil.fftn(u, N, P, alignment="X", P1=P1, communication="Alltoallw")
        cshape = [g.complex_shape(r) for r in range(P)]
    A = np.random.default_rng(11).random(N)
    u = [np.ascontiguousarray(A[g.real_local_slice(r)]) for r in range(P)]
    ref = fwd(u)
    R = Ranks(P)

    def rank(r):
        h, _ = _make_plan(L, R, r, kind, N, P, D.TRANSPORT_P2P, P1=P1, P2=P2, chunks=2)
        for rep in range(3):
            c = _exec(L, h, 0, D.DEALIAS_NONE, u[r], np.full(cshape[r], np.nan, dtype=np.complex128))
            assert oracle.rel_l2(c, ref[r]) <= TOL
            assert oracle.rel_l2(_exec(L, h, 1, D.DEALIAS_NONE, c, np.full(g.real_shape(), np.nan)), u[r]) <= TOL
        R.bar.wait()
        assert L.b200fft_plan_destroy(h) == 0

    R.run(rank)
