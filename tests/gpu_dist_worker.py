"""Multi-GPU worker (one process per GPU, NCCL): every class / alignment / dealias mode against the
oracle on seeded inputs, plus the reference's golden vectors.  Launched by torch.distributed.run
from tests/test_gpu_multi.py (or directly under gpurun --gpus N)."""
import glob
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import mpifft4py_b200 as m  # noqa: E402
import oracle  # noqa: E402
import ref_procedures  # noqa: E402
from mpifft4py_b200.comm import world  # noqa: E402

TOL = {"double": 1e-12, "single": 1e-5}
DEVICE_TENSORS = True  # tests/test_worker_lists_cpu.py runs these checks on the host build of the engine: numpy callers only
L3 = np.array([2 * np.pi] * 3)


def rand_c(rng, shape, ct):
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(ct)


def check(name, got, ref, tol, rank):
    err = oracle.rel_l2(got, ref)
    if not (err <= tol) or got.shape != ref.shape:
        raise SystemExit("rank %d: %s: rel L2 %.3e > %.1e (shapes %r %r)" % (rank, name, err, tol, got.shape, ref.shape))


_T0 = [None]


def note(comm, msg):
    """progress line of rank 0 (a run that is cut off by a timeout still shows how far it got)"""
    import time
    if _T0[0] is None:
        _T0[0] = time.time()
    if comm.Get_rank() == 0:
        print("[%6.1f s] %s" % (time.time() - _T0[0], msg), flush=True)


def run_3d(comm, kind, N, prec, alignment=None, P1=None, communication=None, transport=None, pipeline=None, chunks=0):
    P, r = comm.Get_size(), comm.Get_rank()
    note(comm, "run_3d %s %s P1=%s %s %s transport=%s pipeline=%s chunks=%s" % (kind, alignment, P1, communication, prec, transport,
                                                                            pipeline, chunks))
    rt, ct = oracle.common.dtypes(prec)
    tol = TOL[prec]
    rng = np.random.default_rng(99)  # same stream on every rank: global arrays are identical
    if kind == "slab":
        F = m.Slab_R2C(np.array(N), L3, comm, prec, communication=communication or "Alltoallw")
        if transport:
            F.transport = transport  # read when the device plan is created (first transform)
        if pipeline:
            F.exchange_pipeline = pipeline
        g = oracle.slab.Geometry(N, P)
        cshape = [g.complex_shape()] * P
        fwd = lambda u, d=None: oracle.slab.fftn(u, N, P, dealias=d, precision=prec)
        inv = lambda f, d=None: oracle.slab.ifftn(f, N, P, dealias=d, precision=prec)
    else:
        F = m.Pencil_R2C(np.array(N), L3, comm, prec, P1=P1, communication=communication, alignment=alignment)
        if transport:
            F.transport = transport
        if chunks:
            F.exchange_chunks = chunks  # pipelined pencil programs
        g = oracle.pencil.Geometry(N, P, alignment, P1, communication)
        cshape = [g.complex_shape(q) for q in range(P)]
        kw = dict(alignment=alignment, P1=P1, communication=communication, precision=prec)
        fwd = lambda u, d=None: oracle.pencil.fftn(u, N, P, dealias=d, **kw)
        inv = lambda f, d=None: oracle.pencil.ifftn(f, N, P, dealias=d, **kw)
    tag = "%s %s %s %s chunks=%d %s P=%d" % (kind, alignment, communication, transport, chunks, prec, P)
    assert tuple(int(s) for s in F.complex_shape()) == tuple(cshape[r])
    A = rng.random(N).astype(rt)
    u = [np.ascontiguousarray(A[g.real_local_slice(q)]) for q in range(P)]
    c = F.fftn(u[r], np.zeros(cshape[r], dtype=ct))
    check(tag + " fftn", c, fwd(u)[r], tol, r)
    a = F.ifftn(c, np.zeros(F.real_shape(), dtype=rt))
    # 'AlltoallN' drops the Nyquist plane (pencil.py:909-910), so its round trip is the oracle's, not u
    check(tag + " roundtrip", a, inv(fwd(u))[r] if communication == "AlltoallN" else u[r], tol, r)
    fu = [rand_c(rng, s, ct) for s in cshape]
    for d in (None, "2/3-rule", "3/2-rule"):
        shp = F.real_shape_padded() if d == "3/2-rule" else F.real_shape()
        got = F.ifftn(fu[r], np.zeros(shp, dtype=rt), dealias=d)
        check(tag + " ifftn %s" % d, got, inv(fu, d)[r], tol, r)
    up = [rng.random(g.real_shape_padded()).astype(rt) for _ in range(P)]
    got = F.fftn(up[r], np.zeros(cshape[r], dtype=ct), dealias="3/2-rule")
    check(tag + " fftn 3/2", got, fwd(up, "3/2-rule")[r], tol, r)
    if not DEVICE_TENSORS:
        return
    # CUDA tensors in place of numpy arrays
    tu = torch.from_numpy(u[r]).cuda()
    tf = torch.zeros(tuple(int(s) for s in cshape[r]), dtype=torch.complex128 if prec == "double" else torch.complex64,
                     device="cuda")
    F.fftn(tu, tf)
    check(tag + " fftn tensors", tf.cpu().numpy(), c, 1e-15, r)
    if transport:
        assert F.transport_used == transport, (F.transport_used, transport)
        # back-to-back transforms without host synchronisation: the credit / sequence flags alone keep
        # the peers out of buffers that are still being read
        ta = torch.empty_like(tu)
        for _ in range(6):
            F.fftn(tu, tf)
            F.ifftn(tf, ta)
        # ('AlltoallN': the round trip drops the Nyquist plane, so it reproduces the single round trip `a`, not u)
        check(tag + " 6 round trips", ta.cpu().numpy(), a if communication == "AlltoallN" else u[r], tol, r)


def run_line(comm, N, prec, transport=None):
    P, r = comm.Get_size(), comm.Get_rank()
    rt, ct = oracle.common.dtypes(prec)
    tol = TOL[prec]
    rng = np.random.default_rng(17)
    F = m.Line_R2C(np.array(N), L3[:2], comm, prec)
    if transport:
        F.transport = transport
    g = oracle.line.Geometry(N, P)
    cshape = [g.complex_shape(q) for q in range(P)]
    A = rng.random(N).astype(rt)
    u = [np.ascontiguousarray(A[g.real_local_slice(q)]) for q in range(P)]
    c = F.fft2(u[r], np.zeros(cshape[r], dtype=ct))
    check("line fft2", c, oracle.line.fft2(u, N, P, precision=prec)[r], tol, r)
    check("line roundtrip", F.ifft2(c, np.zeros(F.real_shape(), dtype=rt)), u[r], tol, r)
    fu = [rand_c(rng, s, ct) for s in cshape]
    for d in (None, "2/3-rule", "3/2-rule"):
        shp = F.real_shape_padded() if d == "3/2-rule" else F.real_shape()
        got = F.ifft2(fu[r], np.zeros(shp, dtype=rt), dealias=d)
        check("line ifft2 %s" % d, got, oracle.line.ifft2(fu, N, P, dealias=d, precision=prec)[r], tol, r)
    up = [rng.random(g.real_shape_padded()).astype(rt) for _ in range(P)]
    got = F.fft2(up[r], np.zeros(cshape[r], dtype=ct), dealias="3/2-rule")
    check("line fft2 3/2", got, oracle.line.fft2(up, N, P, dealias="3/2-rule", precision=prec, exact=True)[r], tol, r)


def run_golden(comm):
    P, r = comm.Get_size(), comm.Get_rank()
    for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz"))):
        z = np.load(path)
        meta = json.loads(str(z["meta"]))
        if meta["P"] != P:
            continue
        prec = meta["precision"]
        tol = TOL[prec]
        N = meta["N"]
        if meta["kind"] == "slab":
            F = m.Slab_R2C(np.array(N), L3, comm, prec, communication=meta["communication"])
            fwd, inv = F.fftn, F.ifftn
        elif meta["kind"] == "pencil":
            F = m.Pencil_R2C(np.array(N), L3, comm, prec, P1=meta["P1"], communication=meta["communication"],
                             alignment=meta["alignment"])
            fwd, inv = F.fftn, F.ifftn
        else:
            F = m.Line_R2C(np.array(N), L3[:2], comm, prec)
            fwd, inv = F.fft2, F.ifft2
        info = meta["ranks"][r]
        rs = tuple(slice(*s) for s in info["real_local_slice"])
        rps = tuple(slice(*s) for s in info["real_local_slice_padded"])
        cs = tuple(slice(*s) for s in info["complex_local_slice"])
        A, Cg = z["A"], z["C"]
        name = os.path.basename(path)
        c = fwd(np.ascontiguousarray(A[rs]), np.zeros(Cg[cs].shape, dtype=Cg.dtype))
        check(name + " C", c, Cg[cs], tol, r)
        check(name + " A2", inv(np.ascontiguousarray(Cg[cs]), np.zeros(A[rs].shape, dtype=A.dtype)), z["A2"][rs], tol, r)
        Cin = Cg.copy()
        if meta["kind"] == "line":
            Cin[-N[0] // 2] = 0
        Ap = z["Ap"]
        check(name + " Ap", inv(np.ascontiguousarray(Cin[cs]), np.zeros(Ap[rps].shape, dtype=Ap.dtype), dealias="3/2-rule"),
              Ap[rps], tol, r)
        check(name + " Cp", fwd(np.ascontiguousarray(Ap[rps]), np.zeros(Cg[cs].shape, dtype=Cg.dtype), dealias="3/2-rule"),
              z["Cp"][cs], 10 * tol, r)
        if meta["has23"]:
            check(name + " A23", inv(np.ascontiguousarray(Cg[cs]), np.zeros(A[rs].shape, dtype=A.dtype), dealias="2/3-rule"),
                  z["A23"][rs], tol, r)


def run_c2c(comm, N, prec, transport=None, pipeline=None):
    """slab.C2C (slab.py:538-825) on all ranks against the oracle, every dealias mode."""
    P, r = comm.Get_size(), comm.Get_rank()
    rt, ct = oracle.common.dtypes(prec)
    tol = TOL[prec]
    rng = np.random.default_rng(41)
    F = m.Slab_C2C(np.array(N), L3, comm, prec)
    if transport:
        F.transport = transport
    if pipeline:
        F.exchange_pipeline = pipeline
    g = oracle.slab.GeometryC2C(N, P)
    A = rand_c(rng, N, ct)
    u = [np.ascontiguousarray(A[g.real_local_slice(q)]) for q in range(P)]
    c = F.fftn(u[r], np.zeros(g.complex_shape(), dtype=ct))
    check("c2c fftn", c, oracle.slab.c2c_fftn(u, N, P, precision=prec)[r], tol, r)
    check("c2c roundtrip", F.ifftn(c, np.zeros(g.real_shape(), dtype=ct)), u[r], tol, r)
    fu = [rand_c(rng, g.complex_shape(), ct) for _ in range(P)]
    for d in (None, "2/3-rule", "3/2-rule"):
        shp = g.real_shape_padded() if d == "3/2-rule" else g.real_shape()
        got = F.ifftn(fu[r], np.zeros(shp, dtype=ct), dealias=d)
        check("c2c ifftn %s" % d, got, oracle.slab.c2c_ifftn(fu, N, P, dealias=d, precision=prec)[r], tol, r)
    up = [rand_c(rng, g.real_shape_padded(), ct) for _ in range(P)]
    got = F.fftn(up[r], np.zeros(g.complex_shape(), dtype=ct), dealias="3/2-rule")
    check("c2c fftn 3/2", got, oracle.slab.c2c_fftn(up, N, P, dealias="3/2-rule", precision=prec)[r], tol, r)


def run_known_answer(comm):
    """Taylor-Green kinetic energy of the reference's demo (demo/spectral_dns_solver.py:103-105) with the
    slab transforms distributed over all ranks, CUDA tensors end to end."""
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    import spectral_dns_solver as sds
    N = np.array([32, 32, 32], dtype=int)
    F = m.Slab_R2C(N, L3, comm, "double")
    k = sds.solve(F, torch, lambda a: torch.from_numpy(a).cuda(), N)
    k = comm.reduce(k)
    if comm.Get_rank() == 0 and round(k - sds.KNOWN_ANSWER, 7) != 0:
        raise SystemExit("Taylor-Green known answer: got %.12f, expected %.12f" % (k, sds.KNOWN_ANSWER))


def run_known_answer_kernels(comm):
    """The same known answer through mpifft4py_b200.ns.Solver: the library's three elementwise kernels per RK stage
    around the distributed transforms (each rank holds its own wavenumber vectors)."""
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    import spectral_dns_solver as sds
    N = np.array([32, 32, 32], dtype=int)
    F = m.Slab_R2C(N, L3, comm, "double")
    S = m.ns.Solver(F, nu=0.000625, dt=0.01)
    X = [torch.from_numpy(np.ascontiguousarray(np.broadcast_to(x, F.real_shape()))).cuda() for x in F.get_local_mesh()]
    S.set_velocity(torch.stack([torch.sin(X[0]) * torch.cos(X[1]) * torch.cos(X[2]),
                                -torch.cos(X[0]) * torch.sin(X[1]) * torch.cos(X[2]), torch.zeros_like(X[0])]))
    for _ in range(10):
        S.step()
    k = comm.reduce(S.kinetic_energy())
    if comm.Get_rank() == 0 and round(k - sds.KNOWN_ANSWER, 7) != 0:
        raise SystemExit("Taylor-Green known answer (kernels): got %.12f, expected %.12f" % (k, sds.KNOWN_ANSWER))


def shared_gpu(comm):
    """All ranks on ONE GPU (tests/test_gpu_multi.py::test_multi_rank_shared_gpu): what a single-GPU box can say
    about the distributed half of the path.  The ranks are separate processes with separate CUDA contexts on the same
    device; the copy-engine and fused-store transports work unchanged there (CUDA IPC mappings of the peers' work
    buffers and flag words, cudaMemcpyAsync pushes, stream memory operations) -- only NCCL refuses two ranks on one
    device, so torch.distributed runs over gloo and the NCCL transport is left to the multi-GPU runs.  Plan programs,
    per-peer chunk layouts, sub-communicator addressing, pipelines, sequence flags and credits are the very code an
    8-GPU box executes; what this mode cannot show is NVLink."""
    P = comm.Get_size()
    N = (32, 64, 128)
    run_3d(comm, "slab", N, "double", transport="p2p")
    run_line(comm, (64, 128), "double", transport="p2p")
    if P < 8:  # (eight contexts on one device take turns: the 8-rank run keeps to what only it can show -- the 4x2 and
        #         2x4 pencil grids and the 8-rank goldens)
        run_3d(comm, "slab", N, "single", transport="p2p")
        run_3d(comm, "slab", N, "double", communication="Alltoall", transport="p2p", pipeline="kz")
        run_3d(comm, "slab", (64, 64, 64), "double", transport="store")
        run_line(comm, (64, 128), "single", transport="store")
        run_c2c(comm, N, "double", transport="p2p")
    if P >= 4:
        grids = [None] + ([2] if P == 8 else [])
        for al in "XY":
            for P1 in grids:
                for cm in ("Alltoall", "Alltoallw", "AlltoallN"):
                    run_3d(comm, "pencil", N, "double", al, P1, cm, transport="p2p")
            run_3d(comm, "pencil", N, "single", al, None, "Alltoall", transport="p2p", chunks=2)
    note(comm, "goldens")
    run_golden(comm)
    # the reference's own test procedures (tests/test_FFT.py) on this communicator: all 16 + 2 + 2 fixture parameters
    ref_procedures.run_all(comm, lambda msg: note(comm, msg))
    if P < 8:
        note(comm, "known answer")
        run_known_answer(comm)
        run_known_answer_kernels(comm)
    note(comm, "done")


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if "--share-gpu" in sys.argv:
        os.environ["B200FFT_STRICT_TRANSPORT"] = "1"  # a transport that cannot be set up is an error, not an NCCL fallback
        torch.cuda.set_device(0)
        dist.init_process_group("gloo")
        comm = world()
        shared_gpu(comm)
        comm.barrier()
        dist.destroy_process_group()
        print("GPU_WORKER_OK", local)
        return
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = world()
    P = comm.Get_size()
    N = (32, 64, 128)
    if len(sys.argv) > 1 and sys.argv[1] == "--transport":
        # one slab transport x pipeline only (tests/test_zz_gpu_transports.py): "store" = fused peer stores;
        # "kz" = three-stage pipeline over kz ranges
        tr, pipe = sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else None)
        if pipe == "pencil-chunks":  # pipelined pencil programs (x-plane chunks for 'X', kz sub-ranges for 'Y')
            if P >= 4:
                for al in "XY":
                    for cm in ("Alltoallw", "AlltoallN"):
                        for ch in (2, 4):
                            run_3d(comm, "pencil", (32, 64, 128), "double", al, None, cm, transport=tr, chunks=ch)
                run_3d(comm, "pencil", (32, 64, 128), "single", "Y", None, "Alltoall", transport=tr, chunks=2)
            comm.barrier()
            dist.destroy_process_group()
            print("GPU_WORKER_OK", local)
            return
        for prec in ("double", "single"):
            run_3d(comm, "slab", N, prec, transport=tr, pipeline=pipe)
        run_3d(comm, "slab", (64, 64, 64), "double", transport=tr, pipeline=pipe)
        run_c2c(comm, N, "double", transport=tr, pipeline=pipe)
        if tr != "nccl" and pipe != "kz":  # pencil / line over the same peer mappings (their default is NCCL)
            run_line(comm, (64, 128), "double", transport=tr)
            if P >= 4:
                for al in "XY":
                    for cm in ("Alltoall", "AlltoallN"):
                        run_3d(comm, "pencil", N, "double", al, None, cm, transport=tr)
                if P == 8:
                    run_3d(comm, "pencil", N, "single", "X", 2, "Alltoall", transport=tr)
        comm.barrier()
        dist.destroy_process_group()
        print("GPU_WORKER_OK", local)
        return
    # defaults (copy-engine transport for every class) ...
    for prec in ("double", "single"):
        run_3d(comm, "slab", N, prec)
        run_line(comm, (64, 128), prec)
    run_3d(comm, "slab", N, "double", communication="Alltoall")
    run_c2c(comm, N, "double")
    run_c2c(comm, (32, 32, 32), "single")
    if P >= 4:
        grids = [None] + ([2] if P == 8 else [])
        for al in "XY":
            for P1 in grids:
                for cm in ("Alltoall", "Alltoallw", "AlltoallN"):
                    run_3d(comm, "pencil", N, "double", al, P1, cm)
            run_3d(comm, "pencil", N, "single", al, None, "Alltoall")
    # ... and NCCL send / recv for one object of each class
    run_3d(comm, "slab", N, "double", transport="nccl")
    run_line(comm, (64, 128), "double", transport="nccl")
    if P >= 4:
        for al in "XY":
            run_3d(comm, "pencil", N, "double", al, None, "Alltoallw", transport="nccl")
    note(comm, "goldens")
    run_golden(comm)
    ref_procedures.run_all(comm, lambda msg: note(comm, msg))
    note(comm, "known answer")
    run_known_answer(comm)
    run_known_answer_kernels(comm)
    note(comm, "done")
    comm.barrier()
    dist.destroy_process_group()
    print("GPU_WORKER_OK", local)


if __name__ == "__main__":
    main()
