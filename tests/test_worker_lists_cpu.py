"""The checks of the multi-GPU parity worker (tests/gpu_dist_worker.py: every class / alignment / communication layout /
dealias mode against the oracle, the goldens of the unmodified reference with the matching rank count) executed on the
CPU with 2, 4 and 8 thread-ranks through the product's own stack in the host build (tests/cpu_engine.py): the Python
layer with its transport handshakes -- copy-engine / fused-store handle exchange, NCCL bootstrap of the world and of the
pencil sub-communicators -- the C-ABI layer, plan programs and emulated kernels.  What the device adds to this is the
CUDA runtime, NVLink and the compiled kernels; the lists, the host code and the index maps are the same."""
import threading

import pytest

import cpu_engine
import gpu_dist_worker as W
from test_ref_procedures_oracle import ThreadComm, ThreadWorld

N = (32, 64, 128)


def run_ranks(P, body):
    tw = ThreadWorld(P)

    def rank_main(r):
        try:
            body(ThreadComm(tw, r))
        except BaseException as e:  # noqa: BLE001 - SystemExit from the worker's check() included
            tw.failed.append((r, repr(e)[:400]))
            tw.barrier.abort()

    ts = [threading.Thread(target=rank_main, args=(r,)) for r in range(P)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=900)
    real = [f for f in tw.failed if "BrokenBarrierError" not in f[1]]
    assert not tw.failed, real or tw.failed


@pytest.fixture
def engine(monkeypatch):
    monkeypatch.setattr(W, "DEVICE_TENSORS", False)
    monkeypatch.setattr(W, "note", lambda comm, msg: None)
    cleanup = cpu_engine.install(monkeypatch)
    yield cleanup.calls
    cleanup()


@pytest.mark.parametrize("P", [2, 4, 8])
def test_goldens_of_the_unmodified_reference(engine, P):
    run_ranks(P, W.run_golden)
    assert engine["b200fft_exec_forward"] > 0 and engine["b200fft_plan_p2p_connect"] > 0


@pytest.mark.parametrize("P", [2, 4, 8])
def test_default_transport_lists(engine, P):
    def body(comm):
        W.run_3d(comm, "slab", N, "double")
        W.run_3d(comm, "slab", N, "single", communication="Alltoall", transport="p2p", pipeline="kz")
        W.run_3d(comm, "slab", (64, 64, 64), "double", transport="store")
        W.run_line(comm, (64, 128), "double", transport="p2p")
        W.run_c2c(comm, N, "double", transport="p2p")
        if P >= 4:
            for al in "XY":
                for P1 in [None] + ([2] if P == 8 else []):
                    for cm in ("Alltoall", "Alltoallw", "AlltoallN"):
                        W.run_3d(comm, "pencil", N, "double", al, P1, cm, transport="p2p")
                W.run_3d(comm, "pencil", N, "single", al, None, "Alltoall", transport="p2p", chunks=2)
    run_ranks(P, body)


@pytest.mark.parametrize("P", [2, 4])
def test_nccl_bootstrap_lists(engine, P):
    """transport='nccl': comm.nccl_handle draws the unique id on rank 0 of the world -- and of each pencil
    sub-communicator -- and carries it with the communicator's own bcast."""
    def body(comm):
        W.run_3d(comm, "slab", N, "double", transport="nccl")
        W.run_line(comm, (64, 128), "double", transport="nccl")
        if P >= 4:
            for al in "XY":
                W.run_3d(comm, "pencil", N, "double", al, None, "Alltoallw", transport="nccl")
    run_ranks(P, body)
    assert engine["b200fft_plan_p2p_connect"] == 0


@pytest.mark.parametrize("P", [1, 2, 4, 8])
def test_bench_golden_check(engine, P):
    """bench.py's `reference_goldens` key (the stored outputs of the unmodified reference at the run's rank count,
    through the classes) on the host build: files found for every rank count the driver benches, errors at rounding."""
    import bench
    import mpifft4py_b200 as m
    results = [None] * P

    def body(comm):
        def reduce_max(x):
            return max((getattr(comm, "allgather_world", None) or (lambda v: [v]))(x))
        results[comm.Get_rank()] = bench.golden_check(m, comm, reduce_max)

    if P == 1:
        body(m.comm.COMM_SELF)
    else:
        run_ranks(P, body)
    g = results[0]
    assert all(r == g for r in results)
    assert len(g["files"]) == {1: 2, 2: 3, 4: 10, 8: 6}[P], g["files"]
    e = g["max_rel_l2"]
    assert e["double"] < 1e-13 and e["double_forward_3_2"] < 1e-12
    assert e["single"] < 2e-6 and e["single_forward_3_2"] < 1e-5


@pytest.mark.parametrize("P", [2, 4, 8])
def test_the_shared_gpu_list_itself(monkeypatch, P):
    """gpu_dist_worker.shared_gpu -- exactly what tests/test_gpu_multi.py::test_multi_rank_shared_gpu and
    tests/test_zz_shared_gpu_8.py run on GPU 0 -- with tensors in place of numpy arrays, the back-to-back round trips
    and the known answers included, on thread-ranks of the host build (host tensors where the device has CUDA ones)."""
    import torch
    import mpifft4py_b200 as m
    from test_bench_cpu_smoke import host_cuda
    monkeypatch.setattr(W, "note", lambda comm, msg: None)
    cleanup = cpu_engine.install(monkeypatch)
    host_cuda(monkeypatch)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(m.ns.Solver, "_device", lambda self: torch.device("cpu"))
    try:
        run_ranks(P, W.shared_gpu)
    finally:
        cleanup()
    assert cleanup.calls["b200fft_plan_p2p_connect"] >= 20 * P
