"""Worker for tests/test_dist_gloo.py: exercises mpifft4py_b200.comm.TorchComm and the multi-rank
host logic over torch.distributed (gloo on CPU).  Launched by torch.distributed.run."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import mpifft4py_b200 as m  # noqa: E402
import oracle  # noqa: E402
from mpifft4py_b200.comm import TorchComm, world  # noqa: E402


def main():
    dist.init_process_group("gloo")
    comm = world()
    assert isinstance(comm, TorchComm)
    P, r = comm.Get_size(), comm.Get_rank()
    assert P == int(os.environ["WORLD_SIZE"]) and r == int(os.environ["RANK"])
    # Bcast / bcast / reduce / barrier: the calls tests/test_FFT.py:77-78 and the demos make
    a = np.arange(12, dtype=np.float64).reshape(3, 4) if r == 0 else np.zeros((3, 4))
    comm.Bcast(a, root=0)
    assert np.array_equal(a, np.arange(12).reshape(3, 4))
    z = (np.arange(6) * (1 - 2j)).reshape(2, 3) if r == 0 else np.zeros((2, 3), dtype=complex)   # complex arrays too
    comm.Bcast(z, root=0)                                                                          # (tests/test_FFT.py:78)
    assert np.array_equal(z, (np.arange(6) * (1 - 2j)).reshape(2, 3))
    z0 = np.array(3 - 4j) if r == 0 else np.array(0j)            # 0-d
    comm.Bcast(z0, root=0)
    assert z0 == 3 - 4j
    z32 = np.full((4,), 1 + 1j, dtype=np.complex64) if r == 0 else np.zeros((4,), dtype=np.complex64)
    comm.Bcast(z32, root=0)
    assert z32.dtype == np.complex64 and np.all(z32 == 1 + 1j)
    assert comm.bcast({"id": 7} if r == 0 else None, root=0) == {"id": 7}
    tot = comm.reduce(float(r + 1))
    if r == 0:
        assert tot == P * (P + 1) / 2
    assert comm.reduce(r, op="MIN", root=0) in (0, None)
    comm.barrier()
    # slab + line geometry per rank equals the oracle's
    N = np.array([8, 16, 32])
    L = np.array([2 * np.pi] * 3)
    F = m.Slab_R2C(N, L, comm, "double")
    g = oracle.slab.Geometry(N, P)
    assert tuple(F.real_local_slice()) == tuple(g.real_local_slice(r))
    assert tuple(F.complex_local_slice()) == tuple(g.complex_local_slice(r))
    assert tuple(F.real_local_slice(1.5)) == tuple(g.real_local_slice(r, 1.5))
    F2 = m.Line_R2C(N[:2], L[:2], comm, "single")
    g2 = oracle.line.Geometry(N[:2], P)
    assert tuple(int(x) for x in F2.complex_shape()) == g2.complex_shape(r)
    assert tuple(F2.complex_local_slice()) == tuple(g2.complex_local_slice(r))
    # pencil: Split into comm0 / comm1 (pencil.py:192-195)
    if P >= 4:
        for al in "XY":
            Fp = m.Pencil_R2C(N, L, comm, "double", alignment=al, communication="Alltoallw")
            gp = oracle.pencil.Geometry(N, P, al, None, "Alltoallw")
            assert (Fp.comm0_rank, Fp.comm1_rank) == gp.coords(r)
            assert Fp.comm0.Get_size() == gp.P1 and Fp.comm1.Get_size() == gp.P2
            assert tuple(int(x) for x in Fp.complex_shape()) == gp.complex_shape(r)
            assert tuple(Fp.complex_local_slice()) == tuple(gp.complex_local_slice(r))
            assert tuple(Fp.real_local_slice(1.5)) == tuple(gp.real_local_slice(r, 1.5))
            # sub-communicator collectives work (used to bootstrap the NCCL sub-communicators)
            assert Fp.comm0.bcast(("c0", Fp.comm1_rank) if Fp.comm0_rank == 0 else None, root=0) == ("c0", Fp.comm1_rank)
            assert sorted(Fp.comm1.allgather(r)) == sorted(gp.comm1_groups()[Fp.comm0_rank])
    else:
        try:
            m.Pencil_R2C(N, L, comm, "double")
            raise SystemExit("pencil on 2 ranks must raise IOError")
        except IOError:
            pass
    # the transforms themselves have no CPU path
    try:
        F.fftn(np.zeros(F.real_shape()), np.zeros(F.complex_shape(), dtype=complex))
        raise SystemExit("transform ran without a GPU")
    except m._lib.B200FFTError:
        pass
    comm.barrier()
    dist.destroy_process_group()
    print("WORKER_OK", r)


if __name__ == "__main__":
    main()
