"""A few small launches of every kernel family and both single-rank layouts, sized for compute-sanitizer
(memcheck / racecheck / synccheck run each kernel tens of times slower):

    compute-sanitizer --tool racecheck --error-exitcode 9 python tests/gpu_sanitize_worker.py

Results are still checked against numpy, so a clean sanitizer run is also a parity run."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import torch  # noqa: E402

import oracle  # noqa: E402
import test_passes as tp  # noqa: E402


def main():
    import mpifft4py_b200 as m
    from mpifft4py_b200.comm import SelfComm
    assert torch.cuda.is_available()
    be = tp._Gpu()
    L3 = np.array([2 * np.pi] * 3)
    for prec in "ds":
        for n in (64, 1024, 1536):
            tp.test_strided_c2c_all_plans(be, n, prec)
        for h in (32, 512, 768, 2048):  # R2C; C2R register-staged (512, 768) and staging (32, 2048) kernels
            tp.test_rows_r2c_c2r(be, h, prec)
    tp.test_pad_on_load_and_truncate_fold_on_store(be, 1024)
    tp.test_peer_chunk_store_and_gather_load(be, 8, 1024)
    tp.test_rows_uneven_kz_chunks(be)
    tp.test_rows_row_map(be, 512, "d")
    tp.test_rows_c2c(be, 1024, "d")
    for inverse in (0, 1):
        tp.test_four_step_long_axis(be, 16, 8, "d", inverse)
        tp.test_four_step_long_axis(be, 128, 128, "s", inverse)
    print("kernels ok", flush=True)
    # Navier-Stokes elementwise kernels: one RK4 step of the 16^3 Taylor-Green problem
    N16 = np.array([16, 16, 16])
    FN = m.Slab_R2C(N16, L3, SelfComm(), "double")
    S = m.ns.Solver(FN, nu=0.01, dt=0.01)
    X = [torch.from_numpy(np.ascontiguousarray(np.broadcast_to(x, FN.real_shape()))).cuda() for x in FN.get_local_mesh()]
    S.set_velocity(torch.stack([torch.sin(X[0]) * torch.cos(X[1]) * torch.cos(X[2]), -torch.cos(X[0]) * torch.sin(X[1]) * torch.cos(X[2]),
                                torch.zeros_like(X[0])]))
    S.step()
    assert np.isfinite(S.kinetic_energy())
    print("ns ok", flush=True)
    # whole transforms, y-blocked and natural intermediate
    N = (4, 512, 512)
    A = np.random.default_rng(0).random(N)
    ref = oracle.slab.fftn([A], N, 1)[0]
    for layout in ("yblock", "natural"):
        F = m.Slab_R2C(np.array(N), L3, SelfComm(), "double")
        F.layout = layout
        c = F.fftn(A, np.zeros(F.complex_shape(), dtype=np.complex128))
        assert oracle.rel_l2(c, ref) <= 1e-12, layout
        assert oracle.rel_l2(F.ifftn(c, np.zeros(F.real_shape())), A) <= 1e-12, layout
        print("layout", layout, "ok", flush=True)
    print("SANITIZE_WORKER_OK")


if __name__ == "__main__":
    main()
