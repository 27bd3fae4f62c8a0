"""A few small launches of every kernel family -- default and opt-in variants -- sized for compute-sanitizer
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
    for variant in (0, 21, 23, 30, 31, 32, 33, 34, 35):
        be.L.b200fft_set_variant(variant)
        for prec in "ds":
            for n in (64, 1024, 1536):
                tp.test_strided_c2c_all_plans(be, n, prec)
            for h in (32, 512, 768):
                tp.test_rows_r2c_c2r(be, h, prec)
        tp.test_pad_on_load_and_truncate_fold_on_store(be, 1024)
        tp.test_peer_chunk_store_and_gather_load(be, 8, 1024)
        tp.test_rows_uneven_kz_chunks(be)
        tp.test_rows_c2c(be, 1024, "d")
        print("variant", variant, "ok", flush=True)
    be.L.b200fft_set_variant(0)
    # whole transforms, incl. the fused z+y kernel (l2_mode 3) and the grouped / two-stream schedules
    N = (4, 512, 512)
    A = np.random.default_rng(0).random(N)
    ref = oracle.slab.fftn([A], N, 1)[0]
    for planes, mode in ((0, 0), (2, 1), (2, 2), (1, 3), (3, 3)):
        F = m.Slab_R2C(np.array(N), L3, SelfComm(), "double")
        F.l2_planes, F.l2_mode = planes, mode
        c = F.fftn(A, np.zeros(F.complex_shape(), dtype=np.complex128))
        assert oracle.rel_l2(c, ref) <= 1e-12, (planes, mode)
        assert oracle.rel_l2(F.ifftn(c, np.zeros(F.real_shape())), A) <= 1e-12, (planes, mode)
        print("l2", planes, mode, "ok", flush=True)
    print("SANITIZE_WORKER_OK")


if __name__ == "__main__":
    main()
