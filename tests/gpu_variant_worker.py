"""GPU worker for kernel variants that are not the default (run in a subprocess with a timeout by
tests/test_zz_gpu_experimental.py, so that a fault cannot take the main test process down).

    python tests/gpu_variant_worker.py cluster    # 2-CTA cluster strided pass (variant 21 / 20)
    python tests/gpu_variant_worker.py rowbar     # row kernels: per-row named barriers (30), register-staged C2R (31)
    python tests/gpu_variant_worker.py l2         # L2-blocked z / y passes: grouped launches, two streams, fused kernel
    python tests/gpu_variant_worker.py kzblock    # kz-blocked intermediate array
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import torch  # noqa: E402

import cluster_checks as cc  # noqa: E402
import oracle  # noqa: E402
import test_passes as tp  # noqa: E402


def cluster():
    import mpifft4py_b200 as m
    from mpifft4py_b200.comm import SelfComm
    be = tp._Gpu()
    be.L.b200fft_set_variant(35)  # (not a cluster kernel: the strided pass with its first stage fed from HBM)
    cc.run_all(be)
    be.L.b200fft_set_variant(23)  # 64-byte tile rows
    cc.run_all(be)
    be.L.b200fft_set_variant(21)  # every strided pass whose length has a cluster plan, 128-byte rows
    cc.run_all(be)
    # whole transforms whose x pass is a cluster launch (1024 plain, 1536 with the 3/2-rule), against the oracle
    N = (1024, 16, 16)
    F = m.Slab_R2C(np.array(N), np.array([2 * np.pi] * 3), SelfComm(), "double")
    A = np.random.default_rng(3).random(N)
    c = F.fftn(A, np.zeros(F.complex_shape(), dtype=np.complex128))
    ref = oracle.slab.fftn([A], N, 1)[0]
    assert oracle.rel_l2(c, ref) <= 1e-12
    assert oracle.rel_l2(F.ifftn(c, np.zeros(F.real_shape())), A) <= 1e-12
    for d in ("2/3-rule", "3/2-rule"):
        shp = F.real_shape_padded() if d == "3/2-rule" else F.real_shape()
        got = F.ifftn(ref, np.zeros(shp), dealias=d)
        assert oracle.rel_l2(got, oracle.slab.ifftn([ref], N, 1, dealias=d)[0]) <= 1e-12, d
    up = np.random.default_rng(4).random(F.real_shape_padded())
    got = F.fftn(up, np.zeros(F.complex_shape(), dtype=np.complex128), dealias="3/2-rule")
    assert oracle.rel_l2(got, oracle.slab.fftn([up], N, 1, dealias="3/2-rule")[0]) <= 1e-12
    # variant 20: only launches whose rows are >= 1 MB apart take the cluster kernel
    be.L.b200fft_set_variant(20)
    n, pitch, J = 1024, 1 << 16, 24  # rows 1 MB apart (complex128), 24 live columns
    x = torch.zeros((n, pitch), dtype=torch.complex128, device="cuda")
    x[:, :J] = torch.from_numpy(tp._cplx(np.random.default_rng(5), (n, J), np.complex128)).cuda()
    y = torch.zeros_like(x)
    from mpifft4py_b200 import _cdefs as D
    out = tp.run_strided(be, None, n, None, B=1, J=J, prec=D.DOUBLE,
                         in_side=D.plain_side(x.data_ptr(), 0, pitch, n), out_side=D.plain_side(y.data_ptr(), 0, pitch, n))
    del out
    ref = np.fft.fft(x[:, :J].cpu().numpy(), axis=0)
    assert tp._rel(y[:, :J].cpu().numpy(), ref) < 4e-14
    assert float(y[:, J:].abs().max()) == 0.0  # nothing outside the live columns was written
    be.L.b200fft_set_variant(0)


def rowbar():
    """Row-kernel variants: 30 = the rows of a CTA synchronise on their own named barrier (many more rows
    than one wave of persistent CTAs, so that rows really run out of step); 31 = register-staged C2R."""
    be = tp._Gpu()
    for variant in (30, 31, 32, 33, 34):  # 32: radix-12 row kernels at four resident CTAs; 33 / 34: paired R2C / C2R
        _rows_checks(be, variant)
    be.L.b200fft_set_variant(0)


def _rows_checks(be, variant):
    import mpifft4py_b200 as m
    from mpifft4py_b200 import _cdefs as D
    from mpifft4py_b200.comm import SelfComm
    be.L.b200fft_set_variant(variant)
    for prec in "ds":
        for h in (256, 384, 512, 768, 1024, 1536):
            tp.test_rows_r2c_c2r(be, h, prec)
    tp.test_rows_truncate_and_zero_pad(be, 256)
    tp.test_rows_uneven_kz_chunks(be)
    for n, rows in ((1024, 40000), (1536, 30011)):
        x = torch.rand((rows, n), dtype=torch.float64, device="cuda")
        X = torch.zeros((rows, n // 2 + 1), dtype=torch.complex128, device="cuda")
        y = torch.zeros_like(x)
        tp.run_rows(be, "exec_r2c", _P(x), D.plain_side(X.data_ptr(), n // 2 + 1, 1, n // 2 + 1), n, rows,
                    n // 2 + 1, D.DOUBLE)
        ref = torch.fft.rfft(x, dim=1)
        assert float(torch.linalg.vector_norm(X - ref) / torch.linalg.vector_norm(ref)) < 1e-14
        tp.run_rows(be, "exec_c2r", _P(y), D.plain_side(X.data_ptr(), n // 2 + 1, 1, n // 2 + 1), n, rows,
                    n // 2 + 1, D.DOUBLE, scale=1.0 / n)
        assert float(torch.linalg.vector_norm(y - x) / torch.linalg.vector_norm(x)) < 1e-14
    N = (64, 64, 1024)
    F = m.Slab_R2C(np.array(N), np.array([2 * np.pi] * 3), SelfComm(), "double")
    A = np.random.default_rng(8).random(N)
    c = F.fftn(A, np.zeros(F.complex_shape(), dtype=np.complex128))
    assert oracle.rel_l2(c, oracle.slab.fftn([A], N, 1)[0]) <= 1e-12
    assert oracle.rel_l2(F.ifftn(c, np.zeros(F.real_shape())), A) <= 1e-12
    up = F.ifftn(c, np.zeros(F.real_shape_padded()), dealias="3/2-rule")
    assert oracle.rel_l2(up, oracle.slab.ifftn([c], N, 1, dealias="3/2-rule")[0]) <= 1e-12


def l2():
    """L2 blocking of the z and y passes (plan options l2_planes / l2_mode) against the oracle and against
    the unblocked plan; the fused persistent kernel (mode 3) is run repeatedly and must reproduce its own
    result bit for bit (a dependency race would show up as a varying result)."""
    import mpifft4py_b200 as m
    from mpifft4py_b200.comm import SelfComm
    L3 = np.array([2 * np.pi] * 3)
    for N, prec, planes in (((16, 512, 512), "double", 2), ((16, 512, 512), "single", 3), ((12, 1024, 1024), "double", 5),
                            ((4, 1024, 1024), "double", 1)):
        rt, ct = oracle.common.dtypes(prec)
        tol = 1e-12 if prec == "double" else 1e-5
        A = np.random.default_rng(21).random(N).astype(rt)
        ref = oracle.slab.fftn([A], N, 1, precision=prec)[0]
        for mode in (1, 2, 3):
            F = m.Slab_R2C(np.array(N), L3, SelfComm(), prec)
            F.l2_planes, F.l2_mode = planes, mode
            c = F.fftn(A, np.zeros(F.complex_shape(), dtype=ct))
            assert oracle.rel_l2(c, ref) <= tol, (N, prec, mode)
            if mode == 3:
                k, _ = F.last_launches()
                assert k == 2, "fused launch expected (z+y in one kernel, then x): %d kernels" % k
            assert oracle.rel_l2(F.ifftn(c, np.zeros(F.real_shape(), dtype=rt)), A) <= tol, (N, prec, mode)
            for d in ("2/3-rule", "3/2-rule"):
                shp = F.real_shape_padded() if d == "3/2-rule" else F.real_shape()
                got = F.ifftn(ref, np.zeros(shp, dtype=rt), dealias=d)
                assert oracle.rel_l2(got, oracle.slab.ifftn([ref], N, 1, dealias=d, precision=prec)[0]) <= tol, (N, prec, mode, d)
                if d == "3/2-rule":
                    back = F.fftn(got, np.zeros(F.complex_shape(), dtype=ct), dealias=d)
                    assert oracle.rel_l2(back, oracle.slab.fftn([got], N, 1, dealias=d, precision=prec)[0]) <= 10 * tol
            if mode == 3:
                tu = torch.from_numpy(A).cuda()
                tf = torch.zeros(tuple(int(s) for s in F.complex_shape()), dtype=torch.complex128 if prec == "double" else torch.complex64,
                                 device="cuda")
                F.fftn(tu, tf)
                first = tf.clone()
                for _ in range(20):
                    tf.zero_()
                    F.fftn(tu, tf)
                    assert torch.equal(tf, first)


def kzblock():
    """kz-blocked intermediate array (plan option kz_block), alone and with L2 grouping, against the oracle."""
    import mpifft4py_b200 as m
    from mpifft4py_b200.comm import SelfComm
    L3 = np.array([2 * np.pi] * 3)
    for N, prec in (((16, 512, 512), "double"), ((8, 1024, 1024), "double"), ((16, 512, 512), "single")):
        rt, ct = oracle.common.dtypes(prec)
        tol = 1e-12 if prec == "double" else 1e-5
        A = np.random.default_rng(22).random(N).astype(rt)
        ref = oracle.slab.fftn([A], N, 1, precision=prec)[0]
        for jc, planes, mode in ((48, 0, 0), (64, 0, 0), (48, 2, 1), (48, 3, 2)):
            F = m.Slab_R2C(np.array(N), L3, SelfComm(), prec)
            F.kz_block, F.l2_planes, F.l2_mode = jc, planes, mode
            c = F.fftn(A, np.zeros(F.complex_shape(), dtype=ct))
            assert oracle.rel_l2(c, ref) <= tol, (N, prec, jc, planes, mode)
            assert oracle.rel_l2(F.ifftn(c, np.zeros(F.real_shape(), dtype=rt)), A) <= tol, (N, prec, jc, planes, mode)
            got = F.ifftn(ref, np.zeros(F.real_shape(), dtype=rt), dealias="2/3-rule")
            assert oracle.rel_l2(got, oracle.slab.ifftn([ref], N, 1, dealias="2/3-rule", precision=prec)[0]) <= tol
        if N[1] == 1024:  # padded lengths 1536 have the blocked-column kernels too
            F = m.Slab_R2C(np.array(N), L3, SelfComm(), prec)
            F.kz_block = 48
            up = F.ifftn(ref, np.zeros(F.real_shape_padded(), dtype=rt), dealias="3/2-rule")
            assert oracle.rel_l2(up, oracle.slab.ifftn([ref], N, 1, dealias="3/2-rule", precision=prec)[0]) <= tol
            back = F.fftn(up, np.zeros(F.complex_shape(), dtype=ct), dealias="3/2-rule")
            assert oracle.rel_l2(back, oracle.slab.fftn([up], N, 1, dealias="3/2-rule", precision=prec)[0]) <= 10 * tol


class _P(object):
    """device tensor with the two attributes run_rows reads from a numpy array"""

    def __init__(self, t):
        self.t, self.shape = t, tuple(t.shape)
        self.ctypes = type("c", (), {"data": t.data_ptr()})()


if __name__ == "__main__":
    assert torch.cuda.is_available()
    {"cluster": cluster, "rowbar": rowbar, "l2": l2, "kzblock": kzblock}[sys.argv[1]]()
    print("VARIANT_WORKER_OK")
