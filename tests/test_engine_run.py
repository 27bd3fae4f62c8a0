"""Transform._run -- the Python between the classes and the C ABI -- on the CPU, through tests/fake_device.py: argument
checks, dtype / contiguity handling of numpy callers, staging buffers, return conventions (slab.py:214,349: the method
fills and returns the caller's output array and leaves the input alone)."""
import threading

import numpy as np
import pytest

import fake_device
import mpifft4py_b200 as m
import oracle
from mpifft4py_b200.comm import COMM_SELF
from test_ref_procedures_oracle import ThreadComm, ThreadWorld

L3 = np.array([2 * np.pi] * 3)
N = (8, 16, 32)


@pytest.fixture
def fake(monkeypatch):
    return fake_device.install(monkeypatch)


def test_numpy_callers_are_staged_and_get_their_own_array_back(fake):
    F = m.Slab_R2C(np.array(N), L3, COMM_SELF, "double")
    rng = np.random.default_rng(0)
    u = rng.random(N)
    keep = u.copy()
    fu = np.zeros(F.complex_shape(), dtype=complex)
    out = F.fftn(u, fu)
    assert out is fu and np.array_equal(u, keep)
    ref = oracle.slab.fftn([u], N, 1)[0]
    assert oracle.rel_l2(fu, ref) < 1e-14
    u2 = F.ifftn(fu, np.zeros(N))
    assert oracle.rel_l2(u2, u) < 1e-14 and fake.execs == 2 and fake.copies == 4
    # one real and one complex staging buffer serve both directions
    assert sorted((k[0], k[1]) for k in F._stage) == [("a", N), ("b", tuple(int(s) for s in F.complex_shape()))]
    up = F.ifftn(fu, np.zeros(F.real_shape_padded()), dealias="3/2-rule")
    assert up.shape == (12, 24, 48) and len(F._stage) == 3
    assert oracle.rel_l2(up, oracle.slab.ifftn([fu], N, 1, dealias="3/2-rule")[0]) < 1e-14


def test_awkward_arrays(fake):
    """Fortran-ordered and strided inputs, inputs of another precision, outputs that are views or of another type: the
    reference's numpy backend takes them all (numpy_fft.py:25-107 converts); so does the staging path."""
    F = m.Slab_R2C(np.array(N), L3, COMM_SELF, "double")
    rng = np.random.default_rng(1)
    u = rng.random(N)
    ref = oracle.slab.fftn([u], N, 1)[0]
    big = np.zeros((8, 16, 40), dtype=complex)
    view = big[:, :, 3:20]                                  # not contiguous
    assert F.fftn(np.asfortranarray(u), view) is view and oracle.rel_l2(view, ref) < 1e-14 and not big[:, :, :3].any()
    wide = rng.random((8, 16, 64))
    assert oracle.rel_l2(F.fftn(wide[:, :, ::2], np.zeros_like(ref)), oracle.slab.fftn([wide[:, :, ::2]], N, 1)[0]) < 1e-14
    u32 = u.astype(np.float32)                               # single-precision data into a double-precision object
    assert oracle.rel_l2(F.fftn(u32, np.zeros_like(ref)), oracle.slab.fftn([u32.astype(float)], N, 1)[0]) < 1e-14
    c64 = np.zeros(ref.shape, dtype=np.complex64)            # ... and a single-precision output array
    assert F.fftn(u, c64) is c64 and oracle.rel_l2(c64, ref) < 1e-6
    ro = np.zeros(ref.shape, dtype=complex)
    ro.flags.writeable = False
    with pytest.raises(ValueError):
        F.fftn(u, ro)


def test_argument_errors(fake):
    F = m.Slab_R2C(np.array(N), L3, COMM_SELF, "double")
    u, fu = np.zeros(N), np.zeros(F.complex_shape(), dtype=complex)
    with pytest.raises(AssertionError):
        F.fftn(u, fu, dealias="4/3-rule")                    # slab.py:235
    with pytest.raises(AssertionError):
        F.fftn(np.zeros((8, 16, 30)), fu)
    with pytest.raises(AssertionError):
        F.fftn(u, fu, dealias="3/2-rule")                    # needs real_shape_padded(), slab.py:447
    with pytest.raises(AssertionError):
        F.ifftn(fu, np.zeros((4, 16, 32)))
    import torch
    with pytest.raises(AssertionError):
        F.fftn(torch.zeros(N, dtype=torch.float64), fu)      # numpy arrays or CUDA tensors, not a mix
    with pytest.raises(AssertionError):   # tensors of another device than the transform's
        F.fftn(torch.zeros(N, dtype=torch.float64, device="meta"), torch.zeros(tuple(fu.shape), dtype=torch.complex128, device="meta"))
    with pytest.raises(AssertionError):   # ... of the wrong precision
        F.fftn(torch.zeros(N, dtype=torch.float32), torch.zeros(tuple(fu.shape), dtype=torch.complex64))
    with pytest.raises(AssertionError):   # ... not contiguous
        F.fftn(torch.zeros((8, 16, 64), dtype=torch.float64)[:, :, ::2], torch.zeros(tuple(fu.shape), dtype=torch.complex128))
    assert fake.execs == 0


def test_device_tensors_are_used_in_place(fake):
    """Tensors on the transform's device (here: the host tensors the fake device works on) go to the C ABI by pointer:
    no staging buffer, no copy, the output tensor itself is filled and returned."""
    import torch
    F = m.Slab_R2C(np.array(N), L3, COMM_SELF, "double")
    u = torch.rand(N, dtype=torch.float64)
    keep = u.clone()
    fu = torch.zeros(tuple(int(s) for s in F.complex_shape()), dtype=torch.complex128)
    assert F.fftn(u, fu) is fu and fake.copies == 0 and not F._stage
    assert torch.equal(u, keep)
    assert oracle.rel_l2(fu.numpy(), oracle.slab.fftn([keep.numpy()], N, 1)[0]) < 1e-14
    back = F.ifftn(fu, torch.empty_like(u))
    assert oracle.rel_l2(back.numpy(), keep.numpy()) < 1e-14 and fake.copies == 0


@pytest.mark.parametrize("prec", ["double", "single"])
def test_line_and_c2c_objects(fake, prec):
    rt, ct = oracle.common.dtypes(prec)
    rng = np.random.default_rng(2)
    Fl = m.Line_R2C(np.array(N[:2]), L3[:2], COMM_SELF, prec)
    a = rng.random(N[:2]).astype(rt)
    c = Fl.fft2(a, np.zeros(Fl.complex_shape(), dtype=ct))
    assert c.dtype == ct and oracle.rel_l2(c, oracle.line.fft2([a], N[:2], 1, precision=prec)[0]) < 1e-6
    assert oracle.rel_l2(Fl.ifft2(c, np.zeros_like(a)), a) < 1e-5
    Fc = m.Slab_C2C(np.array(N), L3, COMM_SELF, prec)
    z = (rng.random(N) + 1j * rng.random(N)).astype(ct)
    zc = Fc.fftn(z, np.zeros(N, dtype=ct))
    assert oracle.rel_l2(zc, oracle.slab.c2c_fftn([z], N, 1, precision=prec)[0]) < 1e-6
    assert oracle.rel_l2(Fc.ifftn(zc, np.zeros_like(z)), z) < 1e-5
    assert len(Fc._stage) == 2  # same shape and type on both sides: still two distinct buffers, never in place


def test_four_thread_ranks_pencil(fake):
    P = 4
    tw = ThreadWorld(P)
    A = np.random.default_rng(3).random(N)
    errs = [None] * P

    def rank_main(r):
        try:
            comm = ThreadComm(tw, r)
            F = m.Pencil_R2C(np.array(N), L3, comm, "double", alignment="Y", communication="AlltoallN")
            g = oracle.pencil.Geometry(N, P, "Y", None, "AlltoallN")
            u = [np.ascontiguousarray(A[g.real_local_slice(q)]) for q in range(P)]
            c = F.fftn(u[r], np.zeros(F.complex_shape(), dtype=complex))
            errs[r] = oracle.rel_l2(c, oracle.pencil.fftn(u, N, P, alignment="Y", communication="AlltoallN")[r])
        except BaseException as e:  # noqa: BLE001
            tw.failed.append((r, repr(e)))
            tw.barrier.abort()

    ts = [threading.Thread(target=rank_main, args=(r,)) for r in range(P)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    assert not tw.failed, tw.failed
    assert max(errs) < 1e-14
