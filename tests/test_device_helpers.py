"""Device-side mesh / wavenumber / mask / work-array helpers (mpifft4py_b200/device.py).
The values are the host methods' (bit-exact vs the reference: test_host_api.py); these tests pin the
residence change: same structure, same numbers, broadcast axes kept compact.  Host tensors here, CUDA
tensors in the gpu-marked test."""
import numpy as np
import pytest

import mpifft4py_b200 as m
from mpifft4py_b200 import device as dev
from mpifft4py_b200.comm import SelfComm

L3 = np.array([2 * np.pi] * 3)


def _objects():
    N = np.array([8, 16, 32])
    yield m.Slab_R2C(N, L3, SelfComm(), "double")
    yield m.Slab_R2C(N, L3, SelfComm(), "single")
    yield m.Line_R2C(N[:2], L3[:2], SelfComm(), "double")


def _same(t, a):
    a = np.asarray(a)
    assert tuple(t.shape) == a.shape
    assert np.array_equal(t.cpu().numpy(), a.astype(np.uint8) if a.dtype == np.bool_ else a)


@pytest.mark.parametrize("device", ["cpu", pytest.param("cuda", marks=pytest.mark.gpu)])
def test_mesh_wavenumbers_mask_match_the_host_methods(device):
    for F in _objects():
        X, Xd = F.get_local_mesh(), dev.local_mesh(F, device)
        if isinstance(X, (list, tuple)):
            assert len(X) == len(Xd)
            for t, a in zip(Xd, X):
                _same(t, a)
        else:
            _same(Xd, X)
        kws = [{}] if isinstance(F, m.Line_R2C) else [{}, {"scaled": True}, {"scaled": True, "broadcast": True}]
        for kw in kws:
            K, Kd = F.get_local_wavenumbermesh(**kw), dev.local_wavenumbermesh(F, device, **kw)
            if isinstance(K, (list, tuple)):
                for t, a in zip(Kd, K):
                    _same(t, a)
            else:
                _same(Kd, K)
        _same(dev.dealias_filter(F, device), F.get_dealias_filter())


def test_broadcast_axes_stay_compact():
    F = m.Slab_R2C(np.array([8, 16, 32]), L3, SelfComm(), "double")
    K = dev.local_wavenumbermesh(F, "cpu", scaled=True, broadcast=True)
    for k in K:
        assert tuple(k.shape) == tuple(F.complex_shape())
        assert 0 in k.stride()  # expanded view of a 1-D array, not a dense copy


@pytest.mark.parametrize("device", ["cpu", pytest.param("cuda", marks=pytest.mark.gpu)])
def test_device_work_arrays_keys_and_zero_on_fetch(device):
    import torch
    W = dev.work_arrays(device)
    a = W[((4, 5), np.complex128, 0)]
    assert a.dtype == torch.complex128 and tuple(a.shape) == (4, 5) and float(a.abs().sum()) == 0
    a += 1
    assert float(W[((4, 5), np.complex128, 0, False)].abs().sum()) == 20  # fillzero=False keeps the content
    assert float(W[((4, 5), np.complex128, 0)].abs().sum()) == 0          # default: zeroed on fetch
    b = W[(a, 1)]                                                           # keyed by an existing tensor
    assert b is not a and b.dtype == a.dtype and b.shape == a.shape
    assert W[(np.zeros((4, 5), dtype=np.complex128), 1)] is b               # ... or by a numpy array
    assert len(W) == 2
    with pytest.raises(TypeError):
        W[("bad", 0)]
    with pytest.raises(TypeError):
        W.values()
