"""Device-resident companions of the host helpers on either side of the transforms in a solver loop
(SURVEY.md section 8f-3): the mesh, the wavenumber mesh, the 2/3-rule mask and the work-array cache,
as ``torch`` tensors on the GPU that owns the plan.

The arithmetic is the classes' own host code (``slab.py:146-197``, ``pencil.py:289-349,945-969``,
``line.py:105-136`` upstream; bit-exact against the reference, tests/test_host_api.py) evaluated
once; only the residence changes, so that a caller such as ``examples/spectral_dns_solver.py``
never touches host memory between transforms.  ``device`` defaults to the current CUDA device;
passing ``"cpu"`` gives the same tensors on the host (used by the CPU tests of this module).
"""
import collections.abc

import numpy as np


def _torch():
    import torch
    return torch


def _device(device):
    torch = _torch()
    if device is not None:
        return torch.device(device)
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device: pass device='cpu' for host tensors")
    return torch.device("cuda", torch.cuda.current_device())


def to_device(a, device=None):
    """numpy array (or broadcast view) -> tensor on ``device``; broadcast axes stay broadcast (stride 0)
    so that a sparse mesh costs O(N) device memory, not O(N^3)."""
    torch = _torch()
    a = np.asarray(a)
    # find the axes along which the view is a broadcast (stride 0) and ship the compact array only
    idx = tuple(slice(0, 1) if (s == 0 and n > 1) else slice(None) for s, n in zip(a.strides, a.shape))
    compact = np.array(a[idx], copy=True, order="C")
    if compact.dtype == np.bool_:
        compact = compact.astype(np.uint8)
    t = torch.from_numpy(compact).to(_device(device))
    return t.expand(*a.shape) if compact.shape != a.shape else t


def local_mesh(FFT, device=None):
    """``FFT.get_local_mesh()`` on the device: a list of broadcastable tensors (slab, pencil 'Y') or one
    dense tensor (pencil 'X', line), exactly as the host method returns them."""
    X = FFT.get_local_mesh()
    if isinstance(X, (list, tuple)):
        return [to_device(x, device) for x in X]
    return to_device(X, device)


def local_wavenumbermesh(FFT, device=None, **kw):
    """``FFT.get_local_wavenumbermesh(**kw)`` on the device (same list-or-array structure)."""
    K = FFT.get_local_wavenumbermesh(**kw)
    if isinstance(K, (list, tuple)):
        return [to_device(k, device) for k in K]
    return to_device(K, device)


def dealias_filter(FFT, device=None):
    """The 2/3-rule mask of ``get_dealias_filter`` (uint8, ``complex_shape``) on the device.  The
    transforms do not need it -- ``ifftn(dealias='2/3-rule')`` folds the mask bands into the first
    pass's load -- but solvers that filter spectra themselves do."""
    return to_device(np.asarray(FFT.get_dealias_filter()), device)


class work_arrays(collections.abc.MutableMapping):
    """Device counterpart of ``mpibase.work_arrays`` (``mpibase.py:61-131``): tensors keyed
    ``(shape, dtype, index[, fillzero])`` or ``(tensor_or_array, index[, fillzero])``, created on first
    use, zeroed on every fetch unless ``fillzero`` is False."""

    def __init__(self, device=None):
        self.device = device
        self.store = {}
        self.fillzero = True

    @staticmethod
    def _tdtype(dtype):
        torch = _torch()
        if isinstance(dtype, torch.dtype):
            return dtype
        return {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
                np.dtype(np.complex64): torch.complex64, np.dtype(np.complex128): torch.complex128,
                np.dtype(np.uint8): torch.uint8, np.dtype(np.int32): torch.int32,
                np.dtype(np.int64): torch.int64}[np.dtype(dtype)]

    def __keytransform__(self, key):
        if hasattr(key[0], "shape") and hasattr(key[0], "dtype"):
            shape, dtype, i = tuple(key[0].shape), key[0].dtype, key[1]
            zero = True if len(key) == 2 else key[2]
        elif isinstance(key[0], tuple):
            if len(key) == 3:
                shape, dtype, i = key
                zero = True
            elif len(key) == 4:
                shape, dtype, i, zero = key
            else:
                raise TypeError("Wrong type of key for work array")
        else:
            raise TypeError("Wrong type of key for work array")
        assert isinstance(zero, bool)
        assert isinstance(i, int)
        self.fillzero = zero
        return (tuple(int(s) for s in shape), self._tdtype(dtype), i)

    def __getitem__(self, key):
        k = self.__keytransform__(key)
        t = self.store.get(k)
        if t is None:
            t = self.store[k] = _torch().zeros(k[0], dtype=k[1], device=_device(self.device))
        elif self.fillzero is True:
            t.zero_()
        return t

    def __setitem__(self, key, value):
        self.store[self.__keytransform__(key)] = value

    def __delitem__(self, key):
        del self.store[self.__keytransform__(key)]

    def __iter__(self):
        return iter(self.store)

    def __len__(self):
        return len(self.store)

    def values(self):
        raise TypeError('Work arrays not iterable')
