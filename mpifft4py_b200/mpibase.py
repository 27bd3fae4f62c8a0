"""Work-array cache, datatype map and (pinned) host allocators.

Mirrors ``mpiFFT4py/mpibase.py:36-137``: ``work_arrays`` keeps its two key forms and its
zero-on-fetch behaviour, ``datatypes(precision)`` returns the (real, complex, wire) triple --
the wire type is a name since NCCL moves bytes, and ``empty``/``zeros`` hand out host arrays.
Where the reference aligns them for FFTW (pyfftw.empty_aligned, ``mpibase.py:38-45``) these are
page-locked when a CUDA device is present so host<->device staging runs at DMA speed.
"""
import collections.abc

import numpy as np

_pinned_keepalive = {}


def _pinned(shape, dtype):
    try:
        import torch
        if not torch.cuda.is_available():
            return None
        nbytes = int(np.prod(shape, dtype=np.int64)) * np.dtype(dtype).itemsize
        t = torch.empty(max(nbytes, 1), dtype=torch.uint8, pin_memory=True)
        a = t.numpy()[:nbytes].view(dtype).reshape(shape)
        _pinned_keepalive[a.ctypes.data] = t
        return a
    except Exception:  # noqa: BLE001 - pinned memory is an optimisation only
        return None


def empty(N, dtype=float, bytes=None):
    a = _pinned(N, dtype)
    return a if a is not None else np.empty(N, dtype=dtype)


def zeros(N, dtype=float, bytes=None):
    a = _pinned(N, dtype)
    if a is None:
        return np.zeros(N, dtype=dtype)
    a.fill(0)
    return a


class work_array_dict(dict):
    """Dictionary of work arrays indexed by their shape, type and an indicator i."""

    def __missing__(self, key):
        shape, dtype, i = key
        a = np.zeros(shape, dtype=dtype)
        self[key] = a
        return self[key]


class work_arrays(collections.abc.MutableMapping):
    """Host work arrays keyed ``(shape, dtype, index[, fillzero])`` or ``(ndarray, index[, fillzero])``
    (``mpibase.py:61-131``); fetched arrays are zeroed unless ``fillzero`` is False."""

    def __init__(self):
        self.store = work_array_dict()
        self.fillzero = True

    def __getitem__(self, key):
        val = self.store[self.__keytransform__(key)]
        if self.fillzero is True:
            val.fill(0)
        return val

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

    def __keytransform__(self, key):
        if isinstance(key[0], np.ndarray):
            shape = key[0].shape
            dtype = key[0].dtype
            i = key[1]
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
        return (tuple(int(s) for s in shape), np.dtype(dtype), i)


def datatypes(precision):
    """Return datatypes associated with precision (``mpibase.py:133-137``)."""
    assert precision in ("single", "double")
    return {"single": (np.float32, np.complex64, "C_FLOAT_COMPLEX"),
            "double": (np.float64, np.complex128, "C_DOUBLE_COMPLEX")}[precision]
