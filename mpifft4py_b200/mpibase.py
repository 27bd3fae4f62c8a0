"""Work-array cache, datatype map and (pinned) host allocators.

Mirrors ``mpiFFT4py/mpibase.py:36-137``: ``work_arrays`` keeps its two key forms and its
zero-on-fetch behaviour, ``datatypes(precision)`` returns the (real, complex, wire) triple --
the wire type is a name since the exchanges move bytes, and ``empty``/``zeros`` hand out host arrays.
Where the reference aligns them for FFTW (pyfftw.empty_aligned, ``mpibase.py:38-45``) these are
page-locked when a CUDA device is present so host<->device staging runs at DMA speed.
"""
import collections.abc

import numpy as np

def _pinned(shape, dtype):
    try:
        import torch
        if not torch.cuda.is_available():
            return None
        nbytes = int(np.prod(shape, dtype=np.int64)) * np.dtype(dtype).itemsize
        t = torch.empty(max(nbytes, 1), dtype=torch.uint8, pin_memory=True)
        # the array (through its .base chain) keeps the page-locked storage alive and releases it when it dies
        return t.numpy()[:nbytes].view(dtype).reshape(shape)
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


class work_arrays(collections.abc.MutableMapping):
    """Host work arrays (``mpibase.py:53-131``), created on first use and handed out ZEROED on every fetch.

    A key names an array by ``(shape, dtype, index)`` or by example, ``(ndarray, index)``; a trailing ``False``
    (``fillzero``) keeps the content of an existing array.  The index tells apart arrays of equal shape and type."""

    def __init__(self):
        self.store = {}
        self.fillzero = True

    def __keytransform__(self, key):
        like, rest = key[0], key[1:]
        if isinstance(like, np.ndarray):
            shape, dtype = like.shape, like.dtype
        elif isinstance(like, tuple) and len(rest) in (2, 3):
            shape, dtype, rest = like, rest[0], rest[1:]
        else:
            raise TypeError("Wrong type of key for work array")
        if len(rest) not in (1, 2):
            raise TypeError("Wrong type of key for work array")
        index, zero = rest[0], (rest[1] if len(rest) == 2 else True)
        assert isinstance(zero, bool)
        assert isinstance(index, int)
        self.fillzero = zero
        return (tuple(int(s) for s in shape), np.dtype(dtype), index)

    def __getitem__(self, key):
        k = self.__keytransform__(key)
        a = self.store.get(k)
        if a is None:
            a = self.store[k] = zeros(k[0], dtype=k[1])   # page-locked where a CUDA device is present
        elif self.fillzero is True:
            a.fill(0)
        return a

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


def datatypes(precision):
    """Return datatypes associated with precision (``mpibase.py:133-137``)."""
    assert precision in ("single", "double")
    return {"single": (np.float32, np.complex64, "C_FLOAT_COMPLEX"),
            "double": (np.float64, np.complex128, "C_DOUBLE_COMPLEX")}[precision]
