"""The product's whole stack on the CPU (TEST INFRASTRUCTURE): the Python classes and serial functions, Transform._run
with its staging and -- for several ranks -- the transport handshake of Transform._ensure_plan, then the C-ABI layer
itself (mpifft4py_b200/csrc/b200fft.cu built for the host, tests/emu/host_shim.cpp) executing the plan programs with the
kernels' own phase bodies in the emulator.  Only the CUDA runtime is a stand-in (host memory, ranks as threads, CUDA IPC
handles as plain pointers, stream memory operations as polls).  tests/fake_device.py answers the C-ABI calls with the
oracle; this runs the engine's own arithmetic, so whatever passes here was computed by the code that runs on the GPU,
index map for index map.  Nothing here is reachable from the product."""
import collections
import ctypes as C
import weakref

import numpy as np
import torch

import host_shim_util
from mpifft4py_b200 import _engine, _lib, serialFFT


class ShimLibrary(object):
    """libb200fft_hostshim.so behind the names of libb200fft.so.  The stand-in runtime numbers the exec calls (an event
    recorded by an earlier call must not satisfy a wait of a later one), hence the wrappers."""

    def __init__(self):
        self._L = host_shim_util.load()
        self.calls = collections.Counter()  # entry point -> number of calls (tests assert that the engine really ran)

    def __getattr__(self, name):
        fn = getattr(self._L, name)
        if not name.startswith("b200fft_exec_") and name not in ("b200fft_plan_create", "b200fft_plan_p2p_connect"):
            return fn

        def counted(*args):
            self.calls[name] += 1
            if name in ("b200fft_exec_forward", "b200fft_exec_inverse"):
                self._L.shim_next_epoch()
            return fn(*args)
        return counted


def _host_tensor(a, dtype):
    """serialFFT._to_device without a device: always a fresh contiguous tensor (the device copy is one, too)."""
    if serialFFT._is_tensor(a):
        return a.contiguous().to(serialFFT._tdtype(dtype)).clone()
    return torch.from_numpy(np.array(a, dtype=dtype, order="C", copy=True))


def install(monkeypatch):
    """Route this process's transform objects and serial functions through the host build.  Returns a cleanup
    callable for the END of the test: plans created meanwhile are destroyed through the host build (an object
    finalised later would hand its handle to the real CUDA library)."""
    shim = ShimLibrary()
    made = []
    real_ensure = _engine.Transform._ensure_plan

    def ensure_plan(self):
        if self._plan is not None:
            return
        real_ensure(self)  # the real code: descriptor, transport choice, handle exchange over self.comm
        self.device = torch.device("cpu")
        made.append(weakref.ref(self))

    monkeypatch.setattr(_lib, "lib", lambda: shim)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    monkeypatch.setattr(_engine.Transform, "_ensure_plan", ensure_plan)
    monkeypatch.setattr(_engine.Transform, "_stream", lambda self: C.c_void_p(None))
    monkeypatch.setattr(serialFFT, "_to_device", _host_tensor)
    monkeypatch.setattr(serialFFT, "_stream", lambda: C.c_void_p(None))

    def cleanup():
        cleanup.calls = shim.calls
        for ref in made:
            F = ref()
            if F is not None and F._plan is not None:
                shim.b200fft_plan_destroy(F._plan)
                F._plan = None

    cleanup.calls = shim.calls
    return cleanup
