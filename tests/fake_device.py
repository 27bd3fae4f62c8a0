"""A stand-in for libb200fft.so + the CUDA device UNDER mpifft4py_b200._engine.Transform._run (TEST INFRASTRUCTURE).

The classes' bookkeeping is CPU code and the plan programs / kernels have their emulator, but the piece between them
-- Transform._run: argument checks, dtype and contiguity handling, the staging buffers of numpy callers, the calls
into the C ABI -- only ever ran on a GPU.  With this module it runs on the CPU: `torch` tensors on the host play the
device buffers, `b200fft_copy` is a memmove, and `b200fft_exec_forward / _inverse` hand the staged input of all ranks
to the ORACLE and write this rank's block of the answer into the staged output.  Nothing here is reachable from the
product."""
import ctypes as C
import weakref

import numpy as np
import torch

import oracle
from mpifft4py_b200 import _cdefs as D
from mpifft4py_b200 import _engine, _lib, line, pencil, slab

MODES = {D.DEALIAS_NONE: None, D.DEALIAS_3_2: "3/2-rule", D.DEALIAS_2_3: "2/3-rule"}


class Handle(C.c_void_p):
    """Plan handle of the fake: a NULL pointer that knows its transform object.  NULL because an object that outlives
    the test is finalised by the REAL Transform.__del__, which hands the handle to the real b200fft_plan_destroy --
    harmless for NULL."""
    owner = None


class FakeLibrary(object):
    """The entry points Transform._run uses.  Plans are the transform objects themselves."""

    def __init__(self):
        self.copies = 0
        self.execs = 0
        self.error = b""

    def b200fft_last_error(self):
        return self.error

    def b200fft_copy(self, dst, src, nbytes, stream):
        C.memmove(dst.value, src.value, int(nbytes))
        self.copies += 1
        return 0

    def b200fft_stream_sync(self, stream):
        return 0

    def _exec(self, inverse, plan, src, dst, mode):
        F = plan.owner()
        dealias = MODES[int(mode)]
        padded = dealias == "3/2-rule"
        c2c = isinstance(F, slab.C2C)
        real_shape = tuple(int(s) for s in (F.real_shape_padded() if padded else F.real_shape()))
        cshape = tuple(int(s) for s in F.complex_shape())
        rdt = F.complex if c2c else F.float
        (ishape, idt), (oshape, odt) = ((cshape, F.complex), (real_shape, rdt)) if inverse else ((real_shape, rdt), (cshape, F.complex))

        def view(ptr, shape, dt):
            n = int(np.prod(shape)) * np.dtype(dt).itemsize
            return np.frombuffer((C.c_char * n).from_address(ptr.value), dtype=dt).reshape(shape)

        mine = np.array(view(src, ishape, idt))
        comm = F.comm
        P = comm.Get_size()
        blocks = (getattr(comm, "allgather_world", None) or comm.allgather)(mine) if P > 1 else [mine]
        kw = dict(dealias=dealias, precision="double" if F.float is np.float64 else "single")
        N = tuple(int(n) for n in F.N)
        if c2c:
            fn = oracle.slab.c2c_ifftn if inverse else oracle.slab.c2c_fftn
        elif isinstance(F, slab.R2C):
            fn = oracle.slab.ifftn if inverse else oracle.slab.fftn
        elif isinstance(F, line.R2C):
            fn = oracle.line.ifft2 if inverse else oracle.line.fft2
            if not inverse:
                kw["exact"] = True
        else:
            fn = oracle.pencil.ifftn if inverse else oracle.pencil.fftn
            kw.update(alignment="X" if isinstance(F, pencil.R2CX) else "Y", P1=F.P1, communication=F.communication)
        view(dst, oshape, odt)[...] = fn(blocks, N, P, **kw)[comm.Get_rank() if P > 1 else 0]
        assert np.array_equal(view(src, ishape, idt), mine), "a transform must not modify its (staged) input"
        self.execs += 1
        return 0

    def b200fft_exec_forward(self, plan, src, dst, mode, stream):
        return self._exec(0, plan, src, dst, mode)

    def b200fft_exec_inverse(self, plan, src, dst, mode, stream):
        return self._exec(1, plan, src, dst, mode)


def install(monkeypatch):
    """Route every Transform object of this process through a FakeLibrary (returned)."""
    fake = FakeLibrary()

    def ensure_plan(self):
        if self._plan is None:
            self._plan = Handle(None)
            self._plan.owner = weakref.ref(self)
            self.device = torch.device("cpu")

    monkeypatch.setattr(_engine.Transform, "_ensure_plan", ensure_plan)
    monkeypatch.setattr(_engine.Transform, "_stream", lambda self: None)
    monkeypatch.setattr(_lib, "lib", lambda: fake)
    return fake
