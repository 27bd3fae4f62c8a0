"""Build + load the HOST build of the C-ABI layer (tests/emu/host_shim.cpp): the product's own
mpifft4py_b200/csrc/b200fft.cu compiled with g++ against an inert CUDA runtime stand-in, kernels bound to
the CPU emulator.  TEST INFRASTRUCTURE: checks the control flow of plan creation / program execution /
timing records without a GPU; never loaded by the product."""
import os
import subprocess

import ctypes

from mpifft4py_b200 import _lib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "emu", "host_shim.cpp")
LIB = os.path.join(HERE, "emu", "libb200fft_hostshim.so")
CS = os.path.join(ROOT, "mpifft4py_b200", "csrc")
DEPS = [SRC, os.path.join(HERE, "emu", "emu.cpp")] + \
       [os.path.join(HERE, "emu", "cuda_shim", f) for f in ("cuda_runtime.h", "cuda.h", "nccl.h")] + \
       [os.path.join(CS, f) for f in ("b200fft.cu", "fft_kernels.cuh", "fft_radix.cuh", "fft_plans.h", "fft_dispatch.h",
                                      "desc_convert.h", "plan_program.h", "ns_ops.cuh")] + [os.path.join(ROOT, "include", "b200fft.h")]

_shim = None


def load():
    global _shim
    if _shim is not None:
        return _shim
    stale = (not os.path.exists(LIB)) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in DEPS)
    if stale:
        subprocess.run(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-x", "c++", "-I", os.path.join(HERE, "emu", "cuda_shim"),
                        SRC, "-o", LIB, "-ldl", "-pthread"], check=True)
    L = ctypes.CDLL(LIB)
    _lib.declare(L)
    for name in _lib.SYMBOLS:
        getattr(L, name)
    _shim = L
    return L
