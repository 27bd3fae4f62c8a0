"""Build + load the CPU emulator of the CUDA kernels (tests/emu/emu.cpp).  TEST INFRASTRUCTURE."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "emu", "emu.cpp")
LIB = os.path.join(HERE, "emu", "libb200fft_emu.so")
DEPS = [SRC] + [os.path.join(ROOT, "mpifft4py_b200", "csrc", f) for f in
                ("fft_kernels.cuh", "fft_radix.cuh", "fft_plans.h", "desc_convert.h", "plan_program.h")] + \
       [os.path.join(ROOT, "include", "b200fft.h")]

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    stale = (not os.path.exists(LIB)) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in DEPS)
    if stale:
        subprocess.run(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", SRC, "-o", LIB], check=True)
    _lib = ctypes.CDLL(LIB)
    return _lib
