#!/usr/bin/env python
"""Micro-benchmark of one strided C2C pass (b200fft_exec_strided) as a function of the row stride.

Same transform length, same bytes, only the distance between consecutive rows of the transformed
axis changes -- separates HBM/L2 effects from address-translation effects of the long-stride x pass.
Prints one line per case: GB/s of algorithmic bytes (read once + written once).

    python scripts/microbench_strided.py [n] [precision d|s]
"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpifft4py_b200 import _cdefs as D  # noqa: E402
from mpifft4py_b200 import _lib  # noqa: E402


def run(n, prec, B, J, inplace, reps=5):
    L = _lib.lib()
    cdt = torch.complex128 if prec == "d" else torch.complex64
    esz = 16 if prec == "d" else 8
    x = torch.randn(B * n * J, dtype=torch.float64 if prec == "d" else torch.float32, device="cuda")
    x = torch.complex(x, x).to(cdt)
    y = x if inplace else torch.empty_like(x)
    d = D.StridedDesc()
    d.precision = D.DOUBLE if prec == "d" else D.SINGLE
    d.n, d.B, d.J = n, B, J
    d.inverse, d.fold_mode, d.scale = 0, 0, 1.0
    d.inp = D.plain_side(x.data_ptr(), n * J, J, n)
    d.out = D.plain_side(y.data_ptr(), n * J, J, n)
    d.mask = D.no_mask()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(2):
        _lib.check(L.b200fft_exec_strided(C.byref(d), st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        _lib.check(L.b200fft_exec_strided(C.byref(d), st))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    by = 2.0 * B * n * J * esz
    return ms, by / ms / 1e6


def run_sides(n, prec, B, J, far_in, far_out, reps=5):
    """Same bytes as `run`, but the load and the store side get their own row stride: 'near' = J
    elements (batch-major array [B][n][J]), 'far' = B*J elements (array [n][B][J], the batch entry being
    a column block) -- tells how much of the long-stride penalty belongs to the loads and how much to
    the stores."""
    L = _lib.lib()
    cdt = torch.complex128 if prec == "d" else torch.complex64
    esz = 16 if prec == "d" else 8
    x = torch.randn(B * n * J, dtype=torch.float64 if prec == "d" else torch.float32, device="cuda")
    x = torch.complex(x, x).to(cdt)
    y = torch.empty_like(x)
    d = D.StridedDesc()
    d.precision = D.DOUBLE if prec == "d" else D.SINGLE
    d.n, d.B, d.J = n, B, J
    d.inverse, d.fold_mode, d.scale = 0, 0, 1.0
    d.inp = D.plain_side(x.data_ptr(), J, B * J, n) if far_in else D.plain_side(x.data_ptr(), n * J, J, n)
    d.out = D.plain_side(y.data_ptr(), J, B * J, n) if far_out else D.plain_side(y.data_ptr(), n * J, J, n)
    d.mask = D.no_mask()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(2):
        _lib.check(L.b200fft_exec_strided(C.byref(d), st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        _lib.check(L.b200fft_exec_strided(C.byref(d), st))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return ms, 2.0 * B * n * J * esz / ms / 1e6


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    prec = sys.argv[2] if len(sys.argv) > 2 else "d"
    esz = 16 if prec == "d" else 8
    total = (4 << 30) // esz // n  # B*J for 4 GiB
    print("n", n, prec)
    for J in (512, 513, 4096, 32768, 131072, 131072 + 8, 524288, total):
        B = max(1, total // J)
        for inplace in (True, False):
            ms, gbs = run(n, prec, B, J, inplace)
            print("  B=%-6d J=%-8d row stride %9.1f KB  %s  %7.3f ms  %7.1f GB/s" %
                  (B, J, J * esz / 1024.0, "in-place " if inplace else "out-of-pl", ms, gbs))
    J = 4096
    B = total // J
    for far_in in (False, True):
        for far_out in (False, True):
            ms, gbs = run_sides(n, prec, B, J, far_in, far_out)
            print("  loads %-4s (%8.1f KB)  stores %-4s (%8.1f KB)  %7.3f ms  %7.1f GB/s" %
                  ("far" if far_in else "near", (B * J if far_in else J) * esz / 1024.0,
                   "far" if far_out else "near", (B * J if far_out else J) * esz / 1024.0, ms, gbs))


if __name__ == "__main__":
    main()
