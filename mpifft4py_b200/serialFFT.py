"""Serial (single-GPU) transforms with the reference's duck-typed signature
``fn(a, b=None, axis=... | axes=..., overwrite_input=False, threads=1, planner_effort=...) -> b``
(``mpiFFT4py/serialFFT/pyfftw_fft.py:26-203``, ``numpy_fft.py:25-107``).

Each call is one or more fused passes of ``libb200fft.so``: ``b200fft_exec_strided`` for complex
axes, ``b200fft_exec_r2c`` / ``b200fft_exec_c2r`` for the real last axis.  numpy arguments are
staged through device memory; CUDA tensors are used in place.  Forward transforms are
unnormalised, inverses carry 1/n per axis (numpy convention).  Lengths must be 2^k or 3*2^k.
"""
import ctypes as C

import numpy as np

from . import _cdefs as D
from . import _lib

__all__ = ['dct', 'fft', 'ifft', 'fft2', 'ifft2', 'fftn', 'ifftn',
           'rfft', 'irfft', 'rfft2', 'irfft2', 'rfftn', 'irfftn']


def _torch():
    import torch
    return torch


_NP2T = None


def _tdtype(dt):
    global _NP2T
    torch = _torch()
    if _NP2T is None:
        _NP2T = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
                 np.dtype(np.complex64): torch.complex64, np.dtype(np.complex128): torch.complex128}
    return _NP2T[np.dtype(dt)]


def _is_tensor(x):
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


def _to_device(a, dtype):
    torch = _torch()
    if not torch.cuda.is_available():
        raise _lib.B200FFTError("no CUDA device: mpifft4py_b200 has no CPU path")
    if _is_tensor(a):
        assert a.is_cuda
        return a.contiguous().to(_tdtype(dtype))
    return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).cuda()


def _stream():
    return C.c_void_p(_torch().cuda.current_stream().cuda_stream)


def _c2c_axis(x, axis, inverse):
    """In-place complex FFT of device tensor ``x`` along ``axis``."""
    shape = tuple(x.shape)
    n = shape[axis]
    B = int(np.prod(shape[:axis], dtype=np.int64))
    J = int(np.prod(shape[axis + 1:], dtype=np.int64))
    d = D.StridedDesc()
    d.precision = D.DOUBLE if x.dtype == _tdtype(np.complex128) else D.SINGLE
    d.n, d.B, d.J = n, B, J
    d.inverse = int(inverse)
    d.fold_mode = 0
    d.scale = 1.0 / n if inverse else 1.0
    d.inp = D.plain_side(x.data_ptr(), n * J, J, n)
    d.out = D.plain_side(x.data_ptr(), n * J, J, n)
    d.mask = D.no_mask()
    _lib.check(_lib.lib().b200fft_exec_strided(C.byref(d), _stream()))
    return x


def _rows(real, cplx, n, forward):
    rows = int(np.prod(real.shape[:-1], dtype=np.int64))
    d = D.RowsDesc()
    d.precision = D.DOUBLE if real.dtype == _tdtype(np.float64) else D.SINGLE
    d.n, d.rows, d.nk = n, rows, n // 2 + 1
    d.scale = 1.0 if forward else 1.0 / n
    d.real_base = real.data_ptr()
    d.rpitch = n
    d.cside = D.plain_side(cplx.data_ptr(), n // 2 + 1, 1, n // 2 + 1)
    L = _lib.lib()
    _lib.check((L.b200fft_exec_r2c if forward else L.b200fft_exec_c2r)(C.byref(d), _stream()))


def _finish(res, a, b):
    """Return convention of the reference: fill and return ``b`` when given."""
    if b is None:
        return res if _is_tensor(a) else res.cpu().numpy()
    if _is_tensor(b):
        b.copy_(res)
    else:
        b[...] = res.cpu().numpy()
    return b


def _cdt(a):
    dt = a.dtype if not _is_tensor(a) else np.dtype(str(a.dtype).replace("torch.", ""))
    return np.complex64 if np.dtype(dt) in (np.dtype(np.float32), np.dtype(np.complex64)) else np.complex128


def _rdt(ct):
    return np.float32 if np.dtype(ct) == np.dtype(np.complex64) else np.float64


def _cn(a, axes, inverse):
    x = _to_device(a, _cdt(a))
    if not _is_tensor(a) or x.data_ptr() == a.data_ptr():
        x = x.clone() if _is_tensor(a) else x
    for ax in axes:
        _c2c_axis(x, ax % x.dim(), inverse)
    return x


def fft(a, b=None, axis=0, overwrite_input=False, threads=1, planner_effort=None, **kw):
    return _finish(_cn(a, (axis,), False), a, b)


def ifft(a, b=None, axis=0, overwrite_input=False, threads=1, planner_effort=None, **kw):
    return _finish(_cn(a, (axis,), True), a, b)


def fft2(a, b=None, axes=(0, 1), overwrite_input=False, threads=1, planner_effort=None, **kw):
    return _finish(_cn(a, tuple(axes)[::-1], False), a, b)


def ifft2(a, b=None, axes=(0, 1), overwrite_input=False, threads=1, planner_effort=None, **kw):
    return _finish(_cn(a, tuple(axes), True), a, b)


def fftn(a, b=None, axes=(0, 1, 2), overwrite_input=False, threads=1, planner_effort=None, **kw):
    return _finish(_cn(a, tuple(axes)[::-1], False), a, b)


def ifftn(a, b=None, axes=(0, 1, 2), overwrite_input=False, threads=1, planner_effort=None, **kw):
    return _finish(_cn(a, tuple(axes), True), a, b)


def _rn(a, axes):
    torch = _torch()
    ct = _cdt(a)
    x = _to_device(a, _rdt(ct))
    axes = [ax % x.dim() for ax in axes]
    assert axes[-1] == x.dim() - 1, "the real transform runs along the last axis (as in every reference call site)"
    n = x.shape[-1]
    out = torch.empty(tuple(x.shape[:-1]) + (n // 2 + 1,), dtype=_tdtype(ct), device=x.device)
    _rows(x, out, n, True)
    for ax in axes[-2::-1]:
        _c2c_axis(out, ax, False)
    return out


def _irn(a, axes):
    torch = _torch()
    ct = _cdt(a)
    x = _to_device(a, ct)
    if _is_tensor(a) and x.data_ptr() == a.data_ptr():
        x = x.clone()  # C2R and the in-place complex passes must not destroy the caller's input
    axes = [ax % x.dim() for ax in axes]
    assert axes[-1] == x.dim() - 1, "the real transform runs along the last axis (as in every reference call site)"
    for ax in axes[:-1]:
        _c2c_axis(x, ax, True)
    n = 2 * (x.shape[-1] - 1)
    out = torch.empty(tuple(x.shape[:-1]) + (n,), dtype=_tdtype(_rdt(ct)), device=x.device)
    _rows(out, x, n, False)
    return out


def rfft(a, b=None, axis=-1, overwrite_input=False, threads=1, planner_effort=None, **kw):
    return _finish(_rn(a, (axis,)), a, b)


def irfft(a, b=None, axis=-1, overwrite_input=False, threads=1, planner_effort=None, **kw):
    return _finish(_irn(a, (axis,)), a, b)


def rfft2(a, b=None, axes=(0, 1), overwrite_input=False, threads=1, planner_effort=None, **kw):
    return _finish(_rn(a, tuple(axes)), a, b)


def irfft2(a, b=None, axes=(0, 1), overwrite_input=False, threads=1, planner_effort=None, **kw):
    return _finish(_irn(a, tuple(axes)), a, b)


def rfftn(a, b=None, axes=(0, 1, 2), overwrite_input=False, threads=1, planner_effort=None, **kw):
    return _finish(_rn(a, tuple(axes)), a, b)


def irfftn(a, b=None, axes=(0, 1, 2), overwrite_input=False, threads=1, planner_effort=None, **kw):
    return _finish(_irn(a, tuple(axes)), a, b)


# ------------------------------------------------------------------------------------------------
# dct (serialFFT/pyfftw_fft.py:205-244, numpy_fft.py:11-22): types 1 to 4 with scipy.fftpack's unnormalised
# convention.  Types 2 and 3: ONE complex FFT of the engine along `axis` (Makhoul's reordering), real and
# imaginary parts of a complex input sharing it; type 1: the FFT of the even extension (length 2(N-1): the
# Chebyshev-Gauss-Lobatto sizes N = 2^k + 1 have kernels); type 4: a pre- and post-twiddled FFT of length 2N.
# Complex input is transformed part by part like upstream.
# Not on the R2C hot path (no caller in slab / pencil / line); provided because the function table of
# the reference's serialFFT module has it.
# ------------------------------------------------------------------------------------------------
def _dct_core(x, type, axis, fft_inplace):
    """``x``: complex tensor (device of the FFT); ``fft_inplace(t, axis, inverse)``: unnormalised forward /
    1/n-normalised inverse complex FFT in place.  Returns the complex result (imaginary part = DCT of the
    imaginary part of ``x``)."""
    torch = _torch()
    x = x.movedim(axis, -1)
    N = x.shape[-1]
    k = torch.arange(N, device=x.device, dtype=x.real.dtype)
    ang = torch.pi * k / (2 * N)
    if type == 2:
        v = torch.cat([x[..., 0::2], x[..., 1::2].flip(-1)], dim=-1).contiguous()
        Z = fft_inplace(v, v.dim() - 1, False)
        Zr = torch.roll(Z.flip(-1), 1, -1).conj()            # conj Z[(N - k) % N]
        e = torch.complex(torch.cos(ang), -torch.sin(ang))   # exp(-i pi k / 2N)
        yr = 2 * ((Z + Zr) * 0.5 * e).real                   # FFT of the real part of v, rotated
        yi = 2 * ((Z - Zr) * (-0.5j) * e).real               # ... of the imaginary part
        y = torch.complex(yr, yi)
    elif type == 3:
        xr = torch.cat([torch.zeros_like(x[..., :1]), x[..., 1:].flip(-1)], dim=-1)   # x[N - k], x[N] := 0
        e = torch.complex(torch.cos(ang), torch.sin(ang))    # exp(+i pi k / 2N)
        W = ((x - 1j * xr) * e).contiguous()
        w = fft_inplace(W, W.dim() - 1, True) * N            # unnormalised inverse
        y = torch.empty_like(w)
        y[..., 0::2] = w[..., :(N + 1) // 2]
        y[..., 1::2] = w.flip(-1)[..., :N // 2]
    elif type == 1:
        # even extension [x0 .. x_{N-1}, x_{N-2} .. x1] of length 2(N-1): its FFT is real for real input, so the
        # real and the imaginary part of a complex x come out of ONE transform as the real and imaginary part
        assert N >= 2, "dct type 1 needs at least two points"
        v = torch.cat([x, x[..., 1:N - 1].flip(-1)], dim=-1).contiguous()
        y = fft_inplace(v, v.dim() - 1, False)[..., :N]
    elif type == 4:
        # y[k] = 2 Re( exp(-i pi (2k+1) / 4N) * FFT_2N(x[n] exp(-i pi n / 2N), zero-padded)[k] ); the pre-twiddled
        # sequence has no symmetry to separate two real inputs with, so a complex x takes two transforms
        pre = torch.complex(torch.cos(ang), -torch.sin(ang))                 # exp(-i pi n / 2N)
        post = torch.complex(torch.cos(ang + torch.pi / (4 * N)), -torch.sin(ang + torch.pi / (4 * N)))

        def part(r):
            w = torch.cat([r * pre, torch.zeros_like(x)], dim=-1).contiguous()
            return 2 * (fft_inplace(w, w.dim() - 1, False)[..., :N] * post).real

        yr = part(x.real)
        y = torch.complex(yr, part(x.imag) if bool((x.imag != 0).any()) else torch.zeros_like(yr))
    else:
        raise NotImplementedError("dct type %r: scipy.fftpack has types 1 to 4" % (type,))
    return y.movedim(-1, axis)


def dct(a, b, type=2, axis=0, overwrite_input=False, threads=1, planner_effort=None, **kw):
    ct = _cdt(a)
    is_complex = (a.is_complex() if _is_tensor(a) else np.iscomplexobj(a))
    x = _to_device(a, ct)
    if _is_tensor(a) and x.data_ptr() == a.data_ptr():
        x = x.clone()
    y = _dct_core(x, type, axis % x.dim(), _c2c_axis)
    return _finish(y if is_complex else y.real.contiguous(), a, b)
