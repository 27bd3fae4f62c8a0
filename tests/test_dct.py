"""serialFFT.dct (types 1 to 4, scipy.fftpack convention; pyfftw_fft.py:205-244): the reordering / rotation
around the engine's complex FFT is checked on the CPU with torch.fft standing in for the kernel, the whole
function on the device."""
import numpy as np
import pytest
import torch
from scipy.fftpack import dct as scipy_dct

from mpifft4py_b200 import serialFFT


def _torch_fft(t, axis, inverse):
    return torch.fft.ifft(t, dim=axis) if inverse else torch.fft.fft(t, dim=axis)


@pytest.mark.parametrize("type", [1, 2, 3, 4])
@pytest.mark.parametrize("N,axis", [(8, 0), (16, 1), (12, 2), (48, 1), (7, 0), (2, 1), (33, 2)])
def test_dct_reordering_against_scipy(N, axis, type):
    rng = np.random.default_rng(N + type)
    shape = [3, 4, 5]
    shape[axis] = N
    for cplx in (False, True):
        x = rng.standard_normal(shape) + (1j * rng.standard_normal(shape) if cplx else 0)
        y = serialFFT._dct_core(torch.from_numpy(x.astype(np.complex128)), type, axis, _torch_fft).numpy()
        ref = scipy_dct(x.real, type=type, axis=axis) + 1j * scipy_dct(x.imag, type=type, axis=axis)
        assert np.abs(y - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())


def test_dct_type_errors():
    with pytest.raises(NotImplementedError):
        serialFFT._dct_core(torch.zeros(8, dtype=torch.complex128), 5, 0, _torch_fft)


@pytest.mark.gpu
@pytest.mark.parametrize("type", [1, 2, 3, 4])
@pytest.mark.parametrize("prec", ["double", "single"])
def test_dct_on_device(prec, type):
    rt, ct, tol = (np.float64, np.complex128, 1e-12) if prec == "double" else (np.float32, np.complex64, 2e-5)
    rng = np.random.default_rng(5)
    # (type 1 runs a complex FFT of length 2(N-1): the Gauss-Lobatto sizes 2^k + 1 and 3*2^k + 1)
    cases = (((65, 6, 5), 0), ((4, 97, 7), 1), ((3, 5, 1025), 2)) if type == 1 else (((64, 6, 5), 0), ((4, 96, 7), 1), ((3, 5, 1024), 2))
    for shape, axis in cases:
        x = rng.standard_normal(shape).astype(rt)
        ref = scipy_dct(x.astype(np.float64), type=type, axis=axis)
        got = serialFFT.dct(x, np.zeros(shape, dtype=rt), type=type, axis=axis)
        assert np.linalg.norm(got - ref) <= tol * np.linalg.norm(ref)
        z = (x + 1j * rng.standard_normal(shape)).astype(ct)
        refz = scipy_dct(z.real.astype(np.float64), type=type, axis=axis) + 1j * scipy_dct(z.imag.astype(np.float64), type=type, axis=axis)
        gotz = serialFFT.dct(z, np.zeros(shape, dtype=ct), type=type, axis=axis)
        assert np.linalg.norm(gotz - refz) <= tol * np.linalg.norm(refz)
        t = torch.from_numpy(z).cuda()
        out = torch.zeros_like(t)
        assert serialFFT.dct(t, out, type=type, axis=axis) is out
        assert np.linalg.norm(out.cpu().numpy() - refz) <= tol * np.linalg.norm(refz)
