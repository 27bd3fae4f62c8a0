"""mpifft4py_b200.serialFFT -- the reference's serial function table (pyfftw_fft.py:26-244) -- on the CPU through the
product's own C-ABI passes in the host build (tests/cpu_engine.py): all twelve transforms and dct types 1-4 against
numpy / scipy, return conventions (``b`` filled and returned; a new array without ``b``), inputs left untouched.  The
device run of the same functions is tests/test_gpu_transforms.py::test_serial_functions_vs_numpy and tests/test_dct.py."""
import numpy as np
import pytest
from scipy.fftpack import dct as scipy_dct

import cpu_engine
import mpifft4py_b200 as m
import oracle


@pytest.fixture
def engine(monkeypatch):
    cleanup = cpu_engine.install(monkeypatch)
    yield cleanup.calls
    cleanup()


def _c(rng, shape, ct=np.complex128):
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(ct)


def test_all_twelve_transforms(engine):
    rng = np.random.default_rng(11)
    a = rng.standard_normal((16, 32, 64))
    c = _c(rng, (16, 32, 33))
    a0, c0 = a.copy(), c.copy()
    tol = 1e-13
    assert oracle.rel_l2(m.rfftn(a, axes=(0, 1, 2)), np.fft.rfftn(a)) <= tol
    assert oracle.rel_l2(m.irfftn(c, axes=(0, 1, 2)), np.fft.irfftn(c, s=(16, 32, 64), axes=(0, 1, 2))) <= tol
    assert oracle.rel_l2(m.rfft2(a, axes=(1, 2)), np.fft.rfft2(a, axes=(1, 2))) <= tol
    assert oracle.rel_l2(m.irfft2(c, axes=(1, 2)), np.fft.irfft2(c, s=(32, 64), axes=(1, 2))) <= tol
    assert oracle.rel_l2(m.rfft(a, axis=2), np.fft.rfft(a, axis=2)) <= tol
    assert oracle.rel_l2(m.irfft(c, axis=2), np.fft.irfft(c, n=64, axis=2)) <= tol
    z = _c(rng, (16, 48, 64))
    z0 = z.copy()
    for ax in (0, 1, 2):
        assert oracle.rel_l2(m.fft(z, axis=ax), np.fft.fft(z, axis=ax)) <= tol
        assert oracle.rel_l2(m.ifft(z, axis=ax), np.fft.ifft(z, axis=ax)) <= tol
    assert oracle.rel_l2(m.fft2(z, axes=(0, 1)), np.fft.fft2(z, axes=(0, 1))) <= tol
    assert oracle.rel_l2(m.ifft2(z, axes=(1, 2)), np.fft.ifft2(z, axes=(1, 2))) <= tol
    assert oracle.rel_l2(m.fftn(z, axes=(0, 1, 2)), np.fft.fftn(z)) <= tol
    assert oracle.rel_l2(m.ifftn(z, axes=(0, 1, 2)), np.fft.ifftn(z)) <= tol
    # inputs are never modified (C2R destroys its input in FFTW; upstream copies unless overwrite_input, pyfftw_fft.py:76)
    assert np.array_equal(a, a0) and np.array_equal(c, c0) and np.array_equal(z, z0)
    # b given: filled and returned, also when it is not contiguous or of the other precision
    b = np.zeros_like(z)
    assert m.fft(z, b, axis=1) is b and oracle.rel_l2(b, np.fft.fft(z, axis=1)) <= tol
    wide = np.zeros((16, 48, 70), dtype=complex)
    assert m.fftn(z, wide[:, :, 3:67], axes=(0, 1, 2)).base is wide and oracle.rel_l2(wide[:, :, 3:67], np.fft.fftn(z)) <= tol
    a32 = a.astype(np.float32)
    out = m.rfftn(a32, axes=(0, 1, 2))
    assert out.dtype == np.complex64 and oracle.rel_l2(out, np.fft.rfftn(a)) <= 1e-6
    with pytest.raises(NotImplementedError):
        m.fft(_c(rng, (10, 4)), axis=0)                      # length 10 has no radix plan
    with pytest.raises(AssertionError):
        m.rfft(a, axis=0)                                    # the real transform runs along the last axis
    assert engine["b200fft_exec_strided"] > 20 and engine["b200fft_exec_r2c"] >= 4 and engine["b200fft_exec_c2r"] >= 3


@pytest.mark.parametrize("type", [1, 2, 3, 4])
def test_dct(engine, type):
    rng = np.random.default_rng(5)
    cases = (((17, 6, 5), 0), ((4, 25, 7), 1), ((3, 5, 65), 2)) if type == 1 else (((16, 6, 5), 0), ((4, 24, 7), 1), ((3, 5, 64), 2))
    for shape, axis in cases:
        x = rng.standard_normal(shape)
        got = m.dct(x, np.zeros(shape), type=type, axis=axis)
        assert np.linalg.norm(got - scipy_dct(x, type=type, axis=axis)) <= 1e-13 * np.linalg.norm(got)
        z = x + 1j * rng.standard_normal(shape)
        gz = m.dct(z, np.zeros(shape, dtype=complex), type=type, axis=axis)
        ref = scipy_dct(z.real, type=type, axis=axis) + 1j * scipy_dct(z.imag, type=type, axis=axis)
        assert np.linalg.norm(gz - ref) <= 1e-13 * np.linalg.norm(ref)
    assert engine["b200fft_exec_strided"] >= 6
