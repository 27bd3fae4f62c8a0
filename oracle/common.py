"""Shared helpers of the CPU oracle (TEST INFRASTRUCTURE)."""
import numpy as np

try:  # scipy.fft is the same pocketfft with a ``workers`` knob; used only for timing legs
    import scipy.fft as _sfft
except Exception:  # noqa: BLE001
    _sfft = None

_WORKERS = 1


def set_workers(n):
    """Threads used by the 1D transforms (bench cpu_baseline: all host cores)."""
    global _WORKERS
    _WORKERS = int(n)


class _FFT(object):
    """numpy.fft semantics (``serialFFT/numpy_fft.py:25-51``); scipy.fft when workers>1."""

    @staticmethod
    def _mod():
        return _sfft if (_WORKERS != 1 and _sfft is not None) else None

    def fft(self, a, axis):
        m = self._mod()
        return m.fft(a, axis=axis, workers=_WORKERS) if m else np.fft.fft(a, axis=axis)

    def ifft(self, a, axis):
        m = self._mod()
        return m.ifft(a, axis=axis, workers=_WORKERS) if m else np.fft.ifft(a, axis=axis)

    def rfft(self, a, axis):
        m = self._mod()
        return m.rfft(a, axis=axis, workers=_WORKERS) if m else np.fft.rfft(a, axis=axis)

    def irfft(self, a, axis):
        m = self._mod()
        return m.irfft(a, axis=axis, workers=_WORKERS) if m else np.fft.irfft(a, axis=axis)


F = _FFT()


def dtypes(precision):
    """``mpibase.py:133-137`` without the MPI type."""
    assert precision in ("single", "double")
    return {"single": (np.float32, np.complex64), "double": (np.float64, np.complex128)}[precision]


def rel_l2(x, ref):
    x = np.asarray(x)
    ref = np.asarray(ref)
    d = np.linalg.norm((x.astype(np.complex128) - ref.astype(np.complex128)).ravel())
    n = np.linalg.norm(ref.astype(np.complex128).ravel())
    return float(d / n) if n > 0 else float(d)


def global_field(N, seed=1234, dtype=np.float64):
    """The synthetic input of SURVEY.md section 8d: uniform [0,1) like tests/test_FFT.py:63."""
    return np.random.default_rng(seed).random(tuple(int(n) for n in N)).astype(dtype)


def alltoall(groups, send):
    """All-to-all inside each group.

    groups : list of lists of world ranks; send[r][k] is what world rank r sends to the k-th
    member of its own group.  Returns recv with recv[r][k] = block sent to r by the k-th member.
    """
    nranks = sum(len(g) for g in groups)
    recv = [None] * nranks
    for g in groups:
        for ki, r in enumerate(g):
            recv[r] = [send[src][ki] for src in g]
    return recv


def pad_copy(fu, fp, N, axis):
    """Zero-pad copy along ``axis`` (``slab.py:516-525``, ``pencil.py:351-363``): the low half
    keeps its place, the high half (including the Nyquist row N/2) moves to the end."""
    h = int(N) // 2
    lo = [slice(None)] * fu.ndim
    hi_src = [slice(None)] * fu.ndim
    hi_dst = [slice(None)] * fu.ndim
    lo[axis] = slice(0, h)
    hi_src[axis] = slice(h, None)
    hi_dst[axis] = slice(fp.shape[axis] - h, None)
    fp[tuple(lo)] = fu[tuple(lo)]
    fp[tuple(hi_dst)] = fu[tuple(hi_src)]
    return fp


def trunc_fold(fp, fu, N, axis):
    """Truncate with Nyquist fold along ``axis`` (``slab.py:480-482,529-533``,
    ``pencil.py:369-379``): rows 0..N/2 copied, then rows -N/2.. ADDED onto N/2.."""
    h = int(N) // 2
    fu.fill(0)
    a = [slice(None)] * fu.ndim
    b = [slice(None)] * fu.ndim
    a[axis] = slice(0, h + 1)
    fu[tuple(a)] = fp[tuple(a)]
    a[axis] = slice(h, None)
    b[axis] = slice(fp.shape[axis] - h, None)
    fu[tuple(a)] += fp[tuple(b)]
    return fu


def dealias_mask(K, N):
    """2/3-rule mask (``slab.py:191-197``, ``pencil.py:343-349``, ``line.py:131-136``)."""
    kmax = 2. / 3. * (np.asarray(N) // 2 + 1)
    m = np.ones(np.broadcast(*K).shape, dtype=bool)
    for k, km in zip(K, kmax):
        m = m & (np.abs(k) < km)
    return m.astype(np.uint8)
