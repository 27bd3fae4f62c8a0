"""The fused FFT passes against numpy.fft, on two backends:

  emu : the CPU emulator (same phase bodies, g++-compiled, tests/emu) -- runs in `-m "not gpu"`;
  gpu : the real sm_100a kernels through the C ABI (b200fft_exec_strided / _r2c / _c2r) -- `-m gpu`.
        Test arrays live in page-locked host memory, which the device addresses directly (UVA), so
        the very same descriptors are used on both backends.

Covers every radix plan in both precisions plus each fused index map."""
import ctypes as C

import numpy as np
import pytest

import emu_util
from mpifft4py_b200 import _cdefs as D


class _Emu(object):
    name = "emu"

    def arr(self, a):
        return np.ascontiguousarray(a)

    def zeros(self, shape, dtype):
        return np.zeros(shape, dtype=dtype)

    def call(self, fn, desc):
        return getattr(emu_util.load(), "emu_" + fn)(C.byref(desc))


class _Gpu(object):
    name = "gpu"

    def __init__(self):
        from mpifft4py_b200 import _lib, mpibase
        import torch
        assert torch.cuda.is_available()
        torch.cuda.init()
        self.L = _lib.lib()
        self.check = _lib.check
        self.mpibase = mpibase

    def arr(self, a):
        out = self.mpibase.empty(a.shape, dtype=a.dtype)  # pinned: device-addressable
        out[...] = a
        return out

    def zeros(self, shape, dtype):
        return self.mpibase.zeros(shape, dtype=dtype)

    def call(self, fn, desc):
        rc = getattr(self.L, "b200fft_" + fn)(C.byref(desc), None)
        if rc == 0:
            self.check(self.L.b200fft_stream_sync(None))
        else:
            print(self.L.b200fft_last_error())
        return rc


@pytest.fixture(params=["emu", pytest.param("gpu", marks=pytest.mark.gpu)], scope="module")
def be(request):
    return _Emu() if request.param == "emu" else _Gpu()

LENS2 = [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192]
LENS3 = [3, 6, 12, 24, 48, 96, 192, 384, 768, 1536, 3072, 6144, 12288]


def _rel(x, ref):
    return float(np.linalg.norm((x - ref).ravel()) / max(np.linalg.norm(ref.ravel()), 1e-300))


def _ptr(a):
    return a.ctypes.data


def _cplx(rng, shape, ct):
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(ct)


def run_strided(be, x, n, out, inverse=0, scale=1.0, in_side=None, out_side=None, fold=0, mask=None,
                B=None, J=None, prec=None, cross=None):
    d = D.StridedDesc()
    if cross is not None:
        d.cross_n, d.cross_div = cross
    ref = x if x is not None else out
    if prec is None:
        prec = D.DOUBLE if (ref is None or ref.dtype == np.complex128) else D.SINGLE
    d.precision = prec
    d.n = n
    d.B = ref.shape[0] if B is None else B
    d.J = ref.shape[2] if J is None else J
    d.inverse = inverse
    d.fold_mode = fold
    d.scale = scale
    d.inp = in_side if in_side is not None else D.plain_side(_ptr(x), x.shape[1] * x.shape[2], x.shape[2], n)
    d.out = out_side if out_side is not None else D.plain_side(_ptr(out), out.shape[1] * out.shape[2],
                                                               out.shape[2], n)
    d.mask = mask if mask is not None else D.no_mask()
    rc = be.call("exec_strided", d)
    assert rc == 0, rc
    return out


@pytest.mark.parametrize("prec", ["d", "s"])
@pytest.mark.parametrize("n", LENS2 + LENS3)
def test_strided_c2c_all_plans(be, n, prec):
    ct = np.complex128 if prec == "d" else np.complex64
    tol = 2e-15 * max(1, np.log2(n)) if prec == "d" else 6e-7 * max(1, np.log2(n))
    rng = np.random.default_rng(n)
    J = 5 if n > 512 else 19  # not a multiple of the tile width: exercises dead lanes
    B = 1 if n > 512 else 2
    x = be.arr(_cplx(rng, (B, n, J), ct))
    out = be.zeros(x.shape, x.dtype)
    run_strided(be, x, n, out)
    assert _rel(out, np.fft.fft(x.astype(np.complex128), axis=1)) < tol
    if n <= 2048:
        out2 = be.zeros(x.shape, x.dtype)
        run_strided(be, out, n, out2, inverse=1, scale=1.0 / n)
        assert _rel(out2, x) < 2 * tol


def test_strided_c2c_16384_single(be):
    """line.R2C 16384^2 float32 (BASELINE.json config 5a): the x pass is one 16384-point column per
    CTA (128 KB of shared memory, no address table)."""
    n = 16384
    rng = np.random.default_rng(n)
    x = be.arr(_cplx(rng, (1, n, 3), np.complex64))
    out = be.zeros(x.shape, x.dtype)
    run_strided(be, x, n, out)
    assert _rel(out, np.fft.fft(x.astype(np.complex128), axis=1)) < 6e-7 * 14
    back = be.zeros(x.shape, x.dtype)
    run_strided(be, out, n, back, inverse=1, scale=1.0 / n)
    assert _rel(back, x) < 2e-6 * 14


@pytest.mark.parametrize("n", [8, 64, 1024, 12, 96, 1536])
def test_strided_in_place(be, n):
    rng = np.random.default_rng(1)
    x = be.arr(_cplx(rng, (3, n, 7), np.complex128))
    ref = np.fft.fft(x, axis=1)
    run_strided(be, x, n, x)
    assert _rel(x, ref) < 1e-14


@pytest.mark.parametrize("N", [8, 32, 64, 256, 1024])
def test_pad_on_load_and_truncate_fold_on_store(be, N):
    """copy_to_padded + ifft and fft + copy_from_padded (slab.py:516-533) in one pass each."""
    rng = np.random.default_rng(N)
    n = 3 * N // 2
    B, J = 2, 6
    fu = be.arr(_cplx(rng, (B, N, J), np.complex128))
    # inverse with zero padding and the reference's padsize scaling
    up = be.zeros((B, n, J), np.complex128)
    run_strided(be, fu, n, up, inverse=1, scale=1.5 / n,
                in_side=D.plain_side(_ptr(fu), N * J, J, N))
    fp = np.zeros((B, n, J), dtype=np.complex128)
    fp[:, :N // 2] = fu[:, :N // 2]
    fp[:, -(N // 2):] = fu[:, N // 2:]
    assert _rel(up, np.fft.ifft(fp * 1.5, axis=1)) < 1e-14
    # forward with truncation + Nyquist fold
    x = be.arr(_cplx(rng, (B, n, J), np.complex128))
    X = np.fft.fft(x, axis=1)
    ref = np.zeros((B, N, J), dtype=np.complex128)
    ref[:, :N // 2 + 1] = X[:, :N // 2 + 1]
    ref[:, N // 2:] += X[:, -(N // 2):]
    got = be.zeros(ref.shape, ref.dtype)
    run_strided(be, x, n, got, scale=1 / 1.5, fold=1, out_side=D.plain_side(_ptr(got), N * J, J, N))
    assert _rel(got, ref / 1.5) < 1e-14
    # line.py:189 variant (P == 1): mode -N/2 only, no fold
    ref2 = np.ascontiguousarray(X[:, np.r_[0:N // 2, n - N // 2:n]])
    got2 = be.zeros(ref2.shape, ref2.dtype)
    run_strided(be, x, n, got2, fold=2, out_side=D.plain_side(_ptr(got2), N * J, J, N))
    assert _rel(got2, ref2) < 1e-14


@pytest.mark.parametrize("P,n", [(2, 16), (4, 64), (8, 1024), (4, 96)])
def test_peer_chunk_store_and_gather_load(be, P, n):
    """Store into the per-peer send layout (U_mpi of slab.py:394-403 / subarraysB :206-209), gather
    from the receive layout on load (transpose_Uc, maths.pyx:21-31)."""
    rng = np.random.default_rng(P * n)
    B, J = 3, 5
    c = n // P
    x = be.arr(_cplx(rng, (B, n, J), np.complex128))
    X = np.fft.fft(x, axis=1)
    blocks = [be.zeros((B, c, J), np.complex128) for _ in range(P)]
    side = D.chunked_side([_ptr(b) for b in blocks], [c * J] * P, [J] * P, c, n)
    run_strided(be, x, n, None, out_side=side)
    for p in range(P):
        assert _rel(blocks[p], X[:, p * c:(p + 1) * c]) < 1e-14
    # gather: the same blocks as the receive side of the next pass
    y = be.zeros(x.shape, x.dtype)
    run_strided(be, None, n, y, inverse=1, scale=1.0 / n, in_side=side, B=B, J=J)
    assert _rel(y, x) < 1e-14


def test_uneven_last_chunk_with_padding(be):
    """Padded + chunked together: slab 3/2 inverse y pass gathers N1 = P*Np1 physical rows from P
    peers and pads to 1.5*N1 (slab.py:332-338)."""
    rng = np.random.default_rng(5)
    P, N, J, B = 4, 32, 3, 2
    n = 48
    c = N // P
    blocks = [be.arr(_cplx(rng, (B, c, J), np.complex128)) for _ in range(P)]
    full = np.concatenate(blocks, axis=1)
    fp = np.zeros((B, n, J), dtype=np.complex128)
    fp[:, :N // 2] = full[:, :N // 2]
    fp[:, -(N // 2):] = full[:, N // 2:]
    side = D.chunked_side([_ptr(b) for b in blocks], [c * J] * P, [J] * P, c, N)
    out = be.zeros((B, n, J), np.complex128)
    run_strided(be, None, n, out, inverse=1, scale=1.0 / n, in_side=side, B=B, J=J)
    assert _rel(out, np.fft.ifft(fp, axis=1)) < 1e-14


def test_mask_bands(be):
    """2/3-rule mask (slab.py:191-197) folded into the load of the first inverse pass."""
    rng = np.random.default_rng(9)
    N0, N1, Nf = 16, 8, 9  # pass over axis 0 of (N0, N1*Nf): i = kx, j = (ky, kz)
    fu = be.arr(_cplx(rng, (1, N0, N1 * Nf), np.complex128))
    kx = np.fft.fftfreq(N0, 1. / N0)
    ky = np.fft.fftfreq(N1, 1. / N1)
    kz = np.fft.rfftfreq(16, 1. / 16)
    kmax = 2. / 3. * (np.array([N0, N1, 16]) // 2 + 1)
    keep = ((abs(kx) < kmax[0])[:, None, None] * (abs(ky) < kmax[1])[None, :, None] *
            (abs(kz) < kmax[2])[None, None, :])
    ref = np.fft.ifft(fu.reshape(N0, N1, Nf) * keep, axis=0).reshape(1, N0, N1 * Nf)
    m = D.no_mask()
    m.on = 1
    lo = [int(np.ceil(k)) for k in kmax]
    m.i_off, m.i_lo, m.i_hi = 0, lo[0], N0 - lo[0]
    m.jdiv = Nf
    m.jq_off, m.jq_lo, m.jq_hi = 0, lo[1], N1 - lo[1]
    m.jr_off, m.jr_lo, m.jr_hi = 0, lo[2], 1 << 30
    out = be.zeros(fu.shape, fu.dtype)
    run_strided(be, fu, N0, out, inverse=1, scale=1.0 / N0, mask=m)
    assert _rel(out, ref) < 1e-14


def run_rows(be, fn, real, cplx_side, n, rows, nk, prec, scale=1.0, rowmap=None):
    d = D.RowsDesc()
    if rowmap is not None:
        d.rm_period, d.rm_block, d.rm_planes = rowmap
    d.precision = prec
    d.n = n
    d.rows = rows
    d.nk = nk
    d.scale = scale
    d.real_base = _ptr(real)
    d.rpitch = real.shape[1]
    d.cside = cplx_side
    rc = be.call(fn, d)
    assert rc == 0


@pytest.mark.parametrize("prec", ["d", "s"])
@pytest.mark.parametrize("h", [2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 64, 96, 128, 256, 384, 512, 768, 1024,
                               1536, 2048, 8192, 12288])
def test_rows_r2c_c2r(be, h, prec):
    n = 2 * h
    rt, ct = (np.float64, np.complex128) if prec == "d" else (np.float32, np.complex64)
    pr = D.DOUBLE if prec == "d" else D.SINGLE
    tol = 3e-15 * max(1, np.log2(n)) if prec == "d" else 8e-7 * max(1, np.log2(n))
    rng = np.random.default_rng(h)
    rows = 3 if h > 512 else 11
    x = be.arr(rng.standard_normal((rows, n)).astype(rt))
    X = be.zeros((rows, h + 1), ct)
    run_rows(be, "exec_r2c", x, D.plain_side(_ptr(X), h + 1, 1, h + 1), n, rows, h + 1, pr)
    ref = np.fft.rfft(x.astype(np.float64), axis=1)
    assert _rel(X, ref) < tol
    # C2R of an arbitrary (non-Hermitian-clean) spectrum: imag of DC / Nyquist must be ignored
    Y = be.arr(_cplx(rng, (rows, h + 1), ct))
    y = be.zeros((rows, n), rt)
    run_rows(be, "exec_c2r", y, D.plain_side(_ptr(Y), h + 1, 1, h + 1), n, rows, h + 1, pr, scale=1.0 / n)
    assert _rel(y, np.fft.irfft(Y.astype(np.complex128), n=n, axis=1)) < 2 * tol


@pytest.mark.parametrize("N", [8, 32, 256])
def test_rows_truncate_and_zero_pad(be, N):
    """3/2-rule z pass: R2C on 1.5N keeps Nf modes (slab.py:535); C2R pads Nf modes to 1.5N/2+1."""
    n = 3 * N // 2
    Nf = N // 2 + 1
    rng = np.random.default_rng(N)
    rows = 5
    x = be.arr(rng.standard_normal((rows, n)))
    X = be.zeros((rows, Nf), np.complex128)
    run_rows(be, "exec_r2c", x, D.plain_side(_ptr(X), Nf, 1, Nf), n, rows, Nf, D.DOUBLE)
    assert _rel(X, np.fft.rfft(x, axis=1)[:, :Nf]) < 1e-14
    Y = be.arr(_cplx(rng, (rows, Nf), np.complex128))
    y = be.zeros((rows, n), np.float64)
    run_rows(be, "exec_c2r", y, D.plain_side(_ptr(Y), Nf, 1, Nf), n, rows, Nf, D.DOUBLE, scale=1.0 / n)
    Yp = np.zeros((rows, n // 2 + 1), dtype=np.complex128)
    Yp[:, :Nf] = Y
    assert _rel(y, np.fft.irfft(Yp, n=n, axis=1)) < 1e-14


def test_rows_uneven_kz_chunks(be):
    """pencil z pass: kz split over P2 peers, the last one carrying the Nyquist plane
    (pencil.py:80-90 _distribution; subarrays2B :240-244)."""
    rng = np.random.default_rng(3)
    n, P2, rows = 64, 4, 6
    Nf = n // 2 + 1
    c = (n // 2) // P2
    lens = [c] * (P2 - 1) + [c + 1]
    x = be.arr(rng.standard_normal((rows, n)))
    blocks = [be.zeros((rows, l), np.complex128) for l in lens]
    side = D.chunked_side([_ptr(b) for b in blocks], lens, [1] * P2, c, Nf)
    run_rows(be, "exec_r2c", x, side, n, rows, Nf, D.DOUBLE)
    X = np.fft.rfft(x, axis=1)
    for p in range(P2):
        assert _rel(blocks[p], X[:, p * c:p * c + lens[p]]) < 1e-14
    y = be.zeros(x.shape, x.dtype)
    run_rows(be, "exec_c2r", y, side, n, rows, Nf, D.DOUBLE, scale=1.0 / n)
    assert _rel(y, x) < 1e-14


@pytest.mark.parametrize("prec", ["d", "s"])
@pytest.mark.parametrize("n", [4, 8, 64, 512, 1024, 2048, 6, 12, 96, 768, 1536])
def test_rows_c2c(be, n, prec):
    """Contiguous-row C2C (J == 1, unit stride): the z pass of slab.C2C; forward and inverse."""
    ct = np.complex128 if prec == "d" else np.complex64
    tol = 2e-15 * max(1, np.log2(n)) if prec == "d" else 6e-7 * max(1, np.log2(n))
    rng = np.random.default_rng(n)
    rows = 5 if n > 256 else 37
    x = be.arr(_cplx(rng, (rows, n, 1), ct))
    out = be.zeros(x.shape, x.dtype)
    run_strided(be, x, n, out)
    assert _rel(out, np.fft.fft(x.astype(np.complex128), axis=1)) < tol
    back = be.zeros(x.shape, x.dtype)
    run_strided(be, out, n, back, inverse=1, scale=1.0 / n)
    assert _rel(back, x) < 2 * tol


@pytest.mark.parametrize("N", [8, 64, 1024])
def test_rows_c2c_pad_truncate_fold(be, N):
    """3/2-rule along a contiguous axis (slab.C2C z pass, slab.py:803-825): zero pad on load for the
    inverse, truncation with fold (P > 1) or keeping mode -N/2 (P == 1, the `ks` gather) on store."""
    rng = np.random.default_rng(N)
    n = 3 * N // 2
    rows = 6
    fu = be.arr(_cplx(rng, (rows, N, 1), np.complex128))
    up = be.zeros((rows, n, 1), np.complex128)
    run_strided(be, fu, n, up, inverse=1, scale=1.5 / n, in_side=D.plain_side(_ptr(fu), N, 1, N))
    fp = np.zeros((rows, n, 1), dtype=np.complex128)
    fp[:, :N // 2] = fu[:, :N // 2]
    fp[:, -(N // 2):] = fu[:, N // 2:]
    assert _rel(up, np.fft.ifft(fp * 1.5, axis=1)) < 1e-14
    x = be.arr(_cplx(rng, (rows, n, 1), np.complex128))
    X = np.fft.fft(x, axis=1)
    ref = np.zeros((rows, N, 1), dtype=np.complex128)
    ref[:, :N // 2 + 1] = X[:, :N // 2 + 1]
    ref[:, N // 2:] += X[:, -(N // 2):]
    got = be.zeros(ref.shape, ref.dtype)
    run_strided(be, x, n, got, fold=1, out_side=D.plain_side(_ptr(got), N, 1, N))
    assert _rel(got, ref) < 1e-14
    ref2 = np.ascontiguousarray(X[:, np.r_[0:N // 2, n - N // 2:n]])
    got2 = be.zeros(ref2.shape, ref2.dtype)
    run_strided(be, x, n, got2, fold=2, out_side=D.plain_side(_ptr(got2), N, 1, N))
    assert _rel(got2, ref2) < 1e-14


@pytest.mark.parametrize("h,prec", [(8, "d"), (48, "d"), (128, "s"), (512, "d"), (768, "d"), (2048, "s")])
def test_rows_row_map(be, h, prec):
    """Row map of the row kernels (b200fft_rows_desc_t::rm_*): the spectrum of real row r = x * period + y lands at
    complex row ((y / block) * planes + x) * block + y % block -- the y-blocked intermediate [y block][x][y in block][k]
    of single-rank slab plans -- for the R2C store and both C2R kernels' loads (h = 2048: the staging kernel)."""
    n = 2 * h
    rt, ct = (np.float64, np.complex128) if prec == "d" else (np.float32, np.complex64)
    pr = D.DOUBLE if prec == "d" else D.SINGLE
    tol = 3e-15 * np.log2(n) if prec == "d" else 8e-7 * np.log2(n)
    rng = np.random.default_rng(h)
    planes, period, block = 3, 6, 2
    rows = planes * period
    x = be.arr(rng.standard_normal((rows, n)).astype(rt))
    X = be.zeros((period // block, planes, block, h + 1), ct)
    run_rows(be, "exec_r2c", x, D.plain_side(_ptr(X), h + 1, 1, h + 1), n, rows, h + 1, pr, rowmap=(period, block, planes))
    ref = np.fft.rfft(x.astype(np.float64), axis=1).reshape(planes, period // block, block, h + 1).transpose(1, 0, 2, 3)
    assert _rel(X, ref) < tol
    y = be.zeros((rows, n), rt)
    run_rows(be, "exec_c2r", y, D.plain_side(_ptr(X), h + 1, 1, h + 1), n, rows, h + 1, pr, scale=1.0 / n,
             rowmap=(period, block, planes))
    assert _rel(y, x) < 3 * tol


@pytest.mark.parametrize("n1,n2,prec", [(8, 8, "d"), (16, 8, "d"), (4, 32, "s"), (128, 128, "s"), (128, 64, "d")])
@pytest.mark.parametrize("inverse", [0, 1])
def test_four_step_long_axis(be, n1, n2, prec, inverse):
    """A transform of length N = n1 * n2 along the strided axis as TWO launches (columns too long for a tile of
    useful width): launch A runs n1-point transforms over rows n2 apart, the n2 interleaved sub-columns side by side
    as n2 * J columns, and multiplies by the cross twiddles W_N^(x2 * k1) on store; launch B runs n2-point transforms
    over the n2 consecutive rows of each k1 and scatters output k2 to row k1 + n1 * k2."""
    ct = np.complex128 if prec == "d" else np.complex64
    pr = D.DOUBLE if prec == "d" else D.SINGLE
    N = n1 * n2
    tol = 3e-15 * np.log2(N) if prec == "d" else 8e-7 * np.log2(N)
    rng = np.random.default_rng(N + inverse)
    J = 5
    x = be.arr(_cplx(rng, (1, N, J), ct))
    ref = (np.fft.ifft if inverse else np.fft.fft)(x.astype(np.complex128), axis=1)
    w = be.arr(x.copy())
    # A: in place; row i of the launch = rows [i * n2, (i + 1) * n2) of the array = n2 * J columns
    side = D.plain_side(_ptr(w), N * J, n2 * J, n1)
    run_strided(be, None, n1, None, inverse=inverse, in_side=side, out_side=side, B=1, J=n2 * J, prec=pr, cross=(N, J))
    # B: batch entry k1 = n2 consecutive rows; output k2 -> row k1 + n1 * k2
    out = be.zeros((1, N, J), ct)
    run_strided(be, None, n2, None, inverse=inverse, scale=(1.0 / N if inverse else 1.0),
                in_side=D.plain_side(_ptr(w), n2 * J, J, n2), out_side=D.plain_side(_ptr(out), J, n1 * J, n2), B=n1, J=J, prec=pr)
    assert _rel(out, ref) < tol
