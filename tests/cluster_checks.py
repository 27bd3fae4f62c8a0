"""Checks of the 2-CTA cluster strided pass (ClusterStridedK, kernel variant 21) shared by the emulator
tests (tests/test_cluster_pass.py) and the GPU worker (tests/gpu_variant_worker.py).  `be` is a backend
of tests/test_passes.py whose variant has been set to 21 by the caller."""
import numpy as np

import test_passes as tp
from mpifft4py_b200 import _cdefs as D

CLUSTER_LENGTHS = [1024, 1536, 2048, 3072]


def all_plans(be, n, prec):
    tp.test_strided_c2c_all_plans(be, n, prec)


def wide_and_batched(be, n, prec):
    """several column tiles (J not a multiple of the 128-byte tile), several batch entries, in place"""
    ct = np.complex128 if prec == "d" else np.complex64
    tol = 2e-15 * np.log2(n) if prec == "d" else 6e-7 * np.log2(n)
    rng = np.random.default_rng(n + 1)
    x = be.arr(tp._cplx(rng, (2, n, 37), ct))
    ref = np.fft.fft(x.astype(np.complex128), axis=1)
    tp.run_strided(be, x, n, x)
    assert tp._rel(x, ref) < tol


def pad_truncate_fold(be, N):
    tp.test_pad_on_load_and_truncate_fold_on_store(be, N)


def peer_chunks(be, P, n):
    tp.test_peer_chunk_store_and_gather_load(be, P, n)


def uneven_chunks_with_padding(be):
    """slab 3/2 inverse x pass: N0 = 1024 physical rows gathered from 4 peers, padded to 1536"""
    rng = np.random.default_rng(6)
    P, N, J, B, n = 4, 1024, 3, 2, 1536
    c = N // P
    blocks = [be.arr(tp._cplx(rng, (B, c, J), np.complex128)) for _ in range(P)]
    full = np.concatenate(blocks, axis=1)
    fp = np.zeros((B, n, J), dtype=np.complex128)
    fp[:, :N // 2] = full[:, :N // 2]
    fp[:, -(N // 2):] = full[:, N // 2:]
    side = D.chunked_side([tp._ptr(b) for b in blocks], [c * J] * P, [J] * P, c, N)
    out = be.zeros((B, n, J), np.complex128)
    tp.run_strided(be, None, n, out, inverse=1, scale=1.0 / n, in_side=side, B=B, J=J)
    assert tp._rel(out, np.fft.ifft(fp, axis=1)) < 1e-14


def mask_bands(be):
    """2/3-rule mask on the load of a 1024-point inverse x pass: kx band by rows, ky / kz bands by columns"""
    rng = np.random.default_rng(10)
    N0, N1, Nf = 1024, 4, 5
    fu = be.arr(tp._cplx(rng, (1, N0, N1 * Nf), np.complex128))
    kx = np.fft.fftfreq(N0, 1. / N0)
    ky = np.fft.fftfreq(N1, 1. / N1)
    kz = np.fft.rfftfreq(8, 1. / 8)
    kmax = 2. / 3. * (np.array([N0, N1, 8]) // 2 + 1)
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
    tp.run_strided(be, fu, N0, out, inverse=1, scale=1.0 / N0, mask=m)
    assert tp._rel(out, ref) < 1e-14


def run_all(be):
    for n in CLUSTER_LENGTHS:
        for prec in "ds":
            all_plans(be, n, prec)
            wide_and_batched(be, n, prec)
    for N in (1024, 2048):
        pad_truncate_fold(be, N)
    peer_chunks(be, 8, 1024)
    peer_chunks(be, 4, 1536)
    uneven_chunks_with_padding(be)
    mask_bands(be)
