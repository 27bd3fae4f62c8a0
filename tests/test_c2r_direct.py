"""Register-staged C2R kernel (csrc/fft_kernels.cuh: C2RDK, kernel variant 31) in the CPU emulator: every
row plan, zero-padded spectra (3/2-rule), per-peer kz chunks.  The device run is part of
tests/gpu_variant_worker.py rowbar."""
import pytest

import emu_util
import test_passes as tp


@pytest.fixture(scope="module")
def be31():
    lib = emu_util.load()
    old = lib.emu_set_variant(31)
    yield tp._Emu()
    lib.emu_set_variant(old)


@pytest.mark.parametrize("prec", ["d", "s"])
@pytest.mark.parametrize("h", [2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 64, 96, 128, 256, 384, 512, 768, 1024, 1536, 2048])
def test_c2r_direct_all_plans(be31, h, prec):
    tp.test_rows_r2c_c2r(be31, h, prec)


@pytest.mark.parametrize("N", [8, 32, 256, 1024])
def test_c2r_direct_zero_pad(be31, N):
    tp.test_rows_truncate_and_zero_pad(be31, N)


def test_c2r_direct_uneven_kz_chunks(be31):
    tp.test_rows_uneven_kz_chunks(be31)


@pytest.fixture(scope="module")
def be33():
    """variant 33: R2C with the split step folded into a paired last stage (R2CPK)"""
    lib = emu_util.load()
    old = lib.emu_set_variant(33)
    yield tp._Emu()
    lib.emu_set_variant(old)


@pytest.mark.parametrize("prec", ["d", "s"])
@pytest.mark.parametrize("h", [2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 64, 96, 128, 256, 384, 512, 768, 1024, 1536, 2048, 8192])
def test_r2c_paired_all_plans(be33, h, prec):
    tp.test_rows_r2c_c2r(be33, h, prec)


@pytest.mark.parametrize("N", [8, 32, 256, 1024])
def test_r2c_paired_truncation(be33, N):
    tp.test_rows_truncate_and_zero_pad(be33, N)


def test_r2c_paired_uneven_kz_chunks(be33):
    tp.test_rows_uneven_kz_chunks(be33)


@pytest.fixture(scope="module")
def be34():
    """variant 34: C2R with the merge step folded into a paired first stage (C2RPK; plans with a first radix <= 8)"""
    lib = emu_util.load()
    old = lib.emu_set_variant(34)
    yield tp._Emu()
    lib.emu_set_variant(old)


@pytest.mark.parametrize("prec", ["d", "s"])
@pytest.mark.parametrize("h", [32, 48, 64, 96, 384, 512, 768, 6144, 8, 256, 1024])
def test_c2r_paired_all_plans(be34, h, prec):
    # 32 = (4,8), 48 = (4,12), 64 = (8,8), 96 = (8,12), 384 = (4,8,12), 512, 768, 6144 = (8,8,8,12): paired form;
    # 8 (one stage), 256 = (16,16), 1024 = (16,8,8): fall back to C2RK
    tp.test_rows_r2c_c2r(be34, h, prec)


@pytest.mark.parametrize("N", [32, 256, 1024])
def test_c2r_paired_zero_pad(be34, N):
    tp.test_rows_truncate_and_zero_pad(be34, N)


def test_c2r_paired_uneven_kz_chunks(be34):
    tp.test_rows_uneven_kz_chunks(be34)
