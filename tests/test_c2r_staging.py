"""The staging C2R kernel (csrc/fft_kernels.cuh: C2RK -- persistent, double-buffered asynchronous copies) for the
row lengths where the library now launches the register-staged kernel (C2RDK, half-lengths 256 ... 1536; the
default path of tests/test_passes.py): every row plan, zero-padded spectra (3/2-rule), per-peer kz chunks, the row
map.  CPU emulator; on the device C2RK serves the short and the very long rows (test_passes.py[gpu-*])."""
import pytest

import emu_util
import test_passes as tp


@pytest.fixture(scope="module")
def be_staging():
    lib = emu_util.load()
    old = lib.emu_set_c2r_staging(1)
    yield tp._Emu()
    lib.emu_set_c2r_staging(old)


@pytest.mark.parametrize("prec", ["d", "s"])
@pytest.mark.parametrize("h", [256, 384, 512, 768, 1024, 1536])
def test_c2r_staging_all_plans(be_staging, h, prec):
    tp.test_rows_r2c_c2r(be_staging, h, prec)


@pytest.mark.parametrize("N", [256, 1024])
def test_c2r_staging_zero_pad(be_staging, N):
    tp.test_rows_truncate_and_zero_pad(be_staging, N)


def test_c2r_staging_row_map(be_staging):
    tp.test_rows_row_map(be_staging, 512, "d")
