"""Strided pass with its first stage fed straight from memory (csrc/fft_kernels.cuh: StridedDK, kernel variant
35) in the CPU emulator: every plan with two or more stages and every fused index map of the strided pass."""
import pytest

import emu_util
import test_passes as tp


@pytest.fixture(scope="module")
def be35():
    lib = emu_util.load()
    old = lib.emu_set_variant(35)
    yield tp._Emu()
    lib.emu_set_variant(old)


@pytest.mark.parametrize("prec", ["d", "s"])
@pytest.mark.parametrize("n", [32, 64, 128, 256, 512, 1024, 2048, 4096, 24, 48, 96, 192, 384, 768, 1536, 3072])
def test_direct_all_plans(be35, n, prec):
    tp.test_strided_c2c_all_plans(be35, n, prec)


@pytest.mark.parametrize("n", [64, 1024, 96, 1536])
def test_direct_in_place(be35, n):
    tp.test_strided_in_place(be35, n)


@pytest.mark.parametrize("N", [32, 64, 256, 1024])
def test_direct_pad_truncate_fold(be35, N):
    tp.test_pad_on_load_and_truncate_fold_on_store(be35, N)


@pytest.mark.parametrize("P,n", [(4, 64), (8, 1024), (4, 96)])
def test_direct_peer_chunks(be35, P, n):
    tp.test_peer_chunk_store_and_gather_load(be35, P, n)


def test_direct_padding_with_chunks_and_masks(be35):
    tp.test_uneven_last_chunk_with_padding(be35)
    tp.test_mask_bands(be35)
