"""The 2-CTA cluster strided pass (csrc/fft_kernels.cuh: ClusterStridedK) in the CPU emulator: both CTAs
of a cluster are stepped through each phase with the partner's shared memory visible, which checks the
row split, the cross radix-2 stage, the half-length sub-transforms and every fused index map.  The
device run of the same checks is tests/test_zz_gpu_experimental.py."""
import pytest

import cluster_checks as cc
import emu_util
import test_passes as tp


@pytest.fixture(scope="module", params=[21, 23], ids=["rows128", "rows64"])
def be21(request):
    """variant 21: 128-byte tile rows (far strides); 23: 64-byte rows (long columns at near strides)"""
    lib = emu_util.load()
    old = lib.emu_set_variant(request.param)
    yield tp._Emu()
    lib.emu_set_variant(old)


@pytest.mark.parametrize("prec", ["d", "s"])
@pytest.mark.parametrize("n", cc.CLUSTER_LENGTHS)
def test_cluster_all_plans(be21, n, prec):
    cc.all_plans(be21, n, prec)
    cc.wide_and_batched(be21, n, prec)


@pytest.mark.parametrize("N", [1024, 2048])
def test_cluster_pad_truncate_fold(be21, N):
    cc.pad_truncate_fold(be21, N)


@pytest.mark.parametrize("P,n", [(8, 1024), (4, 1536)])
def test_cluster_peer_chunks(be21, P, n):
    cc.peer_chunks(be21, P, n)


def test_cluster_uneven_chunks_with_padding(be21):
    cc.uneven_chunks_with_padding(be21)


def test_cluster_mask_bands(be21):
    cc.mask_bands(be21)
