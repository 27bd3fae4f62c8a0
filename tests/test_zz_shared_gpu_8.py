"""Eight rank processes on GPU 0 (see test_gpu_multi.py::test_multi_rank_shared_gpu): the pencil grids only eight ranks
have -- 4x2 (default) and 2x4 (P1=2), both alignments, all three communication layouts -- and the goldens of the
unmodified reference with P = 8, against the oracle on the real kernels.  Runs last (file name): eight CUDA contexts
taking turns on one device make it the slowest GPU test."""
import pytest

from test_gpu_multi import shared_gpu_run

pytestmark = pytest.mark.gpu


def test_eight_ranks_shared_gpu():
    shared_gpu_run(8, timeout=420)
