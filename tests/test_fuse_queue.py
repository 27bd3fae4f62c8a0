"""Work queue of the fused pair kernel (csrc/fft_kernels.cuh: fuse_decode, FuseSide): for many shapes --
row blocks that straddle plane groups, short last groups, one plane per group, either pass order -- every
block of both passes is visited once and no block of the second pass is queued before a block it waits for
(which is what makes the persistent kernel deadlock-free)."""
import numpy as np
import pytest

import emu_util


@pytest.mark.parametrize("rows_first", [1, 0])
def test_headline_shapes(rows_first):
    lib = emu_util.load()
    # (rows per plane, rows per block, planes, tiles per plane): 1024^3 and 1536^3-padded double, 512^3, 256^3 single
    for rpp, rpc, planes, tiles in ((1024, 3, 1024, 129), (1536, 2, 1536, 129), (512, 6, 512, 65), (256, 16, 256, 17)):
        for ppg in (1, 2, 3, 4, 6, 7, 8, 16, 100, planes):
            assert lib.emu_check_fuse_queue(rpp, rpc, planes, ppg, tiles, rows_first) == 0, (rpp, rpc, planes, ppg, tiles)


def test_random_shapes():
    lib = emu_util.load()
    rng = np.random.default_rng(7)
    for _ in range(3000):
        rpp = int(rng.integers(1, 40))
        rpc = int(rng.integers(1, 9))
        planes = int(rng.integers(1, 30))
        ppg = int(rng.integers(1, planes + 1))
        tiles = int(rng.integers(1, 6))
        if ppg * rpp < rpc:
            continue
        for rows_first in (0, 1):
            rc = lib.emu_check_fuse_queue(rpp, rpc, planes, ppg, tiles, rows_first)
            assert rc == 0, (rc, rpp, rpc, planes, ppg, tiles, rows_first)
