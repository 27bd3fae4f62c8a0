"""The reference's own test procedures (tests/ref_procedures.py, after /root/reference/tests/test_FFT.py) on the
device with one rank: every slab / line / C2C fixture parameter, expected results from the engine's SERIAL transforms
(mpifft4py_b200.rfftn ...) and COMM_SELF objects exactly as upstream computes them, upstream's tolerances.  The
pencil parameters need four ranks: tests/gpu_dist_worker.py runs the whole list (shared-GPU and multi-GPU runs)."""
import numpy as np
import pytest

import ref_procedures as rp
from mpifft4py_b200.comm import COMM_SELF

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("param", rp.params(1)[0] + rp.params(1)[1])
def test_forward_backward_and_padded(param):
    rng = np.random.default_rng(7)
    F = rp.make(param, COMM_SELF)
    rp.forward_backward(F, rng)
    rp.padded(F, rng)


@pytest.mark.parametrize("param", rp.params(1)[2])
def test_c2c(param):
    rp.c2c(rp.make(param, COMM_SELF), np.random.default_rng(8))
