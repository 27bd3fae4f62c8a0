"""Known-answer integration test of the reference: Taylor-Green vortex, 32^3, nu = 0.000625,
dt = 0.01, T = 0.1, RK4, slab R2C with dealias='3/2-rule' must give kinetic energy 0.124953117517
to 7 decimals (``/root/reference/demo/spectral_dns_solver.py:103-105``).  It pins forward + inverse
+ 3/2-rule end to end against an absolute number.

 * CPU: the solver on the numpy oracle (checks the solver and the oracle);
 * GPU: the same solver on mpifft4py_b200 with CUDA tensors (zero-copy path of the engine)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "examples"))

import oracle  # noqa: E402
import spectral_dns_solver as sds  # noqa: E402

N = np.array([32, 32, 32], dtype=int)
L = np.array([2 * np.pi] * 3)


class _OracleSlab(object):
    """oracle.slab (one simulated rank) behind the method names the solver uses."""

    def __init__(self):
        import mpifft4py_b200 as m
        from mpifft4py_b200.comm import SelfComm
        self._host = m.Slab_R2C(N, L, SelfComm(), "double")  # host bookkeeping only (no GPU touched)
        self.Nt = tuple(int(n) for n in N)

    def __getattr__(self, name):
        return getattr(self._host, name)

    def fftn(self, u, fu, dealias=None):
        fu[...] = oracle.slab.fftn([np.asarray(u)], self.Nt, 1, dealias=dealias)[0]
        return fu

    def ifftn(self, fu, u, dealias=None):
        u[...] = oracle.slab.ifftn([np.asarray(fu)], self.Nt, 1, dealias=dealias)[0]
        return u


def test_taylor_green_known_answer_oracle():
    k = sds.solve(_OracleSlab(), np, lambda a: a, N)
    assert round(k - sds.KNOWN_ANSWER, 7) == 0, k


def test_solver_keeps_every_array_on_the_transform_device():
    """The solver under torch, on the CPU: to_xp tags its tensors (a Tensor subclass stands in for
    "lives on the engine's device"); every array handed to fftn / ifftn must carry the tag.  A bare
    xp.zeros(...) would produce an untagged (CPU) tensor -- the round-1 bug that the engine's
    `assert src.is_cuda` only caught on a GPU box."""
    import torch

    class OnDevice(torch.Tensor):
        pass

    def to_xp(a):
        return torch.from_numpy(np.ascontiguousarray(a)).as_subclass(OnDevice)

    seen = []

    class Checked(_OracleSlab):
        def fftn(self, u, fu, dealias=None):
            seen.append((type(u), type(fu)))
            assert isinstance(u, OnDevice) and isinstance(fu, OnDevice)
            fu[...] = torch.from_numpy(oracle.slab.fftn([u.as_subclass(torch.Tensor).numpy()], self.Nt, 1, dealias=dealias)[0])
            return fu

        def ifftn(self, fu, u, dealias=None):
            seen.append((type(fu), type(u)))
            assert isinstance(u, OnDevice) and isinstance(fu, OnDevice)
            u[...] = torch.from_numpy(oracle.slab.ifftn([fu.as_subclass(torch.Tensor).numpy()], self.Nt, 1, dealias=dealias)[0])
            return u

    k = sds.solve(Checked(), torch, to_xp, N, T=0.02)
    assert len(seen) > 20 and np.isfinite(k)


@pytest.mark.gpu
def test_taylor_green_known_answer_gpu():
    import torch
    import mpifft4py_b200 as m
    from mpifft4py_b200.comm import SelfComm
    torch.cuda.set_device(0)
    FFT = m.Slab_R2C(N, L, SelfComm(), "double")
    k = sds.solve(FFT, torch, lambda a: torch.from_numpy(a).cuda(), N)
    assert round(k - sds.KNOWN_ANSWER, 7) == 0, k
    kk, xx = FFT.last_launches()
    assert kk == 3  # the last transform really ran as three fused passes on the device
