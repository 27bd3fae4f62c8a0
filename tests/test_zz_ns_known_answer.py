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


def _taylor_green(FFT, xp, to_xp):
    X = [to_xp(np.ascontiguousarray(np.broadcast_to(x, FFT.real_shape()))) for x in FFT.get_local_mesh()]
    U = xp.stack([xp.sin(X[0]) * xp.cos(X[1]) * xp.cos(X[2]), -xp.cos(X[0]) * xp.sin(X[1]) * xp.cos(X[2]),
                  xp.zeros_like(X[0])])
    return U


def test_solver_kernels_host_build_known_answer():
    """mpifft4py_b200.ns.Solver -- three elementwise kernels per RK stage instead of a dozen array passes -- with the
    kernels' point functions (csrc/ns_ops.cuh) run as host loops by the host build of the library and the transforms
    by the oracle: same known answer, and its right-hand side equals the demo-style solver's array arithmetic."""
    import torch
    import host_shim_util
    import mpifft4py_b200 as m

    class TorchOracle(_OracleSlab):
        def fftn(self, u, fu, dealias=None):
            fu[...] = torch.from_numpy(oracle.slab.fftn([u.numpy()], self.Nt, 1, dealias=dealias)[0])
            return fu

        def ifftn(self, fu, u, dealias=None):
            u[...] = torch.from_numpy(oracle.slab.ifftn([fu.numpy()], self.Nt, 1, dealias=dealias)[0])
            return u

    FFT = TorchOracle()
    S = m.ns.Solver(FFT, nu=0.000625, dt=0.01, lib=host_shim_util.load())
    S.set_velocity(_taylor_green(FFT, torch, torch.from_numpy))
    # right-hand side against plain array arithmetic (demo :53-77)
    K = [torch.from_numpy(np.ascontiguousarray(np.broadcast_to(np.asarray(k, dtype=np.float64), FFT.complex_shape())))
         for k in FFT.get_local_wavenumbermesh(scaled=True)]
    K2 = K[0] ** 2 + K[1] ** 2 + K[2] ** 2
    Uh = S.U_hat.clone()
    Ud = torch.stack([FFT.ifftn(Uh[i], torch.zeros(FFT.real_shape_padded(), dtype=torch.float64), "3/2-rule") for i in range(3)])
    ch = torch.stack([1j * (K[1] * Uh[2] - K[2] * Uh[1]), 1j * (K[2] * Uh[0] - K[0] * Uh[2]), 1j * (K[0] * Uh[1] - K[1] * Uh[0])])
    cd = torch.stack([FFT.ifftn(ch[i], torch.zeros(FFT.real_shape_padded(), dtype=torch.float64), "3/2-rule") for i in range(3)])
    cr = torch.stack([Ud[1] * cd[2] - Ud[2] * cd[1], Ud[2] * cd[0] - Ud[0] * cd[2], Ud[0] * cd[1] - Ud[1] * cd[0]])
    dU = torch.stack([FFT.fftn(cr[i], torch.zeros(FFT.complex_shape(), dtype=torch.complex128), "3/2-rule") for i in range(3)])
    P = (dU[0] * K[0] + dU[1] * K[1] + dU[2] * K[2]) / torch.where(K2 == 0, torch.ones_like(K2), K2)
    ref = torch.stack([dU[i] - P * K[i] - 0.000625 * K2 * Uh[i] for i in range(3)])
    got = S.rhs().clone()
    assert float(torch.linalg.vector_norm(got - ref) / torch.linalg.vector_norm(ref)) < 1e-13
    for _ in range(10):
        S.step()
    assert round(S.kinetic_energy() - sds.KNOWN_ANSWER, 7) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("graph", [False, True])
def test_solver_kernels_known_answer_gpu(graph):
    """The same on the device: library kernels + fused transforms, eager and replayed from a CUDA graph."""
    import torch
    import mpifft4py_b200 as m
    from mpifft4py_b200.comm import SelfComm
    torch.cuda.set_device(0)
    FFT = m.Slab_R2C(N, L, SelfComm(), "double")
    S = m.ns.Solver(FFT, nu=0.000625, dt=0.01, graph=graph)
    S.set_velocity(_taylor_green(FFT, torch, lambda a: torch.from_numpy(a).cuda()))
    for _ in range(10):
        S.step()
    assert (S.graph is not None) == graph
    assert round(S.kinetic_energy() - sds.KNOWN_ANSWER, 7) == 0


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
