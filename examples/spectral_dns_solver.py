#!/usr/bin/env python
"""Taylor-Green vortex, pseudo-spectral Navier-Stokes, RK4 -- the reference's demo
(``/root/reference/demo/spectral_dns_solver.py:53-105``) on the B200 engine.

Same equations, same 9 transforms per Runge-Kutta stage with ``dealias='3/2-rule'``, same known
answer (kinetic energy 0.124953117517 after T = 0.1 at 32^3, ``:105``).  Everything stays on the
device: the transforms are mpifft4py_b200's kernels, called with CUDA tensors (zero copy); the
cross / curl / projection arithmetic between them is plain tensor arithmetic of the caller.

    python examples/spectral_dns_solver.py                 # 1 GPU, the demo's array arithmetic between the transforms
    python examples/spectral_dns_solver.py --kernels       # mpifft4py_b200.ns.Solver: three library kernels per RK stage
    python examples/spectral_dns_solver.py --kernels --graph   # ... each time step replayed from a CUDA graph
    torchrun --nproc-per-node 4 examples/spectral_dns_solver.py [--kernels]

`solve` is written against a tiny array-module interface so that tests can also run it on the CPU
oracle (numpy) and check the solver logic without a GPU.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

KNOWN_ANSWER = 0.124953117517  # demo/spectral_dns_solver.py:105


def solve(FFT, xp, to_xp, N, nu=0.000625, T=0.1, dt=0.01, dealias="3/2-rule"):
    """Integrate to time T; returns this rank's share of sum(U*U)/(N0*N1*N2)/2.

    FFT   : object with the mpiFFT4py slab API (fftn/ifftn/real_shape/complex_shape/work_shape/
            get_local_mesh/get_local_wavenumbermesh)
    xp    : array module (numpy or torch) providing empty/zeros/sin/cos/where/sum
    to_xp : converts a numpy array to an xp array on the right device
    """
    rshape = tuple(int(s) for s in FFT.real_shape())
    cshape = tuple(int(s) for s in FFT.complex_shape())
    wshape = tuple(int(s) for s in FFT.work_shape(dealias))
    rdt, cdt = np.float64, np.complex128

    def zeros(shape, dtype):
        # every state array goes through to_xp, so it lives where the transform's arrays live (a bare
        # xp.zeros under torch would be a CPU tensor, which the engine refuses)
        return to_xp(np.zeros(shape, dtype=dtype))

    U = zeros((3,) + rshape, rdt)
    U_hat = zeros((3,) + cshape, cdt)
    U_hat0 = zeros((3,) + cshape, cdt)
    U_hat1 = zeros((3,) + cshape, cdt)
    dU = zeros((3,) + cshape, cdt)
    U_d = zeros((3,) + wshape, rdt)
    curl_d = zeros((3,) + wshape, rdt)
    tmp_r = zeros(wshape, rdt)
    tmp_c = zeros(cshape, cdt)
    X = [to_xp(np.ascontiguousarray(np.broadcast_to(x, rshape))) for x in FFT.get_local_mesh()]
    K = [to_xp(np.ascontiguousarray(np.broadcast_to(np.asarray(k, dtype=np.float64), cshape)))
         for k in FFT.get_local_wavenumbermesh(scaled=True)]
    K2 = K[0] * K[0] + K[1] * K[1] + K[2] * K[2]
    K2_safe = xp.where(K2 == 0, xp.ones_like(K2), K2)
    K_over_K2 = [k / K2_safe for k in K]
    a = [1. / 6., 1. / 3., 1. / 3., 1. / 6.]
    b = [0.5, 0.5, 1.]

    def cross(x, y, z):  # :53-58
        for i, (p, q) in enumerate(((1, 2), (2, 0), (0, 1))):
            tmp_r[...] = x[p] * y[q] - x[q] * y[p]
            FFT.fftn(tmp_r, z[i], dealias)
        return z

    def curl(x, z):  # :60-64
        for i, (p, q) in enumerate(((1, 2), (2, 0), (0, 1))):
            tmp_c[...] = 1j * (K[p] * x[q] - K[q] * x[p])
            FFT.ifftn(tmp_c, z[i], dealias)
        return z

    def compute_rhs(rhs):  # :66-77
        for i in range(3):
            FFT.ifftn(U_hat[i], U_d[i], dealias)
        curl(U_hat, curl_d)
        cross(U_d, curl_d, rhs)
        P_hat = rhs[0] * K_over_K2[0] + rhs[1] * K_over_K2[1] + rhs[2] * K_over_K2[2]
        for i in range(3):
            rhs[i] -= P_hat * K[i]
            rhs[i] -= nu * K2 * U_hat[i]
        return rhs

    U[0] = xp.sin(X[0]) * xp.cos(X[1]) * xp.cos(X[2])  # :80-82
    U[1] = -xp.cos(X[0]) * xp.sin(X[1]) * xp.cos(X[2])
    U[2] = 0
    for i in range(3):
        FFT.fftn(U[i], U_hat[i])
    t = 0.0
    while t < T - 1e-8:  # :87-98
        t += dt
        U_hat1[...] = U_hat
        U_hat0[...] = U_hat
        for rk in range(4):
            compute_rhs(dU)
            if rk < 3:
                U_hat[...] = U_hat0 + b[rk] * dt * dU
            U_hat1 += a[rk] * dt * dU
        U_hat[...] = U_hat1
    for i in range(3):
        FFT.ifftn(U_hat[i], U[i])
    return float(xp.sum(U * U)) / float(N[0]) / float(N[1]) / float(N[2]) / 2.0


def main():
    import torch
    import torch.distributed as dist
    import mpifft4py_b200 as m
    from mpifft4py_b200.comm import SelfComm, world
    P = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if P > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        comm = world()
    else:
        comm = SelfComm()
    N = np.array([32, 32, 32], dtype=int)
    L = np.array([2 * np.pi] * 3)
    FFT = m.Slab_R2C(N, L, comm, "double")
    if "--kernels" in sys.argv:
        S = m.ns.Solver(FFT, nu=0.000625, dt=0.01, graph="--graph" in sys.argv)
        X = [torch.from_numpy(np.ascontiguousarray(np.broadcast_to(x, FFT.real_shape()))).cuda() for x in FFT.get_local_mesh()]
        S.set_velocity(torch.stack([torch.sin(X[0]) * torch.cos(X[1]) * torch.cos(X[2]),
                                    -torch.cos(X[0]) * torch.sin(X[1]) * torch.cos(X[2]), torch.zeros_like(X[0])]))
        for _ in range(10):
            S.step()
        k = S.kinetic_energy()
    else:
        k = solve(FFT, torch, lambda a: torch.from_numpy(a).cuda(), N)
    k = comm.reduce(k)
    if comm.Get_rank() == 0:
        print("kinetic energy %.12f (reference %.12f, difference %.2e)" % (k, KNOWN_ANSWER, k - KNOWN_ANSWER))
        assert round(k - KNOWN_ANSWER, 7) == 0
    if P > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
