"""Python that only ever ran on a GPU box -- __graft_entry__.smoke(), mpifft4py_b200.tune.autotune, the device helpers
and the demo-style example driven with tensors -- on the host build of the engine (tests/cpu_engine.py) with host
stand-ins for torch's CUDA entry points (tests/test_bench_cpu_smoke.py: host_cuda)."""
import os
import sys

import numpy as np
import pytest
import torch

import cpu_engine
import mpifft4py_b200 as m
from mpifft4py_b200.comm import COMM_SELF
from test_bench_cpu_smoke import host_cuda

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
L3 = np.array([2 * np.pi] * 3)


@pytest.fixture
def engine(monkeypatch):
    cleanup = cpu_engine.install(monkeypatch)
    host_cuda(monkeypatch)
    yield cleanup.calls
    cleanup()


def test_smoke_entry_point(engine, capsys):
    import __graft_entry__ as g
    g.smoke()
    assert "smoke: rel L2 fftn" in capsys.readouterr().out
    assert engine["b200fft_exec_forward"] == 2 and engine["b200fft_exec_inverse"] == 2


def test_autotune_single_rank(engine):
    """One rank: the exchange candidates mean nothing and are dropped, the layout candidate is timed against the default and
    must reproduce its result; the object ends up with a working plan either way."""
    F = m.Slab_R2C(np.array([16, 16, 32]), L3, COMM_SELF, "double")
    rep = m.tune.autotune(F, candidates=m.tune.CANDIDATES["patient"], reps=1)
    names = [c["name"] for c in rep["candidates"]]
    assert names == ["default", "natural_layout"] and all(c["ok"] and c["error"] <= 1e-12 for c in rep["candidates"])
    assert rep["chosen"] in names
    u = torch.rand((16, 16, 32), dtype=torch.float64)
    fu = F.fftn(u, torch.zeros((16, 16, 17), dtype=torch.complex128))
    assert np.allclose(fu.numpy(), np.fft.rfftn(u.numpy()), rtol=0, atol=1e-11)
    rep32 = m.tune.autotune(F, dealias="3/2-rule", reps=1)
    assert [c["name"] for c in rep32["candidates"]] == ["default"] and rep32["chosen"] == "default"


def test_example_solver_with_tensors(engine):
    """examples/spectral_dns_solver.py as a GPU run drives it: every array a tensor on the transform's device, the
    Taylor-Green known answer of demo/spectral_dns_solver.py:105 at the end."""
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    import spectral_dns_solver as sds
    N = np.array([32, 32, 32], dtype=int)
    F = m.Slab_R2C(N, L3, COMM_SELF, "double")
    k = sds.solve(F, torch, lambda a: torch.from_numpy(a), N)
    assert round(float(k) - sds.KNOWN_ANSWER, 7) == 0
    assert engine["b200fft_exec_forward"] >= 120 and engine["b200fft_exec_inverse"] >= 240


def test_ns_solver_kernels_through_the_python_layer(engine, monkeypatch):
    """mpifft4py_b200.ns.Solver on a transform object (not the bare host build as in test_zz_ns_known_answer.py): the
    library's three elementwise kernels per stage around the object's own transforms."""
    monkeypatch.setattr(m.ns.Solver, "_device", lambda self: torch.device("cpu"))
    N = np.array([32, 32, 32], dtype=int)
    F = m.Slab_R2C(N, L3, COMM_SELF, "double")
    S = m.ns.Solver(F, nu=0.000625, dt=0.01)
    X = [torch.from_numpy(np.ascontiguousarray(np.broadcast_to(x, F.real_shape()))) for x in F.get_local_mesh()]
    S.set_velocity(torch.stack([torch.sin(X[0]) * torch.cos(X[1]) * torch.cos(X[2]),
                                -torch.cos(X[0]) * torch.sin(X[1]) * torch.cos(X[2]), torch.zeros_like(X[0])]))
    for _ in range(10):
        S.step()
    assert round(float(S.kinetic_energy()) - 0.124953117517, 7) == 0
