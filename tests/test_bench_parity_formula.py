"""bench.py's forward-parity closed form (separable field -> outer product of 1D spectra) against the oracle,
for the decompositions and dealias modes the bench lines use.  CPU only."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import oracle  # noqa: E402


def _field(vec):
    if len(vec) == 3:
        return vec[0][:, None, None] * vec[1][None, :, None] * vec[2][None, None, :]
    return vec[0][:, None] * vec[1][None, :]


def _outer(spec):
    if len(spec) == 3:
        return spec[0][:, None, None] * spec[1][None, :, None] * spec[2][None, None, :]
    return spec[0][:, None] * spec[1][None, :]


@pytest.mark.parametrize("P", [1, 2, 4])
@pytest.mark.parametrize("dealias", [None, "3/2-rule"])
def test_slab(P, dealias):
    N = (16, 8, 12)
    g = oracle.slab.Geometry(N, P)
    vec = bench.separable_vectors(N, dealias)
    U = _field(vec)
    pad = 1.5 if dealias else 1
    u = [U[g.real_local_slice(r, pad)] for r in range(P)]
    fu = oracle.slab.fftn(u, N, P, dealias=dealias)
    ref = _outer(bench.separable_spectra(vec, N, dealias, False))
    for r in range(P):
        assert oracle.rel_l2(fu[r], ref[g.complex_local_slice(r)]) < 1e-13


@pytest.mark.parametrize("alignment", ["X", "Y"])
def test_pencil(alignment):
    N = (16, 8, 32)
    P, P1 = 4, 2
    g = oracle.pencil.Geometry(N, P, alignment=alignment, P1=P1, communication="Alltoallw")
    vec = bench.separable_vectors(N, None)
    U = _field(vec)
    u = [U[g.real_local_slice(r)] for r in range(P)]
    fu = oracle.pencil.fftn(u, N, P, P1=P1, alignment=alignment, communication="Alltoallw")
    ref = _outer(bench.separable_spectra(vec, N, None, False))
    for r in range(P):
        assert oracle.rel_l2(fu[r], ref[g.complex_local_slice(r)]) < 1e-13


def test_line():
    N = (16, 32)
    P = 2
    g = oracle.line.Geometry(N, P)
    vec = bench.separable_vectors(N, None)
    U = _field(vec)
    u = [U[g.real_local_slice(r)] for r in range(P)]
    fu = oracle.line.fft2(u, N, P)
    ref = _outer(bench.separable_spectra(vec, N, None, False))
    for r in range(P):
        assert oracle.rel_l2(fu[r], ref[g.complex_local_slice(r)]) < 1e-13


def test_other_workloads_are_the_baseline_configs_a_rank_count_can_run():
    """bench.py adds short lines for the BASELINE.json configurations besides the headline one (key other_workloads):
    pencil grids need four ranks, the 2 x 4 grid of config 3 eight."""
    import json
    import os
    import bench
    assert bench.other_workload_names("slab1024_f64", 1) == ["slab1024_f64_32", "slab256_f32", "line16384_f32"]
    assert bench.other_workload_names("slab1024_f64", 2) == bench.other_workload_names("slab1024_f64", 1)
    assert bench.other_workload_names("slab1024_f64", 4)[-2:] == ["pencilX1024_f64", "pencilY2048_f32"]
    assert bench.other_workload_names("slab1024_f64", 8)[-1] == "pencilX512_f64"
    assert "slab1024_f64_32" not in bench.other_workload_names("slab1024_f64_32", 8)
    for P in (1, 2, 4, 8):
        for n in bench.other_workload_names("slab1024_f64", P):
            assert n in bench.WORKLOADS
    # config 3 of BASELINE.json is the P1 = 2 grid
    assert bench.WORKLOADS["pencilX512_f64"][4]["P1"] == 2
    base = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "BASELINE.json")))
    assert "configs" in base
