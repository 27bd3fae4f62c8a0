"""Stream / event schedules of the plan programs (tests/emu/emu.cpp: emu_check_schedule).  The emulator
executes programs in program order, the device on two streams ordered by events: this derives the
happens-before relation the device enforces and requires it between every pair of steps that touch
overlapping parts of a local buffer with at least one write, that every awaited event is recorded by an
earlier step, and that everything is joined back into the caller's stream."""
import ctypes as C
import itertools

import pytest

import emu_util
from mpifft4py_b200 import _cdefs as D
from test_emu_plans import _desc

MODES = (D.DEALIAS_NONE, D.DEALIAS_3_2, D.DEALIAS_2_3)


def _check(d):
    lib = emu_util.load()
    for inverse in (0, 1):
        for m in MODES:
            assert lib.emu_check_schedule(C.byref(d), inverse, m) == 0, (inverse, m)


@pytest.mark.parametrize("chunks", [0, 1, 2, 4])
@pytest.mark.parametrize("pipeline", [D.PIPELINE_X, D.PIPELINE_KZ])
@pytest.mark.parametrize("transport", [D.TRANSPORT_NCCL, D.TRANSPORT_P2P, D.TRANSPORT_STORE])
@pytest.mark.parametrize("P", [1, 2, 4])
@pytest.mark.parametrize("kind", [D.SLAB, D.SLAB_C2C])
def test_slab_schedules(kind, P, transport, pipeline, chunks):
    if P == 1 and (transport or pipeline or chunks):
        pytest.skip("single rank: no exchange")
    for layout in ((D.LAYOUT_YBLOCK, D.LAYOUT_NATURAL) if P == 1 else (D.LAYOUT_YBLOCK,)):
        _check(_desc(kind, (32, 16, 64), P, "double", chunks=chunks, pipeline=pipeline, transport=transport, layout=layout))


def test_headline_sizes():
    for P, transport, pipeline, chunks in itertools.product(
            (1, 2, 8), (D.TRANSPORT_NCCL, D.TRANSPORT_P2P, D.TRANSPORT_STORE), (D.PIPELINE_X, D.PIPELINE_KZ), (0, 8)):
        if P == 1 and (transport or pipeline or chunks):
            continue
        _check(_desc(D.SLAB, (1024, 1024, 1024), P, "double", chunks=chunks, pipeline=pipeline, transport=transport))


@pytest.mark.parametrize("transport", [D.TRANSPORT_NCCL, D.TRANSPORT_P2P, D.TRANSPORT_STORE])
@pytest.mark.parametrize("P,P1,P2", [(4, 2, 2), (8, 4, 2), (8, 2, 4)])
@pytest.mark.parametrize("kind", [D.PENCIL_X, D.PENCIL_Y])
def test_pencil_schedules(kind, P, P1, P2, transport):
    for drop in (0, 1):
        _check(_desc(kind, (16, 16, 32), P, "double", P1, P2, drop, transport=transport))


@pytest.mark.parametrize("transport", [D.TRANSPORT_NCCL, D.TRANSPORT_P2P, D.TRANSPORT_STORE])
@pytest.mark.parametrize("P", [1, 2, 4, 8])
def test_line_schedules(P, transport):
    _check(_desc(D.LINE, (64, 32), P, "double", transport=transport if P > 1 else 0))
