"""The reference's OWN test functions -- /root/reference/tests/test_FFT.py, loaded unmodified from where it lies --
executed on mpifft4py_b200's classes: ``mpifft4py_b200.compat.install()`` makes ``mpiFFT4py`` / ``mpiFFT4py.slab`` /
``.pencil`` / ``.line`` resolve to this package's modules and ``mpi4py.MPI`` to its communicators while that file is
imported, then every test function is called with objects built the way its fixtures build them (``:36-56``), for every
fixture parameter, on 1 rank and on 4 ranks (threads).  Likewise the reference's demo solver, run as a script to its
asserted known answer.

There is no GPU here, so each test runs twice (fixture ``device``): with the ORACLE answering the C-ABI calls
(tests/fake_device.py) -- which pins the drop-in claim at the level of a caller's source code: names, signatures,
return conventions, attributes, shapes, slices are what upstream's own tests and demo expect -- and with the product's
own stack built for the host (tests/cpu_engine.py: the real Transform._run / _ensure_plan handshake, the real C-ABI
layer and plan programs, the kernels' phase bodies in the emulator; only the CUDA runtime is a stand-in) -- so the
numbers upstream's assertions see there were computed by the engine's code.  On the device the same procedures are
tests/ref_procedures.py (the reference tree does not exist on the GPU box).  Skipped where /root/reference is absent."""
import importlib.util
import os
import sys
import threading

import numpy as np
import pytest

import cpu_engine
import fake_device
import mpifft4py_b200 as m
import ref_procedures as rp
from test_ref_procedures_oracle import ThreadComm, ThreadWorld, numpy_serial, SERIAL

REF_TEST = "/root/reference/tests/test_FFT.py"
pytestmark = pytest.mark.skipif(not os.path.isfile(REF_TEST), reason="reference tree not present")

class WorldProxy(object):
    """``MPI.COMM_WORLD`` of the loaded test module: each thread-rank sees its own communicator."""

    def __init__(self, size):
        self.size = size
        self.local = threading.local()

    def _c(self):
        return getattr(self.local, "comm", None)

    def Get_size(self):
        return self.size

    def Get_rank(self):
        return self._c().Get_rank() if self._c() is not None else 0

    def __getattr__(self, name):
        return getattr(self._c(), name)


def load_reference_tests(world, numpy_serial_functions=True):
    """Import the reference's test file after mpifft4py_b200.compat.install() -- the product's own way of running a
    program written for mpiFFT4py unedited -- with this test's communicator as COMM_WORLD and, there being no GPU,
    numpy.fft behind the serial function names."""
    from mpifft4py_b200 import compat
    saved = {k: sys.modules.get(k) for k in compat._NAMES + ("mpi4py", "mpi4py.MPI")}
    gone = [n for n in ("int", "float") if not hasattr(np, n)]  # `from numpy import ... int ...` (:5), removed in numpy 1.24
    try:
        pkg = compat.install(mpi4py=True)
        sys.modules["mpi4py.MPI"].COMM_WORLD = world
        if numpy_serial_functions:
            for name, fn in SERIAL.items():
                setattr(pkg, name, numpy_serial(fn))
        for n in gone:
            setattr(np, n, {"int": int, "float": float}[n])
        spec = importlib.util.spec_from_file_location("reference_test_FFT_%d" % world.Get_size(), REF_TEST)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for n in gone:
            delattr(np, n)
        compat.uninstall()
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
    return mod


def run_module(mod, comm):
    """What pytest would run for this communicator: each test function x each parameter of its fixture."""
    P = comm.Get_size()
    three_d, lines, c2cs = rp.params(P)
    assert sorted(mod.params) == sorted(three_d)  # the module's own parameter list for this rank count (:24-31)
    ran = 0
    for p in three_d:
        F = rp.make(p, comm)
        mod.test_FFT(F)
        mod.test_FFT_padded(F)
        ran += 2
    for p in lines:
        F = rp.make(p, comm)
        mod.test_FFT2(F)
        mod.test_FFT2_padded(F)
        ran += 2
    for p in c2cs:
        mod.test_FFT_C2C(rp.make(p, comm))
        ran += 1
    return ran


@pytest.fixture(params=["oracle", "engine"])
def device(request, monkeypatch):
    """ "oracle": the oracle answers the C-ABI calls under the transform methods and numpy.fft the serial function names
    (tests/fake_device.py).  "engine": nothing of the product is replaced but the CUDA runtime -- its C-ABI layer, plan
    programs and kernel phase bodies run in the host build (tests/cpu_engine.py), serial functions included."""
    if request.param == "oracle":
        fake_device.install(monkeypatch)
        yield request.param
    else:
        cleanup = cpu_engine.install(monkeypatch)
        yield request.param
        cleanup()


def test_reference_tests_pass_on_one_rank(device):
    world = WorldProxy(1)
    world.local.comm = m.comm.COMM_SELF
    mod = load_reference_tests(world, device == "oracle")
    assert {"test_FFT", "test_FFT2", "test_FFT2_padded", "test_FFT_padded", "test_FFT_C2C"} <= set(dir(mod))
    assert run_module(mod, m.comm.COMM_SELF) == 2 * 4 + 2 * 2 + 2


def test_reference_tests_pass_on_four_ranks(device):
    P = 4
    tw = ThreadWorld(P)
    world = WorldProxy(P)
    mod = load_reference_tests(world, device == "oracle")
    assert len(mod.params) == 16
    counts = [None] * P

    def rank_main(r):
        comm = ThreadComm(tw, r)
        world.local.comm = comm
        try:
            counts[r] = run_module(mod, comm)
        except BaseException as e:  # noqa: BLE001
            tw.failed.append((r, repr(e)[:400]))
            tw.barrier.abort()

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(P)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=900)
    real = [f for f in tw.failed if "BrokenBarrierError" not in f[1]]
    assert not tw.failed, real or tw.failed
    assert counts == [2 * 16 + 2 * 2 + 2] * P


def test_reference_demo_runs_unedited_and_reaches_its_known_answer(device):
    """/root/reference/demo/spectral_dns_solver.py as it lies there: a Taylor-Green run of ten RK4 steps with the
    3/2-rule (90 inverse + 90 forward padded transforms through ``work_arrays``, ``get_local_mesh``,
    ``get_local_wavenumbermesh(scaled=True)``, ``P_hat*K`` on the sparse wavenumber list) that ends in
    ``assert round(k - 0.124953117517, 7) == 0`` (:105).  The script no longer runs on the reference itself with a
    current numpy (the ragged ``P_hat*K``); here it does, through compat.install() and the classes' mesh lists."""
    import runpy
    from mpifft4py_b200 import compat
    demo = "/root/reference/demo/spectral_dns_solver.py"
    saved = {k: sys.modules.get(k) for k in compat._NAMES + ("mpi4py", "mpi4py.MPI")}
    try:
        compat.install(mpi4py=True)
        ns = runpy.run_path(demo, run_name="reference_demo")
    finally:
        compat.uninstall()
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
    assert ns["tstep"] == 10 and abs(float(ns["k"]) - 0.124953117517) < 5e-8
    assert type(ns["FFT"]).__module__ == "mpifft4py_b200.slab"
