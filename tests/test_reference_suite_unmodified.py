"""The reference's OWN test functions -- /root/reference/tests/test_FFT.py, loaded unmodified from where it lies --
executed on mpifft4py_b200's classes: ``mpifft4py_b200.compat.install()`` makes ``mpiFFT4py`` / ``mpiFFT4py.slab`` /
``.pencil`` / ``.line`` resolve to this package's modules and ``mpi4py.MPI`` to its communicators while that file is
imported, then every test function is
called with objects built the way its fixtures build them (``:36-56``), for every fixture parameter, on 1 rank and on
4 ranks (threads).  What this proves is the drop-in claim at the level of a caller's source code: names, signatures,
return conventions, attributes (``FFT.N``, ``.float``, ``.comm``, ``.communication`` ...), shapes and slices are what
upstream's own tests expect.  There is no GPU here, so the ORACLE answers the C-ABI calls under the transform methods
(tests/fake_device.py) and numpy.fft the serial functions; the kernels' parity is the business of the `-m gpu`
tests, where tests/ref_procedures.py restates these same procedures (the reference tree does not exist on the GPU
box).  Skipped where /root/reference is absent."""
import importlib.util
import os
import sys
import threading

import numpy as np
import pytest

import fake_device
import mpifft4py_b200 as m
import ref_procedures as rp
from test_ref_procedures_oracle import ThreadComm, ThreadWorld

REF_TEST = "/root/reference/tests/test_FFT.py"
pytestmark = pytest.mark.skipif(not os.path.isfile(REF_TEST), reason="reference tree not present")

SERIAL = {"rfftn": np.fft.rfftn, "irfftn": np.fft.irfftn, "rfft2": np.fft.rfft2, "irfft2": np.fft.irfft2, "fftn": np.fft.fftn,
          "ifftn": np.fft.ifftn, "irfft": np.fft.irfft, "ifft": np.fft.ifft}


def _serial(npfn):
    def f(a, b, axes=None, axis=None, **kw):
        b[...] = npfn(a, axes=axes) if axis is None else npfn(a, axis=axis)
        return b
    return f


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


def load_reference_tests(world):
    """Import the reference's test file after mpifft4py_b200.compat.install() -- the product's own way of running a
    program written for mpiFFT4py unedited -- with this test's communicator as COMM_WORLD and, there being no GPU,
    numpy.fft behind the serial function names."""
    from mpifft4py_b200 import compat
    saved = {k: sys.modules.get(k) for k in compat._NAMES + ("mpi4py", "mpi4py.MPI")}
    gone = [n for n in ("int", "float") if not hasattr(np, n)]  # `from numpy import ... int ...` (:5), removed in numpy 1.24
    try:
        pkg = compat.install(mpi4py=True)
        sys.modules["mpi4py.MPI"].COMM_WORLD = world
        for name, fn in SERIAL.items():
            setattr(pkg, name, _serial(fn))
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


@pytest.fixture
def oracle_device(monkeypatch):
    fake_device.install(monkeypatch)


def test_reference_tests_pass_on_one_rank(oracle_device):
    world = WorldProxy(1)
    world.local.comm = m.comm.COMM_SELF
    mod = load_reference_tests(world)
    assert {"test_FFT", "test_FFT2", "test_FFT2_padded", "test_FFT_padded", "test_FFT_C2C"} <= set(dir(mod))
    assert run_module(mod, m.comm.COMM_SELF) == 2 * 4 + 2 * 2 + 2


def test_reference_tests_pass_on_four_ranks(oracle_device):
    P = 4
    tw = ThreadWorld(P)
    world = WorldProxy(P)
    mod = load_reference_tests(world)
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


def test_reference_demo_runs_unedited_and_reaches_its_known_answer(oracle_device, capsys):
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
