"""Host-side API of the drop-in classes (no GPU needed): shapes, slices, meshes, wavenumbers,
error behaviour, and the C-ABI library's export table.

Multi-rank geometry runs on the thread-per-rank fake communicator of oracle/refshim (test
infrastructure); where /root/reference exists the same calls are also made on the unmodified
reference and compared value-for-value and dtype-for-dtype."""
import ctypes
import glob
import json
import os
import re
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

SHIM = os.path.join(ROOT, "oracle", "refshim")
if SHIM not in sys.path:
    sys.path.insert(0, SHIM)
from mpi4py import MPI as FAKE  # noqa: E402  (the refshim fake, never a real MPI)

import mpifft4py_b200 as m  # noqa: E402
from mpifft4py_b200 import _lib  # noqa: E402
from mpifft4py_b200.comm import SelfComm  # noqa: E402

HAVE_REF = os.path.isdir("/root/reference/mpiFFT4py")
L3 = np.array([2 * np.pi, 2 * np.pi, 2 * np.pi])


def _sl(s):
    return [[int(x.start), int(x.stop), (None if x.step is None else int(x.step))] for x in s]


def _make(kind, N, prec, comm, communication, alignment, P1):
    N = np.array(N, dtype=int)
    if kind == "slab":
        return m.Slab_R2C(N, L3, comm, prec, communication=communication)
    if kind == "pencil":
        return m.Pencil_R2C(N, L3, comm, prec, P1=P1, communication=communication, alignment=alignment)
    return m.Line_R2C(N, L3[:2], comm, prec)


FILES = sorted(glob.glob(os.path.join(GOLDEN, "*.npz")))


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_shapes_and_slices_match_reference_golden(path):
    meta = json.loads(str(np.load(path)["meta"]))
    P = meta["P"]

    def body():
        F = _make(meta["kind"], meta["N"], meta["precision"], FAKE.COMM_WORLD, meta["communication"],
                  meta["alignment"], meta["P1"])
        info = dict(rank=int(F.rank), real_shape=[int(x) for x in F.real_shape()],
                    complex_shape=[int(x) for x in F.complex_shape()],
                    real_shape_padded=[int(x) for x in F.real_shape_padded()],
                    real_local_slice=_sl(F.real_local_slice()),
                    real_local_slice_padded=_sl(F.real_local_slice(padsize=1.5)),
                    complex_local_slice=_sl(F.complex_local_slice()))
        if meta["kind"] == "pencil":
            info.update(P1=int(F.P1), P2=int(F.P2), comm0_rank=int(F.comm0_rank), comm1_rank=int(F.comm1_rank))
        if meta["kind"] != "line":
            info["work_shape_32"] = [int(x) for x in F.work_shape("3/2-rule")]
            info["work_shape_none"] = [int(x) for x in F.work_shape(None)]
        return info

    got = FAKE.run_ranks(P, body)
    for g, e in zip(got, meta["ranks"]):
        if meta["kind"] == "line":  # golden stores these with the step the generator filled in
            for k in ("real_local_slice", "real_local_slice_padded", "complex_local_slice"):
                assert [s[:2] for s in g[k]] == [s[:2] for s in e[k]]
            for k in ("real_shape", "complex_shape", "real_shape_padded"):
                assert g[k] == e[k]
        else:
            for k, v in g.items():
                assert v == e[k], (k, v, e[k])


def _same(a, b):
    if isinstance(a, (list, tuple)):
        assert type(a) is type(b) or isinstance(b, (list, tuple))
        assert len(a) == len(b)
        for x, y in zip(a, b):
            _same(x, y)
        return
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype, (a.shape, b.shape, a.dtype, b.dtype)
    assert np.array_equal(a, b)


CASES = [("slab", 1, None, None), ("slab", 4, None, None), ("pencil", 4, "X", None), ("pencil", 4, "Y", None),
         ("pencil", 8, "X", 2), ("pencil", 8, "Y", None), ("line", 1, None, None), ("line", 4, None, None)]


@pytest.mark.skipif(not HAVE_REF, reason="reference tree not present")
@pytest.mark.parametrize("kind,P,alignment,P1", CASES)
@pytest.mark.parametrize("prec", ["double", "single"])
def test_mesh_helpers_equal_reference(kind, P, alignment, P1, prec):
    import warnings
    import load_reference
    warnings.simplefilter("ignore")
    load_reference.load()
    N = (32, 64, 128) if kind != "line" else (32, 64)

    def body():
        if kind == "slab":
            from mpiFFT4py.slab import R2C
            R = R2C(np.array(N), L3, FAKE.COMM_WORLD, prec)
        elif kind == "pencil":
            from mpiFFT4py.pencil import R2C
            R = R2C(np.array(N), L3, FAKE.COMM_WORLD, prec, P1=P1, alignment=alignment)
        else:
            from mpiFFT4py.line import R2C
            R = R2C(np.array(N), L3[:2], FAKE.COMM_WORLD, prec)
        F = _make(kind, N, prec, FAKE.COMM_WORLD, "Alltoall" if kind == "pencil" else "Alltoallw", alignment, P1)
        _same(F.get_local_mesh(), R.get_local_mesh())
        if kind == "pencil" and alignment == "X":
            _same(F.get_local_wavenumbermesh(), R.get_local_wavenumbermesh())
        else:
            for kw in (dict(), dict(scaled=True), dict(scaled=True, broadcast=True),
                       dict(eliminate_highest_freq=True), dict(scaled=False, broadcast=True)):
                _same(F.get_local_wavenumbermesh(**kw), R.get_local_wavenumbermesh(**kw))
            _same(F.get_dealias_filter(), R.get_dealias_filter())
        if kind != "line":
            _same(F.complex_local_wavenumbers(), R.complex_local_wavenumbers())
            assert F.global_complex_shape() == R.global_complex_shape()
            assert F.global_complex_shape(1.5) == R.global_complex_shape(1.5)
        assert F.float is R.float and F.complex is R.complex
        assert np.array_equal(F.L, R.L) and F.L.dtype == R.L.dtype
        # every public shape method of the reference class that takes no argument (the 3/2-rule intermediates included)
        names = [n for n in dir(R) if ("shape" in n or n.startswith("complex_padded")) and callable(getattr(R, n))]
        for n in names:
            try:
                want = getattr(R, n)()
            except (TypeError, AttributeError):  # needs an argument / refers to attributes the class never sets upstream
                continue
            got = getattr(F, n)()
            assert tuple(int(x) for x in got) == tuple(int(x) for x in want), (n, got, want)
        return True

    assert all(FAKE.run_ranks(P, body))


def test_error_behaviour():
    N = np.array([8, 16, 32])
    with pytest.raises(AssertionError):
        m.Slab_R2C(N, L3[:2], SelfComm(), "double")
    with pytest.raises(AssertionError):
        m.Slab_R2C(N, L3, SelfComm(), "half")
    F = m.Slab_R2C(N, L3, SelfComm(), "double")
    with pytest.raises(AssertionError):
        F.fftn(np.zeros(F.real_shape()), np.zeros(F.complex_shape(), dtype=complex), dealias="bogus")

    class Three(SelfComm):
        def Get_size(self):
            return 3

    with pytest.raises(IOError):  # slab.py:89-91
        m.Slab_R2C(N, L3, Three(), "double")
    with pytest.raises(AssertionError):  # pencil.py:176
        m.Pencil_R2C(N, L3, SelfComm(), "double")

    def body():
        with pytest.raises(IOError):  # pencil.py:204-205: P1 = 2, P2 = 1 is not even
            m.Pencil_R2C(N, L3, FAKE.COMM_WORLD, "double")
        return True

    assert all(FAKE.run_ranks(2, body))


def test_work_arrays_and_datatypes():
    w = m.work_arrays()
    a = w[((3, 3), np.float64, 0)]
    a[:] = 1
    assert w[(a, 0)] is a and not a.any()  # zeroed on fetch
    a[:] = 2
    assert w[(a, 0, False)].sum() == 18
    b = w[(a, 1)]
    assert b is not a and b.shape == a.shape
    assert m.datatypes("single")[:2] == (np.float32, np.complex64)
    assert m.datatypes("double")[:2] == (np.float64, np.complex128)
    z = m.zeros((4, 5), dtype=np.complex128)
    assert z.shape == (4, 5) and z.dtype == np.complex128 and not z.any()


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "b200fft.h")).read()
    declared = re.findall(r"B200FFT_API\s+[\w\s\*]+?\b(b200fft_\w+)\s*\(", hdr)
    assert len(declared) >= 19
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(declared) == sorted(_lib.SYMBOLS)
    L = _lib.lib()
    assert L.b200fft_version() >= 100
    for n in (2, 3, 1024, 1536, 12288):
        assert L.b200fft_supported_length(n) == 1
    for n in (5, 7, 1000, 9, 18):
        assert L.b200fft_supported_length(n) == 0


def test_no_cpu_fallback():
    """Without a CUDA device the product must fail loudly, never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    F = m.Slab_R2C(np.array([8, 8, 8]), L3, SelfComm(), "double")
    with pytest.raises(_lib.B200FFTError):
        F.fftn(np.zeros(F.real_shape()), np.zeros(F.complex_shape(), dtype=complex))
    with pytest.raises(_lib.B200FFTError):
        m.rfftn(np.zeros((8, 8, 8)))
    src = open(os.path.join(ROOT, "mpifft4py_b200", "_engine.py")).read()
    for mod in glob.glob(os.path.join(ROOT, "mpifft4py_b200", "*.py")):
        text = open(mod).read()
        assert "import oracle" not in text and "from oracle" not in text, mod
        assert "numpy.fft.fft" not in text and "np.fft.fft(" not in text and "np.fft.rfft" not in text, mod
    assert "oracle" not in src


def test_host_allocators_own_their_storage(monkeypatch):
    """empty / zeros / work_arrays hand out arrays backed by a torch allocation (page-locked on a GPU box): the array
    itself keeps that storage alive and nothing else does (no module-level registry that would pin it for good)."""
    import gc
    import torch
    from mpifft4py_b200 import mpibase
    real_empty = torch.empty
    made = []

    def fake_empty(*a, pin_memory=False, **kw):  # a CPU box cannot page-lock: same code path, ordinary memory
        t = real_empty(*a, **kw)
        made.append(t.data_ptr())
        return t

    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch, "empty", fake_empty)
    a = mpibase.zeros((7, 9), dtype=np.complex64)
    assert made and a.ctypes.data == made[-1] and a.shape == (7, 9) and a.dtype == np.complex64 and not a.any()
    gc.collect()
    a[:] = 1 + 2j
    assert a.sum() == 63 * (1 + 2j)
    w = mpibase.work_arrays()
    b = w[((4, 5), np.float64, 0)]
    assert b.ctypes.data == made[-1] and not b.any()
    b[:] = 2
    assert w[((4, 5), np.float64, 0, False)] is b and b.sum() == 40      # kept ...
    assert not w[((4, 5), np.float64, 0)].any()                           # ... and zeroed on an ordinary fetch
    assert not [k for k, v in vars(mpibase).items() if isinstance(v, (dict, list, set)) and not k.startswith("__")]
