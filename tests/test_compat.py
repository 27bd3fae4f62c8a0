"""mpifft4py_b200.compat: a program written for mpiFFT4py imports this package under the reference's names."""
import subprocess
import sys

import numpy as np

PROGRAM = r'''
import sys
import numpy as np
import mpifft4py_b200.compat
mpifft4py_b200.compat.install()

# ---- from here on: source as a user of the reference wrote it (mpiFFT4py/__init__.py:1-8, tests/test_FFT.py:7-13) ----
from mpi4py import MPI
from mpiFFT4py import Slab_R2C, Pencil_R2C, Line_R2C, work_arrays, datatypes, empty, zeros, fftfreq, rfftfreq
from mpiFFT4py import rfft2, rfftn, irfftn, irfft2, fftn, ifftn, irfft, ifft, rfft, fft, dct
from mpiFFT4py.pencil import R2C as P_R2C
from mpiFFT4py.slab import R2C as S_R2C, C2C
from mpiFFT4py.line import R2C as L_R2C
import mpiFFT4py

comm = MPI.COMM_WORLD
assert comm.Get_size() == 1 and comm.Get_rank() == 0 and MPI.COMM_SELF.Get_size() == 1
assert MPI.Compute_dims(8, 2) == [4, 2] and MPI.Compute_dims(16, 2) == [4, 4] and MPI.Compute_dims(4, 2) == [2, 2]
assert comm.reduce(3.0, op=MPI.MIN) == 3.0
N = np.array([8, 16, 32]); L = np.array([2 * np.pi] * 3)
F = Slab_R2C(N, L, comm, "double", communication="Alltoallw")
assert type(F) is S_R2C and F.real_shape() == (8, 16, 32) and F.complex_shape() == (8, 16, 17)
assert mpiFFT4py.__version__ and mpiFFT4py.slab.R2C is S_R2C and C2C(N, L, comm, "single").float is np.float32
assert L_R2C(N[:2], L[:2], MPI.COMM_SELF, "single").complex_shape() == (8, 9)
assert datatypes("double")[:2] == (np.float64, np.complex128) and zeros((2, 3)).shape == (2, 3)
try:
    P_R2C(N, L, comm, "double")
    raise SystemExit("pencil on one rank must be refused")
except AssertionError:
    pass
mpifft4py_b200.compat.uninstall()
assert "mpiFFT4py" not in sys.modules and "mpi4py" not in sys.modules
print("COMPAT_OK")
'''


def test_a_reference_program_imports_unchanged():
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", PROGRAM], cwd=root, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300)
    text = out.stdout.decode("utf-8", "replace")
    assert out.returncode == 0 and "COMPAT_OK" in text, text[-3000:]


def test_stand_in_leaves_a_real_mpi4py_alone(monkeypatch):
    import types
    from mpifft4py_b200 import compat
    real, real_mpi = types.ModuleType("mpi4py"), types.ModuleType("mpi4py.MPI")
    real.MPI = real_mpi
    monkeypatch.setitem(sys.modules, "mpi4py", real)
    monkeypatch.setitem(sys.modules, "mpi4py.MPI", real_mpi)
    try:
        compat.install()
        assert sys.modules["mpi4py"] is real and sys.modules["mpi4py.MPI"] is real_mpi
        import mpiFFT4py
        assert mpiFFT4py.fftfreq is np.fft.fftfreq
    finally:
        compat.uninstall()
    assert sys.modules["mpi4py"] is real  # uninstall only removes its own stand-in
