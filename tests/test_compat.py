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


SCRIPT = r'''
import sys
import numpy as np
from mpi4py import MPI
from mpiFFT4py.slab import R2C
from mpiFFT4py import Line_R2C
comm = MPI.COMM_WORLD
N = np.array([8, 16, 32]); L = np.array([2 * np.pi] * 3)
F = R2C(N, L, comm, "double")
P, r = comm.Get_size(), comm.Get_rank()
assert F.real_shape() == (8 // P, 16, 32) and F.complex_local_slice()[1] == slice(r * 16 // P, (r + 1) * 16 // P, 1)
tot = comm.reduce(r + 1.0)
if r == 0:
    assert tot == P * (P + 1) / 2
print("SCRIPT_OK", r, P, sys.argv[1:], __name__)
'''


def _launch(tmp_path, nproc):
    import os
    import socket
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "user_program.py"
    script.write_text(SCRIPT)
    if nproc == 1:
        cmd = [sys.executable, "-m", "mpifft4py_b200.compat", str(script), "--flag", "7"]
    else:
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
        s.close()
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
               "--master-port", str(port), "-m", "mpifft4py_b200.compat", str(script), "--flag", "7"]
    env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""), OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, cwd=str(tmp_path), env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300)
    return out.returncode, out.stdout.decode("utf-8", "replace")


def test_launcher_runs_a_script_as_main(tmp_path):
    rc, text = _launch(tmp_path, 1)
    assert rc == 0 and "SCRIPT_OK 0 1 ['--flag', '7'] __main__" in text, text[-2000:]


def test_launcher_under_torchrun(tmp_path):
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("CPU check of the launcher (gloo); on a GPU box the ranks would want one GPU each")
    rc, text = _launch(tmp_path, 2)
    assert rc == 0 and text.count("SCRIPT_OK") == 2 and "SCRIPT_OK 1 2" in text, text[-2000:]
