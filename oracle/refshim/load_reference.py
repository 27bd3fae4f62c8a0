"""Import the UNMODIFIED reference (``/root/reference/mpiFFT4py``) in this container.

TEST INFRASTRUCTURE ONLY (golden-vector generation and oracle pinning).  The reference is
Python-2-era code: it needs six compat patches applied *to the interpreter, not to its
sources*  plus the fake ``mpi4py`` next to this file.  Its only native
module, ``mpiFFT4py/cython/maths.pyx``, is compiled with Cython into a scratch copy under
``/tmp`` (``/root/reference`` is read-only); if that fails a 2-function numpy stand-in with
the same semantics (``maths.pyx:9-31``) is injected instead.

Nothing here travels to the GPU box: ``/root/reference`` does not exist there.
"""
import builtins
import collections
import collections.abc
import importlib
import os
import shutil
import subprocess
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("MPIFFT4PY_REFERENCE", "/root/reference")
SCRATCH = os.environ.get("MPIFFT4PY_REF_SCRATCH", "/tmp/mpifft4py_ref_build")
_HERE = os.path.dirname(os.path.abspath(__file__))
# `pip install --no-deps --target baseline/_ref <copy of /root/reference>`: the
# unmodified reference with its Cython extension already built.  Git-ignored, but it travels to
# the GPU box, where it is what bench.py's reference arm times.
INSTALLED = os.path.join(os.path.dirname(os.path.dirname(_HERE)), "baseline", "_ref")


def installed_available():
    return os.path.isfile(os.path.join(INSTALLED, "mpiFFT4py", "slab.py"))


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mpiFFT4py")) or installed_available()


def _apply_compat_patches():
    if not hasattr(np, "float"):
        np.float = float
    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(collections, "MutableMapping"):
        collections.MutableMapping = collections.abc.MutableMapping
    if not hasattr(builtins, "xrange"):
        builtins.xrange = range
    if not getattr(np.meshgrid, "_refshim", False):
        _meshgrid = np.meshgrid

        def meshgrid(*a, **k):
            return list(_meshgrid(*a, **k))

        meshgrid._refshim = True
        np.meshgrid = meshgrid
    if not getattr(np.ogrid, "_refshim", False):
        _ogrid = np.ogrid

        class _OGrid(object):
            _refshim = True

            def __getitem__(self, key):
                out = _ogrid[key]
                return list(out) if isinstance(out, tuple) else out

        np.ogrid = _OGrid()


def _build_scratch_copy():
    """Copy the reference to /tmp and build its Cython extension there.  Returns import root."""
    marker = os.path.join(SCRATCH, ".built")
    if os.path.exists(marker):
        return SCRATCH
    if os.path.isdir(SCRATCH):
        shutil.rmtree(SCRATCH)
    shutil.copytree(REFERENCE_ROOT, SCRATCH)
    pyx_dir = os.path.join(SCRATCH, "mpiFFT4py", "cython")
    try:
        subprocess.run(
            [sys.executable, "-m", "cython", "-3", "maths.pyx"], cwd=pyx_dir, check=True,
            stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
        import sysconfig
        inc = sysconfig.get_paths()["include"]
        ext = sysconfig.get_config_var("EXT_SUFFIX")
        subprocess.run(
            ["gcc", "-shared", "-fPIC", "-O2", "-w", "-I", inc, "-I", np.get_include(),
             "maths.c", "-o", "maths" + ext], cwd=pyx_dir, check=True,
            stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
        with open(marker, "w") as f:
            f.write("cython\n")
    except Exception as e:  # noqa: BLE001
        with open(marker, "w") as f:
            f.write("standin: %r\n" % (e,))
    return SCRATCH


def _inject_maths_standin():
    """numpy stand-in for mpiFFT4py.cython.maths (maths.pyx:9-31) if the build failed."""
    m = types.ModuleType("mpiFFT4py.cython.maths")

    def dealias_filter(fu, dealias):
        s = dealias.shape
        fu[:s[0], :s[1], :s[2]] *= dealias
        return fu

    def transpose_Uc(Uc_hatT, U_mpi, num_processes, Np0, Np1, Nf):
        for i in range(num_processes):
            Uc_hatT[:, i * Np1:(i + 1) * Np1] = U_mpi[i]
        return Uc_hatT

    m.dealias_filter = dealias_filter
    m.transpose_Uc = transpose_Uc
    sys.modules["mpiFFT4py.cython.maths"] = m


_loaded = None


def load():
    """Return the imported reference package ``mpiFFT4py`` (numpy.fft backend, fake MPI)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    _apply_compat_patches()
    if _HERE not in sys.path:
        sys.path.insert(0, _HERE)  # fake mpi4py
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "mpiFFT4py")):
        # GPU box: only the prebuilt install exists
        if INSTALLED not in sys.path:
            sys.path.insert(1, INSTALLED)
        try:
            importlib.import_module("mpiFFT4py.cython.maths")
            how = "cython (baseline/_ref)"
        except Exception:  # noqa: BLE001
            _inject_maths_standin()
            how = "standin (baseline/_ref)"
        _loaded = importlib.import_module("mpiFFT4py")
        _loaded._refshim_maths = how
        return _loaded
    root = _build_scratch_copy()
    with open(os.path.join(root, ".built")) as f:
        how = f.read()
    if root not in sys.path:
        sys.path.insert(1, root)
    if not how.startswith("cython"):
        pkg = importlib.import_module("mpiFFT4py.cython") if os.path.exists(
            os.path.join(root, "mpiFFT4py", "cython", "__init__.py")) else None
        del pkg
        _inject_maths_standin()
    _loaded = importlib.import_module("mpiFFT4py")
    _loaded._refshim_maths = how.strip()
    return _loaded


def run_ranks(nranks, fn, *args, **kwargs):
    load()
    from mpi4py import MPI
    return MPI.run_ranks(nranks, fn, *args, **kwargs)
