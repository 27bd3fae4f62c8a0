"""Run a program written for mpiFFT4py without editing it.

    import mpifft4py_b200.compat
    mpifft4py_b200.compat.install()          # before the program's own imports

    from mpiFFT4py import Slab_R2C, rfftn    # ... now resolve to this package
    from mpiFFT4py.pencil import R2C
    from mpi4py import MPI                   # the real mpi4py if it is installed, else the stand-in below

``install()`` registers this package's modules under the names the reference is imported by
(``mpiFFT4py/__init__.py:1-8``: ``Slab_R2C``, ``Pencil_R2C``, ``Line_R2C``, ``work_arrays``, ``datatypes``, ``empty``,
``zeros``, the serial transform functions, ``fftfreq`` / ``rfftfreq``; the submodules ``slab``, ``pencil``, ``line``,
``mpibase``, ``serialFFT``).  A real ``mpi4py`` is left alone -- its communicators work with the classes as they are
(they need ``Get_size`` / ``Get_rank`` / ``Split`` for the bookkeeping and ``bcast`` / ``allgather`` to set the device
exchanges up).  Where ``mpi4py`` is not installed, a stand-in ``mpi4py.MPI`` provides what callers of the reference use
(its tests and demos: ``COMM_WORLD``, ``COMM_SELF``, ``MIN`` / ``MAX`` / ``SUM``, ``Compute_dims``): ``COMM_WORLD`` is
the ``torch.distributed`` world once that is initialised (one process per GPU under ``torchrun``), a single-process
communicator before.  tests/test_reference_suite_unmodified.py runs the reference's own test file through this.

As a launcher -- where upstream has ``mpirun -np P python script.py``:

    python -m mpifft4py_b200.compat script.py [args ...]                                            # one GPU
    python -m torch.distributed.run --nproc-per-node P -m mpifft4py_b200.compat script.py [args ...]

installs the aliases, binds the process to GPU ``LOCAL_RANK`` and initialises ``torch.distributed`` when launched with
several ranks, then runs the script as ``__main__``.
"""
import importlib
import sys
import types

from . import comm as _comm

_NAMES = ("mpiFFT4py", "mpiFFT4py.slab", "mpiFFT4py.pencil", "mpiFFT4py.line", "mpiFFT4py.mpibase", "mpiFFT4py.serialFFT")


class _World(object):
    """``MPI.COMM_WORLD``: resolved at every use, so that it follows ``torch.distributed.init_process_group``."""

    def __getattr__(self, name):
        return getattr(_comm.world(), name)

    def __repr__(self):
        return "<COMM_WORLD of mpifft4py_b200.compat: %d rank(s)>" % _comm.world().Get_size()


def compute_dims(nnodes, ndims):
    """``MPI.Compute_dims`` for the two-dimensional grids of ``pencil.py:185``: the most balanced factorisation,
    larger factor first."""
    assert int(ndims) == 2, "the reference asks for two-dimensional grids only"
    n = int(nnodes)
    b = max(d for d in range(1, int(n ** 0.5) + 1) if n % d == 0)
    return [n // b, b]


def mpi_stand_in():
    """A module object to stand where ``mpi4py`` would be imported from."""
    mpi = types.ModuleType("mpi4py")
    MPI = types.ModuleType("mpi4py.MPI")
    MPI.COMM_WORLD, MPI.COMM_SELF = _World(), _comm.COMM_SELF
    MPI.SUM, MPI.MIN, MPI.MAX = _comm.SUM, _comm.MIN, _comm.MAX
    MPI.Compute_dims = compute_dims
    mpi.MPI = MPI
    mpi.__doc__ = MPI.__doc__ = "stand-in of mpifft4py_b200.compat (mpi4py is not installed)"
    return mpi, MPI


def install(mpi4py="auto"):
    """Register the aliases.  ``mpi4py``: "auto" (stand-in only if the real package cannot be imported), True (always
    the stand-in) or False (never).  Returns the ``mpiFFT4py`` module object.  ``uninstall()`` removes them again."""
    import mpifft4py_b200 as m
    pkg = types.ModuleType("mpiFFT4py")
    pkg.__path__ = []  # a package: `from mpiFFT4py.slab import R2C` looks the submodule up in sys.modules
    pkg.__doc__ = "mpifft4py_b200 under the reference's name (mpifft4py_b200.compat.install)"
    for name in ("Slab_R2C", "Pencil_R2C", "Line_R2C", "work_arrays", "datatypes", "empty", "zeros", "fftfreq", "rfftfreq",
                 "__version__") + tuple(m.serialFFT.__all__):
        setattr(pkg, name, getattr(m, name))
    subs = {"slab": m.slab, "pencil": m.pencil, "line": m.line, "mpibase": m.mpibase, "serialFFT": m.serialFFT}
    sys.modules["mpiFFT4py"] = pkg
    for name, mod in subs.items():
        setattr(pkg, name, mod)
        sys.modules["mpiFFT4py." + name] = mod
    use_stand_in = bool(mpi4py)
    if mpi4py == "auto":
        try:
            importlib.import_module("mpi4py.MPI")
            use_stand_in = False
        except ImportError:
            use_stand_in = True
    if use_stand_in:
        mpi, MPI = mpi_stand_in()
        sys.modules["mpi4py"], sys.modules["mpi4py.MPI"] = mpi, MPI
    return pkg


def uninstall():
    for name in _NAMES:
        sys.modules.pop(name, None)
    for name in ("mpi4py", "mpi4py.MPI"):
        mod = sys.modules.get(name)
        if mod is not None and "stand-in of mpifft4py_b200.compat" in (getattr(mod, "__doc__", None) or ""):
            del sys.modules[name]


def main(argv=None):
    """``python -m mpifft4py_b200.compat script.py [args ...]`` (see the module docstring)."""
    import os
    import runpy
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] in ("-h", "--help"):
        print(__doc__)
        return 0 if argv else 2
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch
        import torch.distributed as dist
        cuda = torch.cuda.is_available()
        if cuda:
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        if not dist.is_initialized():
            dist.init_process_group(os.environ.get("B200FFT_DIST_BACKEND", "nccl" if cuda else "gloo"))
    install()
    sys.argv = argv
    try:
        runpy.run_path(argv[0], run_name="__main__")
    finally:
        if world > 1:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
