"""CPU oracle for the mpiFFT4py R2C hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy restatement of the reference's distributed-transform algorithms (slab.R2C,
pencil.R2CX/R2CY, line.R2C) with the MPI ranks simulated as entries of Python lists and the
collectives as explicit block shuffles.  Every function cites the reference ``file:line`` it
follows (paths relative to ``/root/reference/``).  The 1D arithmetic is ``numpy.fft``
(pocketfft) exactly as in the reference's own fallback backend
``mpiFFT4py/serialFFT/numpy_fft.py:25-107``.

Parity status: PINNED.  ``tests/golden/*.npz`` hold outputs of the UNMODIFIED reference run in
the build container through ``oracle/refshim`` (fake mpi4py + Cython-built maths.pyx);
``tests/test_oracle_golden.py`` checks this oracle against every one of them, and
``tests/test_oracle_vs_reference.py`` re-runs the live reference when ``/root/reference`` exists.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this package.  The product package ``mpifft4py_b200`` never does.
"""
from . import slab, pencil, line  # noqa: F401
from .common import rel_l2, global_field  # noqa: F401
