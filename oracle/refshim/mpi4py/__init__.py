"""In-process stand-in for ``mpi4py`` -- TEST INFRASTRUCTURE ONLY.

The reference (spectralDNS/mpiFFT4py) needs an MPI library through mpi4py; neither exists
in the build container.  This package implements exactly the MPI surface the reference
touches with one *thread* per rank and memcpy collectives, so that the
UNMODIFIED reference sources can be imported and run to produce golden vectors
(``oracle/refshim/make_golden.py``).  It is never imported by the product package.
"""
from . import MPI  # noqa: F401
