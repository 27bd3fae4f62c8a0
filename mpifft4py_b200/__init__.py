"""mpifft4py_b200 -- B200-native drop-in for mpiFFT4py's R2C transforms.

Same names as ``mpiFFT4py/__init__.py:1-8``: ``Slab_R2C``, ``Pencil_R2C``, ``Line_R2C``,
``work_arrays``, ``datatypes``, ``empty``, ``zeros`` and the serial transform functions, all
running as hand-written sm_100a kernels behind the C ABI in ``include/b200fft.h``.
"""
from .slab import R2C as Slab_R2C
from .slab import C2C as Slab_C2C
from .pencil import R2C as Pencil_R2C
from .line import R2C as Line_R2C
from .mpibase import work_arrays, datatypes, empty, zeros
from .serialFFT import dct, fft, ifft, rfft, irfft, rfft2, irfft2, rfftn, irfftn, fft2, ifft2, fftn, ifftn
from numpy.fft import fftfreq, rfftfreq
from . import comm
from . import device  # device-resident mesh / wavenumber / mask helpers and work arrays
from . import tune    # on-device selection of exchange transport / pipeline depth (the planner_effort of this engine)
from . import ns      # Navier-Stokes right-hand side around the transforms (elementwise kernels of the library)

__version__ = '0.1.0'
