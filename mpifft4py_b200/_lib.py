"""ctypes binding of ``libb200fft.so`` (include/b200fft.h).

There is no CPU fallback: if the CUDA library is missing or a call fails, this raises."""
import ctypes as C
import os

from . import _cdefs as D

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200fft.so")

_lib = None

SYMBOLS = [
    "b200fft_version", "b200fft_last_error", "b200fft_supported_length", "b200fft_copy",
    "b200fft_stream_sync", "b200fft_exec_strided", "b200fft_exec_r2c", "b200fft_exec_c2r",
    "b200fft_comm_unique_id", "b200fft_comm_create", "b200fft_comm_destroy",
    "b200fft_plan_create", "b200fft_plan_destroy", "b200fft_plan_workspace_bytes",
    "b200fft_exec_forward", "b200fft_exec_inverse", "b200fft_plan_last_launches",
    "b200fft_plan_set_timing", "b200fft_plan_last_phase_ms", "b200fft_plan_last_steps",
    "b200fft_plan_p2p_handles", "b200fft_plan_p2p_connect", "b200fft_ns_curl", "b200fft_ns_cross", "b200fft_ns_rhs",
]


class B200FFTError(RuntimeError):
    pass


def declare(L):
    """Argument / result types of include/b200fft.h on a loaded library."""
    L.b200fft_version.restype = C.c_int
    L.b200fft_last_error.restype = C.c_char_p
    L.b200fft_supported_length.argtypes = [C.c_int]
    L.b200fft_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.b200fft_stream_sync.argtypes = [C.c_void_p]
    L.b200fft_ns_curl.argtypes = [C.POINTER(D.NsMesh), C.c_void_p, C.c_void_p, C.c_void_p]
    L.b200fft_ns_cross.argtypes = [C.c_int, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.b200fft_ns_rhs.argtypes = [C.POINTER(D.NsMesh), C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_double, C.c_double, C.c_int, C.c_void_p]
    L.b200fft_exec_strided.argtypes = [C.POINTER(D.StridedDesc), C.c_void_p]
    L.b200fft_exec_r2c.argtypes = [C.POINTER(D.RowsDesc), C.c_void_p]
    L.b200fft_exec_c2r.argtypes = [C.POINTER(D.RowsDesc), C.c_void_p]
    L.b200fft_comm_unique_id.argtypes = [C.c_void_p]
    L.b200fft_comm_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_void_p]
    L.b200fft_comm_destroy.argtypes = [C.c_void_p]
    L.b200fft_plan_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(D.PlanDesc)]
    L.b200fft_plan_destroy.argtypes = [C.c_void_p]
    L.b200fft_plan_workspace_bytes.argtypes = [C.c_void_p]
    L.b200fft_plan_workspace_bytes.restype = C.c_size_t
    L.b200fft_exec_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.b200fft_exec_inverse.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.b200fft_plan_last_launches.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.b200fft_plan_set_timing.argtypes = [C.c_void_p, C.c_int]
    L.b200fft_plan_p2p_handles.argtypes = [C.c_void_p, C.c_void_p]
    L.b200fft_plan_p2p_connect.argtypes = [C.c_void_p, C.c_void_p]
    L.b200fft_plan_last_phase_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.b200fft_plan_last_steps.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                          C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int),
                                          C.POINTER(C.c_int)]


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200FFTError(
            "CUDA library %s is missing; build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C mpifft4py_b200/csrc` (there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    declare(L)
    for name in SYMBOLS:
        getattr(L, name)
    _lib = L
    return L


def check(rc):
    """Map C return codes to the exceptions the reference raises."""
    if rc == 0:
        return
    msg = lib().b200fft_last_error().decode("utf-8", "replace")
    if rc == D.ERR_ARG:
        raise AssertionError(msg)
    if rc == D.ERR_RANKS:
        raise IOError(msg)
    if rc == D.ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise B200FFTError("b200fft error %d: %s" % (rc, msg))
