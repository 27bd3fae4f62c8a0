"""ctypes mirrors of the structs in ``include/b200fft.h`` (the C-ABI drop-in boundary)."""
import ctypes as C

MAXP = 16
SINGLE, DOUBLE = 0, 1
SLAB, PENCIL_X, PENCIL_Y, LINE, SLAB_C2C = 0, 1, 2, 3, 4
DEALIAS_NONE, DEALIAS_3_2, DEALIAS_2_3 = 0, 1, 2
TRANSPORT_NCCL, TRANSPORT_P2P, TRANSPORT_STORE = 0, 1, 2
PIPELINE_AUTO, PIPELINE_X, PIPELINE_KZ = 0, 1, 2
LAYOUT_YBLOCK, LAYOUT_NATURAL = 0, 1

ERR_ARG, ERR_RANKS, ERR_UNSUPPORTED, ERR_CUDA, ERR_NCCL, ERR_NOMEM = 1, 2, 3, 4, 5, 6


class Side(C.Structure):
    _fields_ = [("base", C.c_void_p * MAXP), ("sb", C.c_longlong * MAXP), ("si", C.c_longlong * MAXP),
                ("chunk", C.c_int), ("nchunk", C.c_int), ("nphys", C.c_int)]


class Mask(C.Structure):
    _fields_ = [("on", C.c_int),
                ("i_off", C.c_int), ("i_lo", C.c_int), ("i_hi", C.c_int),
                ("b_off", C.c_int), ("b_lo", C.c_int), ("b_hi", C.c_int),
                ("jdiv", C.c_int),
                ("jq_off", C.c_int), ("jq_lo", C.c_int), ("jq_hi", C.c_int),
                ("jr_off", C.c_int), ("jr_lo", C.c_int), ("jr_hi", C.c_int)]


class StridedDesc(C.Structure):
    _fields_ = [("precision", C.c_int), ("n", C.c_int), ("B", C.c_longlong), ("J", C.c_int),
                ("inverse", C.c_int), ("fold_mode", C.c_int), ("scale", C.c_double),
                ("inp", Side), ("out", Side), ("mask", Mask), ("cross_n", C.c_int), ("cross_div", C.c_int)]


class RowsDesc(C.Structure):
    _fields_ = [("precision", C.c_int), ("n", C.c_int), ("rows", C.c_longlong), ("nk", C.c_int),
                ("scale", C.c_double), ("real_base", C.c_void_p), ("rpitch", C.c_longlong),
                ("cside", Side), ("rm_period", C.c_longlong), ("rm_block", C.c_longlong), ("rm_planes", C.c_longlong)]


class PlanDesc(C.Structure):
    _fields_ = [("kind", C.c_int), ("precision", C.c_int), ("N", C.c_longlong * 3),
                ("nranks", C.c_int), ("rank", C.c_int), ("P1", C.c_int), ("P2", C.c_int),
                ("padsize", C.c_double), ("drop_nyquist", C.c_int), ("transport", C.c_int),
                ("comm", C.c_void_p), ("comm0", C.c_void_p), ("comm1", C.c_void_p), ("chunks", C.c_int), ("pipeline", C.c_int),
                ("layout", C.c_int)]


class NsMesh(C.Structure):
    _fields_ = [("precision", C.c_int), ("n0", C.c_longlong), ("n1", C.c_longlong), ("n2", C.c_longlong),
                ("kx", C.c_void_p), ("ky", C.c_void_p), ("kz", C.c_void_p)]


def no_mask():
    m = Mask()
    m.on = 0
    m.jdiv = 1
    for f in ("i", "b", "jq", "jr"):
        setattr(m, f + "_lo", 1)
        setattr(m, f + "_hi", 0)
    return m


def plain_side(ptr, sb, si, nphys):
    s = Side()
    s.base[0] = ptr
    s.sb[0] = sb
    s.si[0] = si
    s.chunk = nphys
    s.nchunk = 1
    s.nphys = nphys
    return s


def chunked_side(ptrs, sbs, sis, chunk, nphys):
    s = Side()
    for p, (ptr, sb, si) in enumerate(zip(ptrs, sbs, sis)):
        s.base[p] = ptr
        s.sb[p] = sb
        s.si[p] = si
    s.chunk = chunk
    s.nchunk = len(ptrs)
    s.nphys = nphys
    return s
