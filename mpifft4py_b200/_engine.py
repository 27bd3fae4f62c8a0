"""Shared host-side machinery of the three transform classes: plan handle, argument checks,
host<->device staging for numpy callers, and the calls into ``libb200fft.so``.

Callers may pass numpy arrays (what every reference caller does -- they are staged through
device buffers owned by the object) or CUDA ``torch.Tensor``s (used in place, zero copy).
"""
import ctypes as C
import os

import numpy as np

from . import _cdefs as D
from . import _lib
from . import comm as _comm

_DEALIAS = {None: D.DEALIAS_NONE, "None": D.DEALIAS_NONE, "3/2-rule": D.DEALIAS_3_2, "2/3-rule": D.DEALIAS_2_3}


def _torch():
    import torch
    return torch


def _is_tensor(x):
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


class Transform(object):
    """Base of slab.R2C, pencil.R2CX/R2CY and line.R2C."""

    _plan = None

    def _create_plan(self, kind, N, nranks, rank, P1=1, P2=1, drop_nyquist=0, comm=None, comm0=None,
                     comm1=None):
        """Record the plan arguments.  The device plan (and, for several ranks, the NCCL
        communicators -- a collective step) is created on the first transform call, so that the
        integer bookkeeping of the class works without a GPU, exactly like the reference's lazy
        plan cache (``pyfftw_fft.py:28-29``)."""
        self._plan_args = (kind, [int(n) for n in N], nranks, rank, P1, P2, drop_nyquist, comm, comm0, comm1)
        self._plan = None
        self._stage = {}

    def _ensure_plan(self):
        if self._plan is not None:
            return
        kind, N, nranks, rank, P1, P2, drop_nyquist, comm, comm0, comm1 = self._plan_args
        torch = _torch()
        if not torch.cuda.is_available():
            raise _lib.B200FFTError("no CUDA device: mpifft4py_b200 has no CPU path")
        d = D.PlanDesc()
        d.kind = kind
        d.precision = D.DOUBLE if self.float is np.float64 else D.SINGLE
        for i, n in enumerate(N):
            d.N[i] = int(n)
        d.nranks, d.rank = int(nranks), int(rank)
        d.P1, d.P2 = int(P1), int(P2)
        d.padsize = float(self.padsize)
        d.drop_nyquist = int(drop_nyquist)
        d.transport = D.TRANSPORT_NCCL
        d.chunks = int(getattr(self, "exchange_chunks", 0) or os.environ.get("B200FFT_CHUNKS", "0"))
        # slab exchanges are cut by local x planes ("x": z, y | exchange, then x) or by kz ranges ("kz": z, then
        # y | exchange | x -- the exchange overlaps FFT passes on both sides); "auto" (default) lets the plan choose
        # per transform (include/b200fft.h: kz for large plain transforms over the copy engines)
        pipe = str(getattr(self, "exchange_pipeline", None) or os.environ.get("B200FFT_PIPELINE", "auto")).lower()
        assert pipe in ("auto", "x", "kz"), "exchange_pipeline must be 'auto', 'x' or 'kz'"
        d.pipeline = {"auto": D.PIPELINE_AUTO, "x": D.PIPELINE_X, "kz": D.PIPELINE_KZ}[pipe]
        # single-rank slab.R2C plans keep the array between the passes y-blocked (no pass with rows megabytes
        # apart); "natural" runs z, y, x on [x][y][kz] as the reference does (A/B measurements)
        lay = str(getattr(self, "layout", None) or os.environ.get("B200FFT_LAYOUT", "yblock")).lower()
        assert lay in ("yblock", "natural"), "layout must be 'yblock' or 'natural'"
        d.layout = D.LAYOUT_NATURAL if lay == "natural" else D.LAYOUT_YBLOCK
        # Exchanges default to the copy-engine (P2P) transport for every class: DMA pushes over NVLink that do not
        # occupy SMs, pipelined against the FFT passes (8 GPUs, profiles/r02_multi_8: slab 1024^3 5.4 ms against
        # 5.7 over NCCL send/recv, pencil X 5.9 against 7.7, pencil Y 2048^3 single 27.0 against 33.6).
        # B200FFT_TRANSPORT=nccl (or obj.transport = "nccl") selects the NCCL send/recv path; it is also what all
        # ranks agree to use if any of them cannot map its peers' buffers (no IPC / no peer access).
        # "store" is the fused transport: same peer mappings, but the FFT pass in front of an exchange stores each
        # peer's block straight into that peer's receive buffer over NVLink (one kernel does the FFT, the pack and
        # the transfer; measured slower than the copy engines on NVSwitch boxes because the storing pass then runs at
        # link speed -- 6.2 ms against 5.4 at 8 GPUs -- kept for machines without peer DMA overlap).
        choice = str(getattr(self, "transport", None) or os.environ.get("B200FFT_TRANSPORT") or "p2p").lower()
        assert choice in ("p2p", "nccl", "store"), "transport must be 'p2p', 'store' or 'nccl'"
        h = None
        if int(nranks) > 1 and choice in ("p2p", "store"):
            d.transport = D.TRANSPORT_P2P if choice == "p2p" else D.TRANSPORT_STORE
            d.comm = d.comm0 = d.comm1 = None
            h = C.c_void_p()
            L = _lib.lib()
            mine = C.create_string_buffer(256)
            ok = L.b200fft_plan_create(C.byref(h), C.byref(d)) == 0 and L.b200fft_plan_p2p_handles(h, mine) == 0
            err = "" if ok else L.b200fft_last_error().decode("utf-8", "replace")
            replies = comm.allgather((ok, mine.raw, err))
            if all(r[0] for r in replies):
                ok = L.b200fft_plan_p2p_connect(h, b"".join(r[1] for r in replies)) == 0
                err = "" if ok else L.b200fft_last_error().decode("utf-8", "replace")
            else:
                ok = False
                err = next(r[2] for r in replies if not r[0]) or err
            if not all(comm.allgather(ok)):
                if os.environ.get("B200FFT_STRICT_TRANSPORT") == "1":  # tests: never trade the asked-for transport silently
                    raise _lib.B200FFTError("peer-mapped transport '%s' unavailable: %s" % (choice, err))
                import warnings
                warnings.warn("mpifft4py_b200: copy-engine transport unavailable (%s); using NCCL send/recv" % err)
                if h:
                    L.b200fft_plan_destroy(h)
                h = None
        if h is None:
            d.transport = D.TRANSPORT_NCCL
            # (pencil plans exchange within comm0 / comm1 only; their world comm is for the peer-mapped handshake)
            d.comm = _comm.nccl_handle(comm) if (comm is not None and comm0 is None and comm1 is None) else None
            d.comm0 = _comm.nccl_handle(comm0) if comm0 is not None else None
            d.comm1 = _comm.nccl_handle(comm1) if comm1 is not None else None
            h = C.c_void_p()
            _lib.check(_lib.lib().b200fft_plan_create(C.byref(h), C.byref(d)))
        self.transport_used = {D.TRANSPORT_P2P: "p2p", D.TRANSPORT_STORE: "store"}.get(d.transport, "nccl")
        self._plan = h
        self._plan_desc = d
        self.device = torch.device("cuda", torch.cuda.current_device())

    def __del__(self):
        try:
            if self._plan is not None:
                _lib.lib().b200fft_plan_destroy(self._plan)
                self._plan = None
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass

    # ---------------------------------------------------------------- staging
    def _dev(self, tag, shape, dtype):
        torch = _torch()
        key = (tag, tuple(shape), np.dtype(dtype).str)
        t = self._stage.get(key)
        if t is None:
            tdt = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
                   np.dtype(np.complex64): torch.complex64, np.dtype(np.complex128): torch.complex128}[np.dtype(dtype)]
            t = torch.empty(tuple(int(s) for s in shape), dtype=tdt, device=self.device)
            self._stage[key] = t
        return t

    def _stream(self):
        return C.c_void_p(_torch().cuda.current_stream().cuda_stream)

    def _run(self, inverse, src, dst, dealias, src_shape, src_dtype, dst_shape, dst_dtype):
        assert dealias in ('3/2-rule', '2/3-rule', 'None', None)
        mode = _DEALIAS[dealias]
        self._ensure_plan()
        L = _lib.lib()
        st = self._stream()
        fn = L.b200fft_exec_inverse if inverse else L.b200fft_exec_forward
        src_shape = tuple(int(s) for s in src_shape)
        dst_shape = tuple(int(s) for s in dst_shape)
        assert tuple(src.shape) == src_shape, "input has shape %r, expected %r" % (tuple(src.shape), src_shape)
        assert tuple(dst.shape) == dst_shape, "output has shape %r, expected %r" % (tuple(dst.shape), dst_shape)
        if _is_tensor(src) or _is_tensor(dst):
            torch = _torch()
            assert _is_tensor(src) and _is_tensor(dst), "pass either numpy arrays or CUDA tensors, not a mix"
            assert src.device == self.device and dst.device == self.device, \
                "tensors must live on the transform's device (%s), not %s / %s" % (self.device, src.device, dst.device)
            assert src.is_contiguous() and dst.is_contiguous(), "tensors must be contiguous"
            want = {np.float32: torch.float32, np.float64: torch.float64, np.complex64: torch.complex64,
                    np.complex128: torch.complex128}
            assert src.dtype == want[src_dtype] and dst.dtype == want[dst_dtype], "wrong dtype for this precision"
            _lib.check(fn(self._plan, C.c_void_p(src.data_ptr()), C.c_void_p(dst.data_ptr()), mode, st))
            return dst
        # numpy path: H2D, transform, D2H (all ordered on the current stream)
        a = np.ascontiguousarray(src, dtype=src_dtype)
        out = dst if (isinstance(dst, np.ndarray) and dst.dtype == np.dtype(dst_dtype) and dst.flags["C_CONTIGUOUS"]
                      and dst.flags["WRITEABLE"]) else np.empty(dst_shape, dtype=dst_dtype)
        # two staging buffers serve both directions: what fftn stages its result in is what ifftn stages its input in
        # (same shape and type), and the other way round -- half the device memory of a buffer pair per direction
        dsrc = self._dev("ab"[inverse], src_shape, src_dtype)
        ddst = self._dev("ba"[inverse], dst_shape, dst_dtype)
        _lib.check(L.b200fft_copy(C.c_void_p(dsrc.data_ptr()), C.c_void_p(a.ctypes.data), a.nbytes, st))
        _lib.check(fn(self._plan, C.c_void_p(dsrc.data_ptr()), C.c_void_p(ddst.data_ptr()), mode, st))
        _lib.check(L.b200fft_copy(C.c_void_p(out.ctypes.data), C.c_void_p(ddst.data_ptr()), out.nbytes, st))
        _lib.check(L.b200fft_stream_sync(st))
        if out is not dst:
            dst[...] = out
        return dst

    # ---------------------------------------------------------------- introspection for bench / tests
    def workspace_bytes(self):
        self._ensure_plan()
        return int(_lib.lib().b200fft_plan_workspace_bytes(self._plan))

    def last_launches(self):
        k, x = C.c_int(), C.c_int()
        _lib.check(_lib.lib().b200fft_plan_last_launches(self._plan, C.byref(k), C.byref(x)))
        return k.value, x.value

    def set_timing(self, on=True):
        self._ensure_plan()
        _lib.check(_lib.lib().b200fft_plan_set_timing(self._plan, int(bool(on))))

    def last_phase_ms(self):
        f, x = C.c_float(), C.c_float()
        _lib.check(_lib.lib().b200fft_plan_last_phase_ms(self._plan, C.byref(f), C.byref(x)))
        return f.value, x.value

    def last_steps(self):
        """[(type, ms, algorithmic_bytes, length, pass)] of the last transform; type in
        {'c2c', 'r2c', 'c2r', 'exchange'}; ms < 0 unless set_timing(True); the chunks of a
        pipelined pass share `pass`."""
        M = 4096
        n = C.c_int()
        ty = (C.c_int * M)()
        ms = (C.c_float * M)()
        by = (C.c_double * M)()
        ln = (C.c_int * M)()
        ps = (C.c_int * M)()
        _lib.check(_lib.lib().b200fft_plan_last_steps(self._plan, M, C.byref(n), ty, ms, by, ln, ps))
        names = ["c2c", "r2c", "c2r", "exchange"]
        return [(names[ty[i]], float(ms[i]), float(by[i]), int(ln[i]), int(ps[i])) for i in range(min(n.value, M))]
