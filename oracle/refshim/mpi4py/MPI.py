"""Thread-per-rank fake of the ``mpi4py.MPI`` names used by mpiFFT4py -- TEST INFRASTRUCTURE ONLY.

Surface (everything the reference calls):
  COMM_WORLD, COMM_SELF, IN_PLACE, C_FLOAT_COMPLEX, C_DOUBLE_COMPLEX, _typedict, Compute_dims,
  MIN, SUM; Comm.{Get_size, Get_rank, Split, Alltoall, Alltoallw, Sendrecv_replace, Scatter,
  Send, Recv, Bcast, reduce, barrier}; Datatype.Create_subarray(...).Commit().

Every rank is a Python thread started by :func:`run_ranks`; the calling thread's world rank
lives in a ``threading.local``.  Collectives are "post, barrier, read, barrier".
"""
import queue
import threading

import numpy as np

_tls = threading.local()
_TIMEOUT = 120.0

IN_PLACE = "IN_PLACE"
MIN = "MIN"
SUM = "SUM"
MAX = "MAX"


class Datatype(object):
    def __init__(self, name, npchar):
        self.name = name
        self.npchar = npchar

    def Create_subarray(self, sizes, subsizes, starts):
        return Subarray(self, tuple(int(s) for s in sizes), tuple(int(s) for s in subsizes),
                        tuple(int(s) for s in starts))


class Subarray(object):
    def __init__(self, base, sizes, subsizes, starts):
        self.base = base
        self.sizes = sizes
        self.subsizes = subsizes
        self.starts = starts

    def Commit(self):
        return self

    def Free(self):
        pass

    def view(self, arr):
        a = np.asarray(arr)
        assert a.size == int(np.prod(self.sizes)), (a.shape, self.sizes)
        a = a.reshape(self.sizes)
        return a[tuple(slice(s, s + n) for s, n in zip(self.starts, self.subsizes))]


C_FLOAT_COMPLEX = Datatype("C_FLOAT_COMPLEX", "F")
C_DOUBLE_COMPLEX = Datatype("C_DOUBLE_COMPLEX", "D")
FLOAT = Datatype("FLOAT", "f")
DOUBLE = Datatype("DOUBLE", "d")
_typedict = {"F": C_FLOAT_COMPLEX, "D": C_DOUBLE_COMPLEX, "f": FLOAT, "d": DOUBLE}


def Compute_dims(nnodes, dims):
    """MPI_Dims_create for the 2D case: most balanced factorisation, non-increasing order."""
    assert dims == 2
    best = (nnodes, 1)
    for a in range(1, int(nnodes ** 0.5) + 1):
        if nnodes % a == 0:
            best = (nnodes // a, a)
    return [best[0], best[1]]


def _buf(x):
    """mpi4py buffer spec -> ndarray (accepts ``arr`` or ``[arr, type]``)."""
    if isinstance(x, (list, tuple)):
        return x[0]
    return x


class _Shared(object):
    """State shared by all members of one communicator."""

    def __init__(self, members):
        self.members = list(members)
        self.size = len(self.members)
        self.barrier = threading.Barrier(self.size, timeout=_TIMEOUT)
        self.slots = [None] * self.size
        self.mail = {}
        self.lock = threading.Lock()
        self.split_result = {}

    def box(self, src, dst, tag):
        key = (src, dst, tag)
        with self.lock:
            if key not in self.mail:
                self.mail[key] = queue.Queue()
            return self.mail[key]


class Comm(object):
    def __init__(self, shared):
        self._s = shared

    # -- introspection ---------------------------------------------------------------------
    def Get_size(self):
        return self._s.size

    def Get_rank(self):
        return self._s.members.index(_tls.world_rank)

    size = property(Get_size)
    rank = property(Get_rank)

    def _sync(self):
        self._s.barrier.wait()

    def barrier(self):
        self._sync()

    Barrier = barrier

    # -- communicator management -----------------------------------------------------------
    def Split(self, color=0, key=0):
        s = self._s
        me = self.Get_rank()
        s.slots[me] = (int(color), int(key), _tls.world_rank)
        self._sync()
        if me == 0:
            groups = {}
            for c, k, w in s.slots:
                groups.setdefault(c, []).append((k, s.members.index(w), w))
            s.split_result = {c: _Shared([w for _, _, w in sorted(v)]) for c, v in groups.items()}
        self._sync()
        out = Comm(s.split_result[int(color)])
        self._sync()
        return out

    # -- collectives -----------------------------------------------------------------------
    def Alltoall(self, sendbuf, recvbuf):
        s = self._s
        me = self.Get_rank()
        r = _buf(recvbuf)
        if sendbuf is IN_PLACE or (isinstance(sendbuf, str) and sendbuf == IN_PLACE):
            snd = np.array(r, copy=True)
        else:
            snd = np.ascontiguousarray(_buf(sendbuf))
        assert r.flags["C_CONTIGUOUS"], "Alltoall recv buffer must be contiguous"
        s.slots[me] = snd.reshape(s.size, -1)
        self._sync()
        rv = r.reshape(s.size, -1)
        for i in range(s.size):
            rv[i] = s.slots[i][me]
        self._sync()

    def Alltoallw(self, sendspec, recvspec):
        s = self._s
        me = self.Get_rank()
        sbuf, _, stypes = sendspec
        rbuf, _, rtypes = recvspec
        s.slots[me] = [np.array(t.view(sbuf), copy=True) for t in stypes]
        self._sync()
        for i in range(s.size):
            rtypes[i].view(rbuf)[...] = s.slots[i][me]
        self._sync()

    def Scatter(self, sendbuf, recvbuf, root=0):
        s = self._s
        me = self.Get_rank()
        if me == root:
            s.slots[root] = np.array(_buf(sendbuf), copy=True).reshape(s.size, -1)
        self._sync()
        r = _buf(recvbuf)
        r.reshape(-1)[...] = s.slots[root][me]
        self._sync()

    def Bcast(self, buf, root=0):
        s = self._s
        me = self.Get_rank()
        b = _buf(buf)
        if me == root:
            s.slots[root] = np.array(b, copy=True)
        self._sync()
        if me != root:
            b[...] = s.slots[root]
        self._sync()

    def reduce(self, value, op=SUM, root=0):
        s = self._s
        me = self.Get_rank()
        s.slots[me] = value
        self._sync()
        out = None
        if me == root:
            vals = list(s.slots)
            if op == MIN:
                out = min(vals)
            elif op == MAX:
                out = max(vals)
            else:
                out = vals[0]
                for v in vals[1:]:
                    out = out + v
        self._sync()
        return out

    # -- point to point --------------------------------------------------------------------
    def Send(self, buf, dest=0, tag=0):
        self._s.box(self.Get_rank(), dest, tag).put(np.array(_buf(buf), copy=True))

    def Recv(self, buf, source=0, tag=0):
        data = self._s.box(source, self.Get_rank(), tag).get(timeout=_TIMEOUT)
        b = _buf(buf)
        b[...] = data.reshape(b.shape)

    def Sendrecv_replace(self, buf, dest=0, sendtag=0, source=0, recvtag=0):
        b = _buf(buf)
        me = self.Get_rank()
        self._s.box(me, dest, ("sr", sendtag)).put(np.array(b, copy=True))
        data = self._s.box(source, me, ("sr", recvtag)).get(timeout=_TIMEOUT)
        b[...] = data.reshape(b.shape)


class _SelfComm(Comm):
    """COMM_SELF: a size-1 communicator valid on whichever thread uses it."""

    def __init__(self):
        pass

    @property
    def _s(self):
        sh = getattr(_tls, "self_shared", None)
        if sh is None or sh.members != [_tls.world_rank]:
            sh = _Shared([_tls.world_rank])
            _tls.self_shared = sh
        return sh


class _WorldComm(Comm):
    """COMM_WORLD: resolves to the world of the calling thread (set by run_ranks)."""

    def __init__(self):
        pass

    @property
    def _s(self):
        return _tls.world_shared


COMM_WORLD = _WorldComm()
COMM_SELF = _SelfComm()


def run_ranks(nranks, fn, *args, **kwargs):
    """Run ``fn(*args, **kwargs)`` on ``nranks`` threads, one MPI rank each; return results by rank."""
    shared = _Shared(range(nranks))
    results = [None] * nranks
    errors = [None] * nranks

    def body(r):
        _tls.world_rank = r
        _tls.world_shared = shared
        try:
            results[r] = fn(*args, **kwargs)
        except BaseException as e:  # noqa: BLE001 - propagate to the caller below
            errors[r] = e
            shared.barrier.abort()

    threads = [threading.Thread(target=body, args=(r,), daemon=True) for r in range(nranks)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(_TIMEOUT * 4)
    real = [e for e in errors if e is not None and not isinstance(e, threading.BrokenBarrierError)]
    if real:
        raise real[0]
    for e in errors:
        if e is not None:
            raise e
    return results


# Main thread acts as rank 0 of a size-1 world so plain (non run_ranks) use works too.
_tls.world_rank = 0
_tls.world_shared = _Shared([0])
