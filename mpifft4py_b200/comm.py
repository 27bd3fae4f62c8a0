"""Communicators for the SPMD, one-process-per-GPU model.

The reference takes an ``mpi4py.MPI.Comm`` (``slab.py:67``, ``pencil.py:167``, ``line.py:55``).
mpi4py is optional here: any object with ``Get_size``/``Get_rank`` (and ``Split``, ``bcast`` for
multi-rank use) is accepted, and :class:`TorchComm` provides that surface on top of
``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests), including the ``Bcast``/``reduce``/
``barrier`` calls the reference's tests and demos make.  The data path never goes through these
objects: they only bootstrap the exchanges of ``libb200fft.so`` (the CUDA IPC handles of the copy-engine transport, the
unique ids of its NCCL communicators).
"""
import ctypes as C

import numpy as np

SUM, MIN, MAX = "SUM", "MIN", "MAX"


class SelfComm(object):
    """COMM_SELF / single-process COMM_WORLD."""

    def Get_size(self):
        return 1

    def Get_rank(self):
        return 0

    def Split(self, color=0, key=0):
        return SelfComm()

    def Bcast(self, buf, root=0):
        return None

    def bcast(self, obj, root=0):
        return obj

    def barrier(self):
        return None

    Barrier = barrier

    def reduce(self, value, op=SUM, root=0):
        return value

    def allgather(self, obj):
        return [obj]


class TorchComm(object):
    """``torch.distributed`` process group with the mpi4py method names the reference uses."""

    def __init__(self, group=None):
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised (launch with torchrun)")
        self._dist = dist
        self.group = group if group is not None else dist.group.WORLD
        self._ranks = dist.get_process_group_ranks(self.group)

    # -- introspection
    def Get_size(self):
        return len(self._ranks)

    def Get_rank(self):
        return self._ranks.index(self._dist.get_rank())

    def _device(self):
        import torch
        if self._dist.get_backend(self.group) == "nccl":
            return torch.device("cuda", torch.cuda.current_device())
        return torch.device("cpu")

    # -- object collectives (bootstrap, tests)
    def allgather(self, obj):
        out = [None] * self.Get_size()
        self._dist.all_gather_object(out, obj, group=self.group)
        return out

    def bcast(self, obj, root=0):
        box = [obj]
        self._dist.broadcast_object_list(box, src=self._ranks[root], group=self.group, device=self._device())
        return box[0]

    def Bcast(self, buf, root=0):
        import torch
        arr = buf[0] if isinstance(buf, (list, tuple)) else buf
        if isinstance(arr, torch.Tensor):
            self._dist.broadcast(arr, src=self._ranks[root], group=self.group)
            return
        flat = np.ascontiguousarray(arr).reshape(-1)
        if np.iscomplexobj(flat):  # as pairs of reals: not every backend broadcasts complex tensors
            flat = flat.view(flat.real.dtype)
        t = torch.from_numpy(flat).to(self._device())
        self._dist.broadcast(t, src=self._ranks[root], group=self.group)
        arr[...] = t.cpu().numpy().view(arr.dtype).reshape(arr.shape)

    def barrier(self):
        self._dist.barrier(group=self.group)

    Barrier = barrier

    def reduce(self, value, op=SUM, root=0):
        vals = self.allgather(value)
        if self.Get_rank() != root:
            return None
        if op == MIN:
            return min(vals)
        if op == MAX:
            return max(vals)
        out = vals[0]
        for v in vals[1:]:
            out = out + v
        return out

    # -- communicator management (pencil.py:192-193)
    def Split(self, color=0, key=0):
        """Collective over this communicator.  ``new_group`` is collective over the WORLD group in
        torch.distributed, so Split is only supported on communicators spanning all processes."""
        dist = self._dist
        if self.Get_size() != dist.get_world_size():
            raise NotImplementedError("Split of a sub-communicator")
        info = self.allgather((int(color), int(key), dist.get_rank()))
        groups = {}
        for c, k, r in info:
            groups.setdefault(c, []).append((k, r))
        # the same partition again (every pencil object splits the world the same two ways) returns the communicator
        # made the first time: no new process groups, no new NCCL communicators behind them
        part = tuple(tuple(r for _, r in sorted(groups[c])) for c in sorted(groups))
        mine = _split_cache.get(part)
        if mine is None:
            for ranks in part:
                g = dist.new_group(ranks=list(ranks))
                if dist.get_rank() in ranks:
                    mine = TorchComm(g)
            _split_cache[part] = mine
        return mine


_split_cache = {}


def world():
    """COMM_WORLD: the torch.distributed world if initialised, else a single-process communicator."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return TorchComm()
    except ImportError:
        pass
    return SelfComm()


COMM_SELF = SelfComm()

_nccl_cache = {}


def nccl_handle(comm):
    """``b200fft_comm_t`` for ``comm`` (None for a single rank): rank 0 draws the NCCL unique id,
    the communicator's own ``bcast`` carries it to the other ranks (out-of-band bootstrap)."""
    from . import _lib
    size, rank = comm.Get_size(), comm.Get_rank()
    if size == 1:
        return None
    key = id(comm)
    if key in _nccl_cache:
        return _nccl_cache[key][0]
    L = _lib.lib()
    buf = C.create_string_buffer(128)
    if rank == 0:
        _lib.check(L.b200fft_comm_unique_id(buf))
    ident = comm.bcast(bytes(buf.raw) if rank == 0 else None, root=0)
    handle = C.c_void_p()
    idbuf = C.create_string_buffer(ident, 128)
    _lib.check(L.b200fft_comm_create(C.byref(handle), size, rank, idbuf))
    _nccl_cache[key] = (handle, comm)  # keep comm alive so id() stays unique
    return handle
