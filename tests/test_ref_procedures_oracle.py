"""tests/ref_procedures.py (the reference's own test procedures restated for this engine) checked on the CPU: the
ORACLE stands in for the device -- ranks are threads, the C-ABI calls under every transform call of the classes are
answered by the oracle's all-ranks function (tests/fake_device.py), the serial functions by numpy.fft.  What this pins
down without a GPU: the procedures' data flow, slices and tolerances are satisfiable by an implementation that
reproduces the reference (the oracle is pinned to the unmodified reference's goldens), for every fixture parameter at 1,
2 and 4 ranks.  The device runs are tests/test_gpu_reference_procedures.py (P = 1) and tests/gpu_dist_worker.py (P > 1)."""
import threading

import numpy as np
import pytest

import cpu_engine
import fake_device
import mpifft4py_b200 as m
import ref_procedures as rp


class ThreadWorld(object):
    def __init__(self, P):
        self.P = P
        self.barrier = threading.Barrier(P)
        self.slots = [None] * P
        self.failed = []


class ThreadComm(object):
    """The communicator surface the classes and the procedures use, ranks being threads of one process."""

    def __init__(self, world, rank, members=None):
        self.world, self.members = world, members or list(range(world.P))
        self.wrank = rank

    def Get_size(self):
        return len(self.members)

    def Get_rank(self):
        return self.members.index(self.wrank)

    def allgather(self, obj):  # (of the whole world: sub-communicators only report size and rank in these tests)
        return self.allgather_world(obj)

    def allgather_world(self, obj):
        w = self.world
        w.slots[self.wrank] = obj
        w.barrier.wait(timeout=120)
        out = list(w.slots)
        w.barrier.wait(timeout=120)
        return out

    def bcast(self, obj, root=0):  # (works for sub-communicators too: every rank of the world is in one such call)
        return self.allgather_world(obj if self.Get_rank() == root else None)[self.members[root]]

    def barrier(self):
        self.world.barrier.wait(timeout=120)

    def reduce(self, value, op="SUM", root=0):
        vals = [v for r, v in enumerate(self.allgather_world(value)) if r in self.members]
        return None if self.Get_rank() != root else (min(vals) if op == "MIN" else max(vals) if op == "MAX" else sum(vals))

    def Bcast(self, buf, root=0):
        assert len(self.members) == self.world.P
        src = self.allgather_world(buf if self.wrank == root else None)[root]
        if self.wrank != root:
            buf[...] = src
        self.world.barrier.wait(timeout=120)  # root keeps its array untouched until everybody has copied

    def Split(self, color=0, key=0):
        colors = self.allgather_world(int(color))
        return ThreadComm(self.world, self.wrank, [r for r in range(self.world.P) if colors[r] == int(color)])


SERIAL = {"rfftn": np.fft.rfftn, "irfftn": np.fft.irfftn, "rfft2": np.fft.rfft2, "irfft2": np.fft.irfft2, "fftn": np.fft.fftn,
          "ifftn": np.fft.ifftn, "irfft": np.fft.irfft, "ifft": np.fft.ifft}


def numpy_serial(npfn):
    def f(a, b, axes=None, axis=None, **kw):
        b[...] = npfn(a, axes=axes) if axis is None else npfn(a, axis=axis)
        return b
    return f


@pytest.fixture(params=["oracle", "engine"])
def backend(request, monkeypatch):
    """What stands under the classes where there is no GPU.  "oracle": tests/fake_device.py answers the C-ABI calls of
    Transform._run with the oracle, numpy.fft stands behind the serial function names.  "engine": tests/cpu_engine.py --
    the product's own C-ABI layer, plan programs and kernel phase bodies built for the host, serial functions included."""
    if request.param == "oracle":
        fake_device.install(monkeypatch)
        for name, fn in SERIAL.items():
            monkeypatch.setattr(m, name, numpy_serial(fn))
        yield request.param
    else:
        cleanup = cpu_engine.install(monkeypatch)
        yield cleanup.calls
        cleanup()


@pytest.mark.parametrize("P", [1, 2, 4, 8])
def test_reference_procedures_hold(backend, P):
    if P == 1:
        assert rp.run_all(m.comm.COMM_SELF) == 2 * 6 + 2
        if backend != "oracle":
            assert backend["b200fft_exec_forward"] >= 14 and backend["b200fft_exec_strided"] > 0 and backend["b200fft_exec_r2c"] > 0
        return
    world = ThreadWorld(P)
    counts = [None] * P

    def rank_main(r):
        try:
            counts[r] = rp.run_all(ThreadComm(world, r))
        except BaseException as e:  # noqa: BLE001 - reported below; the barrier is broken so the other ranks stop too
            world.failed.append((r, repr(e)))
            world.barrier.abort()

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(P)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=600)
    real = [f for f in world.failed if "BrokenBarrierError" not in f[1]]
    assert not world.failed, real or world.failed
    expect = 2 * ((16 if P >= 4 else 4) + 2) + 2
    assert counts == [expect] * P
    if backend != "oracle":  # the engine's own entry points did the work: plans with connected transports, serial passes
        assert backend["b200fft_exec_forward"] >= expect * P and backend["b200fft_exec_inverse"] >= expect * P
        assert backend["b200fft_plan_p2p_connect"] >= (20 if P >= 4 else 8) * P
        assert backend["b200fft_exec_strided"] > 0 and backend["b200fft_exec_r2c"] > 0
        assert backend["b200fft_exec_c2r"] > 0 or P < 4   # irfftn: only the 'AlltoallN' parameters call it (:66-69)


def test_parameter_lists_are_the_reference_fixtures():
    """tests/test_FFT.py:24-31: 4 slab parameters, 12 pencil ones from four ranks on; :48, :54."""
    three_d, lines, c2cs = rp.params(4)
    assert len(three_d) == 16 and len(set(three_d)) == 16 and len(rp.params(2)[0]) == 4
    assert sorted(three_d) == sorted(["slabas", "slabad", "slabws", "slabwd", "pencilsys", "pencilsyd", "pencilnys", "pencilnyd",
                                      "pencilsxd", "pencilsxs", "pencilnxd", "pencilnxs", "pencilaxd", "pencilaxs", "pencilayd",
                                      "pencilays"])
    F = rp.make("pencilnyd", type("C", (), {"Get_size": lambda s: 4, "Get_rank": lambda s: 0,
                                            "Split": lambda s, c=0, k=0: type("S", (), {"Get_size": lambda t: 2, "Get_rank": lambda t: 0})()})())
    assert F.communication == "AlltoallN" and type(F).__name__ == "R2CY" and F.float is np.float64
    assert lines == ["lines", "lined"] and c2cs == ["c2cd", "c2cs"]
