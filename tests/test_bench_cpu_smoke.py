"""bench.py end to end on the CPU: `main()` with a tiny workload on the host build of the engine (tests/cpu_engine.py),
torch's CUDA entry points replaced by host stand-ins (host tensors for device="cuda", wall-clock events).  The numbers
mean nothing; what is checked is that every leg of the script runs and that the one JSON line has the keys and shapes
the driver's contract asks for -- forward parity, per-pass roofline table, e2e through the numpy API, the extra
BASELINE workloads, the goldens of the unmodified reference -- before a GPU box ever sees it."""
import json
import sys
import time

import pytest
import torch

import bench
import cpu_engine


class Event(object):
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


class Stream(object):
    cuda_stream = 0


def host_cuda(monkeypatch):
    def on_host(fn):
        def f(*a, **kw):
            if str(kw.get("device", "")).startswith("cuda"):
                kw["device"] = "cpu"
            return fn(*a, **kw)
        return f
    for name in ("rand", "empty", "zeros", "ones", "tensor"):
        monkeypatch.setattr(torch, name, on_host(getattr(torch, name)))
    real_generator = torch.Generator
    monkeypatch.setattr(torch, "Generator", lambda device=None: real_generator())
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "empty_cache", lambda: None)
    monkeypatch.setattr(torch.cuda, "Event", Event)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a: Stream())


@pytest.mark.parametrize("workload", ["tiny", "tiny_32"])
def test_bench_main_on_the_host_build(monkeypatch, capsys, workload):
    cleanup = cpu_engine.install(monkeypatch)
    host_cuda(monkeypatch)
    tiny = {"tiny": ("slab", (16, 16, 32), "double", None, {}), "tiny_32": ("slab", (16, 16, 32), "double", "3/2-rule", {}),
            "tiny_f32": ("slab", (16, 32, 16), "single", None, {}), "tiny_line": ("line", (32, 64), "single", None, {})}
    monkeypatch.setattr(bench, "WORKLOADS", dict(bench.WORKLOADS, **tiny))
    monkeypatch.setattr(bench, "other_workload_names", lambda main, P: [n for n in ("tiny_32", "tiny_f32", "tiny_line") if n != main])
    monkeypatch.setattr(sys, "argv", ["bench.py", "--workload", workload, "--steps", "2", "--warmup", "3", "--others-steps", "2",
                                      "--no-cpu-baseline"])
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    try:
        assert bench.main() == 0
    finally:
        cleanup()
    lines = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "roofline", "clocks", "e2e", "gpu_launches", "cpu_baseline", "forward_rel_l2",
                "other_workloads", "reference_goldens", "workspace_bytes"):
        assert key in d, key
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["dtype"] == "f64" and d["unit"] == "GFLOP/s"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["scaling"] == "strong"
    assert d["config"]["workload"].startswith("slab.R2C N=16x16x32 double") and d["config"]["name"] == workload
    assert d["forward_rel_l2"] < 1e-13 and d["gpu_launches"] == 3 * 2 * 2
    r = d["roofline"]
    # (the emulator moves kilobytes per millisecond: `achieved` rounds to zero here)
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] > 0 and r["achieved"] >= 0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    assert r["algorithmic_bytes"] == 16 * 16 * (32 * 8 + 17 * 16) or r["algorithmic_bytes"] == 2 * 16 * 16 * 17 * 16 or workload != "tiny"
    assert len(r["passes"]) == 6 and {p["type"] for p in r["passes"]} == {"r2c", "c2c", "c2r"}
    assert all(p["bytes"] > 0 and p["ms"] > 0 for p in r["passes"])
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == e["d2h_bytes_per_step"] > 0 and e["steps"] == 2
    others = {o["name"]: o for o in d["other_workloads"]}
    assert sorted(others) == sorted(n for n in ("tiny_32", "tiny_f32", "tiny_line") if n != workload)
    for o in others.values():
        assert "error" not in o and o["ms_per_step"] > 0 and o["forward_rel_l2"] < (1e-13 if o["dtype"] == "f64" else 1e-5), o
        assert (o["roundtrip_rel_l2"] is None) == ("3/2-rule" in o["workload"])
    g = d["reference_goldens"]
    assert g["files"] == ["line_P1_d", "slab_P1_Alltoallw_d"] and g["max_rel_l2"]["double"] < 1e-13


# ------------------------------------------------------------------------------------------------------------------
# several ranks: threads of this process, each running bench.main() -- the environment, torch.distributed and the
# communicator are per-thread stand-ins; the engine underneath is the host build with its real peer-mapped transport
# ------------------------------------------------------------------------------------------------------------------
import os
import threading

from test_ref_procedures_oracle import ThreadComm, ThreadWorld


class ThreadEnv(object):
    """os.environ as bench.py reads it, per thread (RANK / LOCAL_RANK / WORLD_SIZE differ between the ranks)."""

    def __init__(self):
        self.local = threading.local()

    def _d(self):
        return getattr(self.local, "env", {})

    def get(self, k, default=None):
        return self._d().get(k, os.environ.get(k, default))

    def __getitem__(self, k):
        return self._d()[k] if k in self._d() else os.environ[k]

    def __contains__(self, k):
        return k in self._d() or k in os.environ


class OsProxy(object):
    def __init__(self, env):
        self.environ = env

    def __getattr__(self, name):
        return getattr(os, name)


@pytest.mark.parametrize("P", [2, 4])
def test_bench_main_multi_rank(monkeypatch, capsys, P):
    import torch.distributed as dist
    import mpifft4py_b200.comm as comm_mod
    cleanup = cpu_engine.install(monkeypatch)
    host_cuda(monkeypatch)
    tw = ThreadWorld(P)
    me = threading.local()
    env = ThreadEnv()

    def all_reduce(t, op=None):
        vals = me.comm.allgather_world(t.clone())
        t.copy_(torch.stack(vals).max(0).values)

    monkeypatch.setattr(dist, "init_process_group", lambda *a, **k: None)
    monkeypatch.setattr(dist, "destroy_process_group", lambda *a, **k: None)
    monkeypatch.setattr(dist, "barrier", lambda *a, **k: me.comm.barrier())
    monkeypatch.setattr(dist, "all_reduce", all_reduce)
    monkeypatch.setattr(comm_mod, "world", lambda: me.comm)
    monkeypatch.setattr(bench, "os", OsProxy(env))
    tiny = {"tiny": ("slab", (16, 16, 32), "double", None, {}), "tiny_32": ("slab", (16, 16, 32), "double", "3/2-rule", {}),
            "tiny_line": ("line", (32, 64), "single", None, {}),
            "tiny_pencil": ("pencil", (16, 32, 32), "double", None, dict(alignment="X", P1=None, communication="Alltoallw")),
            "tiny_pencilY": ("pencil", (16, 32, 32), "single", None, dict(alignment="Y", P1=None, communication="Alltoallw"))}
    monkeypatch.setattr(bench, "WORKLOADS", dict(bench.WORKLOADS, **tiny))
    extra = ["tiny_32", "tiny_line"] + (["tiny_pencil", "tiny_pencilY"] if P >= 4 else [])
    monkeypatch.setattr(bench, "other_workload_names", lambda main, n: list(extra))
    monkeypatch.setattr(sys, "argv", ["bench.py", "--gpus", str(P), "--workload", "tiny", "--steps", "2", "--warmup", "3",
                                      "--others-steps", "2"])
    rcs = [None] * P

    def rank_main(r):
        me.comm = ThreadComm(tw, r)
        env.local.env = {"WORLD_SIZE": str(P), "RANK": str(r), "LOCAL_RANK": str(r)}
        try:
            rcs[r] = bench.main()
        except BaseException as e:  # noqa: BLE001
            tw.failed.append((r, repr(e)[:400]))
            tw.barrier.abort()

    ts = [threading.Thread(target=rank_main, args=(r,)) for r in range(P)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=600)
    cleanup()
    real = [f for f in tw.failed if "BrokenBarrierError" not in f[1]]
    assert not tw.failed, real or tw.failed
    assert rcs == [0] * P
    lines = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, lines   # rank 0 alone prints
    d = json.loads(lines[0])
    assert d["n_gpus"] == P and d["forward_rel_l2"] < 1e-13 and d["cpu_baseline"] is None
    assert d["config"]["exchange"]["transport"] == "p2p" and d["config"]["decomposition"] == "slab P=%d" % P
    assert d["nccl_groups_per_transform"] >= 1 and d["roofline"]["sum_exchange_ms"] > 0
    assert {p["type"] for p in d["roofline"]["passes"]} == {"r2c", "c2c", "c2r", "exchange"}
    assert d["e2e"]["value"] > 0
    others = {o["name"]: o for o in d["other_workloads"]}
    assert sorted(others) == sorted(extra)
    for o in others.values():
        assert "error" not in o, o
        assert o["transport"] == "p2p" and o["forward_rel_l2"] < (1e-13 if o["dtype"] == "f64" else 1e-5), o
    g = d["reference_goldens"]
    assert len(g["files"]) == {2: 3, 4: 10}[P] and g["max_rel_l2"]["double"] < 1e-13 and g["max_rel_l2"]["single"] < 2e-6
