"""On-device selection of the opt-in kernels and schedules -- the role FFTW's planner plays behind the
reference's ``planner_effort`` argument (``pyfftw_fft.py:26-203`` forwards it to ``pyfftw.builders``):
measure candidate configurations of a transform object on the device it runs on, check each against the
default configuration's result, and keep the fastest.

    import mpifft4py_b200 as m
    F = m.Slab_R2C(N, L, comm, "double")
    report = m.tune.autotune(F)            # times the candidates, installs the best one into F
    report = m.tune.autotune(F, dealias="3/2-rule", candidates=m.tune.CANDIDATES["patient"])

A candidate is ``(name, {plan attribute: value})``.  The plan attributes are read when a plan is created, so
installing a candidate drops the object's device plan and lets the next transform build it again.  Every rank of a
multi-rank object must call ``autotune`` collectively; the slowest rank's time decides and a candidate that is wrong
or fails on any rank is rejected on all of them (one ``comm.allgather`` per candidate).

The kernel-level switches round 1 left open were decided by measurement in round 2 and are no longer options
(profiles/r02_single/ab_single.txt); what remains to choose per machine and mesh is the exchange: transport,
pipeline depth and direction -- and, for single-rank slab plans, the intermediate layout.
"""
import numpy as np

from . import _lib

PLAN_ATTRS = ("transport", "exchange_chunks", "exchange_pipeline", "layout")

CANDIDATES = {
    "measure": [("default", {}), ("p2p_c2", {"transport": "p2p", "exchange_chunks": 2}),
                ("p2p_c4", {"transport": "p2p", "exchange_chunks": 4}), ("p2p_c8", {"transport": "p2p", "exchange_chunks": 8}),
                ("nccl_c1", {"transport": "nccl", "exchange_chunks": 1}), ("nccl_c2", {"transport": "nccl", "exchange_chunks": 2})],
}
CANDIDATES["patient"] = CANDIDATES["measure"] + [
    ("p2p_kz2", {"transport": "p2p", "exchange_pipeline": "kz", "exchange_chunks": 2}),
    ("p2p_kz4", {"transport": "p2p", "exchange_pipeline": "kz", "exchange_chunks": 4}),
    ("store_c1", {"transport": "store", "exchange_chunks": 1}),
    ("natural_layout", {"layout": "natural"}),
]


def _install(F, attrs, base=None):
    """Make ``F`` use a candidate: the caller's own plan attributes (``base``) overlaid with the candidate's, and a
    fresh device plan at the next transform."""
    base = base or {}
    for a in PLAN_ATTRS:
        if a in attrs:
            setattr(F, a, attrs[a])
        elif a in base:
            setattr(F, a, base[a])
        elif a in getattr(F, "__dict__", {}):
            delattr(F, a)
    if getattr(F, "_plan", None) is not None:
        _lib.lib().b200fft_plan_destroy(F._plan)
        F._plan = None


def select(candidates, measure, tol, gather=None):
    """Core of the tuner, free of device code: ``measure(candidate) -> (seconds, error vs the default result)``
    (may raise: the candidate is then skipped and reported); a candidate qualifies when its error is within
    ``tol`` ON EVERY RANK; the fastest qualifying one (slowest rank's time) wins, ties and failures fall back to
    the first candidate (the default).  ``gather(x) -> [x of every rank]`` is called exactly once per candidate,
    also when ``measure`` raised on this rank, so the ranks of a multi-rank object stay in step and a candidate
    that is wrong or fails on any one rank is rejected everywhere.  Returns ``(best, report)``."""
    gather = gather or (lambda x: [x])
    report, best, best_t = [], None, None
    for cand in candidates:
        name = cand[0]
        mine = (None, None, None)
        try:
            t, err = measure(cand)
            mine = (float(t), float(err), None)
        except Exception as e:  # noqa: BLE001 - an opt-in kernel that fails must not take the application down
            mine = (None, None, repr(e)[:200])
        everyone = gather(mine)
        failed = [x[2] for x in everyone if x[2] is not None]
        if failed:
            report.append({"name": name, "seconds": None, "error": None, "ok": False, "exception": failed[0]})
            continue
        t = max(x[0] for x in everyone)
        err = max(x[1] for x in everyone)
        ok = bool(err <= tol)
        report.append({"name": name, "seconds": t, "error": err, "ok": ok})
        if ok and (best_t is None or t < best_t):
            best, best_t = cand, t
    if best is None:
        best = candidates[0]
    return best, report


def autotune(F, dealias=None, candidates=None, reps=3, tol=None):
    """Time ``fftn`` + ``ifftn`` (``fft2`` + ``ifft2`` for line objects) of ``F`` for every candidate on random
    data, keep the fastest whose forward result matches the default configuration's, install it into ``F`` and
    return the report.  Device memory: one real and one complex array of the object's local shapes, twice."""
    import torch
    candidates = list(candidates if candidates is not None else CANDIDATES["measure"])
    double = F.float is np.float64
    tol = tol if tol is not None else (1e-12 if double else 1e-5)
    fwd, inv = (F.fft2, F.ifft2) if hasattr(F, "fft2") else (F.fftn, F.ifftn)
    rshape = tuple(int(s) for s in (F.real_shape_padded() if dealias == "3/2-rule" else F.real_shape()))
    cshape = tuple(int(s) for s in F.complex_shape())
    rdt, cdt = (torch.float64, torch.complex128) if double else (torch.float32, torch.complex64)
    u = torch.rand(rshape, dtype=rdt, device="cuda")
    fu = torch.empty(cshape, dtype=cdt, device="cuda")
    u2 = torch.empty_like(u)
    ref = {}
    comm = getattr(F, "comm", None)
    many = comm is not None and getattr(F, "num_processes", 1) > 1
    if not many:  # exchange options mean nothing to a single rank
        candidates = [c for c in candidates if not (set(c[1]) - {"layout"})] or [("default", {})]
    base = {a: F.__dict__[a] for a in PLAN_ATTRS if a in F.__dict__}  # what the caller chose stays unless a candidate overrides it

    def measure(cand):
        _install(F, cand[1], base)
        fwd(u, fu, dealias)
        inv(fu, u2, dealias)  # warm-up: plan creation, kernel set-up
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fwd(u, fu, dealias)
            inv(fu, u2, dealias)
        e1.record()
        torch.cuda.synchronize()
        fwd(u, fu, dealias)
        if "fu" not in ref:
            ref["fu"] = fu.clone()
        err = float(torch.linalg.vector_norm(fu - ref["fu"]) / torch.linalg.vector_norm(ref["fu"]))
        return e0.elapsed_time(e1) * 1e-3 / reps, err

    # (every rank sees the same gathered (time, error, failure) triples, so every rank picks the same candidate)
    best, report = select(candidates, measure, tol, comm.allgather if many else None)
    _install(F, best[1], base)
    return {"chosen": best[0], "candidates": report}
