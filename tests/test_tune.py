"""Selection logic of mpifft4py_b200.tune (the device-free core: measure -> qualify -> pick)."""
import mpifft4py_b200 as m
from mpifft4py_b200 import tune


def test_fastest_qualifying_candidate_wins_and_failures_are_reported():
    cands = [("default", {}), ("fast_but_wrong", {}), ("fast", {}), ("crashes", {}), ("slow", {})]
    times = {"default": (1.0, 0.0), "fast_but_wrong": (0.1, 1e-3), "fast": (0.5, 1e-15), "slow": (2.0, 0.0)}

    def measure(c):
        if c[0] == "crashes":
            raise RuntimeError("launch failed")
        return times[c[0]]

    best, report = tune.select(cands, measure, 1e-12)
    assert best[0] == "fast"
    by = {r["name"]: r for r in report}
    assert by["fast_but_wrong"]["ok"] is False and by["crashes"]["ok"] is False and "launch failed" in by["crashes"]["exception"]
    assert by["fast"]["ok"] and by["default"]["ok"]


def test_default_is_kept_when_nothing_qualifies_and_slowest_rank_decides():
    cands = [("default", {}), ("a", {})]
    best, _ = tune.select(cands, lambda c: (1.0, 1.0), 1e-12)
    assert best[0] == "default"
    # gather models the allgather over ranks: candidate "a" is fast here but slow on another rank
    best, rep = tune.select(cands, lambda c: ((1.0, 0.0) if c[0] == "default" else (0.5, 0.0)), 1e-12,
                            gather=lambda x: [x, (x[0] if x[0] == 1.0 else 3.0, x[1], None)])
    assert best[0] == "default" and rep[1]["seconds"] == 3.0


def test_a_candidate_wrong_or_failing_on_another_rank_is_rejected_and_the_gather_is_always_called():
    cands = [("default", {}), ("wrong_elsewhere", {}), ("fails_elsewhere", {}), ("fails_here", {})]
    calls = []

    def measure(c):
        if c[0] == "fails_here":
            raise RuntimeError("boom")
        return (1.0, 0.0) if c[0] == "default" else (0.1, 0.0)

    def gather(x):
        calls.append(x)
        other = {1: (0.1, 1e-3, None), 2: (None, None, "RuntimeError('peer')")}.get(len(calls) - 1, x)
        return [x, other]

    best, rep = tune.select(cands, measure, 1e-12, gather=gather)
    assert best[0] == "default" and len(calls) == len(cands)  # one collective per candidate, failures included
    assert [r["ok"] for r in rep] == [True, False, False, False]
    assert rep[1]["error"] == 1e-3 and "peer" in rep[2]["exception"] and "boom" in rep[3]["exception"]


def test_install_sets_and_clears_plan_attributes():
    class F(object):
        _plan = None
    f = F()
    tune._install(f, {"exchange_chunks": 4, "transport": "p2p"})
    assert (f.exchange_chunks, f.transport) == (4, "p2p")
    tune._install(f, {"layout": "natural"}, base={"transport": "nccl"})
    assert not hasattr(f, "exchange_chunks") and f.layout == "natural" and f.transport == "nccl"
    assert set(c[0] for c in tune.CANDIDATES["measure"]) <= set(c[0] for c in tune.CANDIDATES["patient"])
    assert m.tune is tune
