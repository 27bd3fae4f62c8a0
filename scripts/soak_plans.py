#!/usr/bin/env python
"""Soak run of the plan programs in the CPU emulator against the oracle, beyond the seeds the test-suite pins
(tests/test_emu_random_plans.py): fresh seeds, AUTO pipeline, 16 ranks, larger meshes, single-rank layouts, line plans.

    python scripts/soak_plans.py [first_seed] [seeds]

One line per seed; a failing case is printed with its parameters (it can be replayed through the functions of
tests/test_emu_random_plans.py / tests/test_emu_plans.py).  Development tool; not part of the product path."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
import test_emu_plans as E  # noqa: E402
import test_emu_random_plans as T  # noqa: E402
from mpifft4py_b200 import _cdefs as D  # noqa: E402


def wide_slab_cases(count, seed):
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < count:
        N = tuple(int(rng.choice([8, 16, 24, 32, 48, 64, 96])) for _ in range(3))
        if np.prod(N) > 64 * 64 * 48:
            continue
        P = int(rng.choice([1, 2, 4, 8, 16]))
        if N[0] % P or N[1] % P or (P > 1 and P > N[0] // 2):
            continue
        transport = int(rng.choice([D.TRANSPORT_NCCL, D.TRANSPORT_P2P, D.TRANSPORT_STORE])) if P > 1 else D.TRANSPORT_NCCL
        pipeline = int(rng.choice([D.PIPELINE_AUTO, D.PIPELINE_X, D.PIPELINE_KZ]))
        chunks = int(rng.choice([0, 1, 2, 3, 4, 6, 8]))
        kind = str(rng.choice(["r2c", "r2c", "c2c"]))
        prec = "double" if rng.random() < 0.6 else "single"
        out.append((N, P, transport, pipeline, chunks, kind, prec))
    return out


def line_cases(count, seed):
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < count:
        N = (int(rng.choice([8, 16, 32, 48, 64, 96, 128])), int(rng.choice([16, 32, 64, 96, 128])))
        P = int(rng.choice([1, 2, 4, 8, 16]))
        if N[1] % (2 * P) or N[0] % P or not all(T._supported(3 * n // 2) for n in N) or (3 * N[0]) % (2 * P):
            continue
        transport = int(rng.choice([D.TRANSPORT_NCCL, D.TRANSPORT_P2P, D.TRANSPORT_STORE]))
        out.append((N, P, "double" if rng.random() < 0.6 else "single", transport))
    return out


def single_rank_cases(count, seed):
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < count:
        N = tuple(int(rng.choice([4, 8, 12, 16, 24, 32, 48, 64])) for _ in range(3))
        out.append((N, "double" if rng.random() < 0.6 else "single", int(rng.choice([D.LAYOUT_YBLOCK, D.LAYOUT_NATURAL]))))
    return out


def main():
    seed0 = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    fails, t0 = 0, time.time()

    def attempt(tag, fn, case):
        nonlocal fails
        try:
            fn(*case)
        except BaseException as e:  # noqa: BLE001 - includes pytest's Skipped (an illegal decomposition) and failures
            if type(e).__name__ == "Skipped":
                return
            fails += 1
            print("%s FAIL %r %s" % (tag, case, repr(e)[:300]), flush=True)

    for seed in range(seed0, seed0 + n):
        for case in wide_slab_cases(16, seed):
            attempt("SLAB", T.test_random_slab_plan, case)
        for case in T._pencil_cases(8, seed + 100000):
            attempt("PENCIL", T.test_random_pencil_plan, case)
        for case in line_cases(6, seed + 200000):
            attempt("LINE", E._line_body, case)
        for case in single_rank_cases(4, seed + 300000):
            attempt("SINGLE", E.test_slab_single_rank_layouts, case)
        print("seed %d done, failures so far %d, %.0f s" % (seed, fails, time.time() - t0), flush=True)
    print("TOTAL FAILURES", fails)
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
