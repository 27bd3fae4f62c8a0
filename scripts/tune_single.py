#!/usr/bin/env python
"""Run the on-device tuner (mpifft4py_b200.tune.autotune) for a bench workload on one GPU and print its report.

    python scripts/tune_single.py slab1024_f64 [measure|patient]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import mpifft4py_b200 as m  # noqa: E402
from mpifft4py_b200.comm import SelfComm  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "slab1024_f64"
effort = sys.argv[2] if len(sys.argv) > 2 else "measure"
F = bench.make_transform(m, SelfComm(), name)
rep = m.tune.autotune(F, dealias=bench.WORKLOADS[name][3], candidates=m.tune.CANDIDATES[effort])
print(json.dumps(rep, indent=1))
