#!/usr/bin/env python
"""A/B of plan options on ONE GPU in ONE process (one torch import, one set of
arrays): for every (workload, configuration) a fresh plan object, warm-up, K timed round trips with CUDA
events, the per-pass table and a parity check (round-trip rel. L2; forward result against the default
configuration's, which the parity suite pins to the oracle).  One JSON line per case on stdout, a table
on stderr.  Timings here guide the choice of defaults; the numbers that are reported come from bench.py.

    python scripts/ab_single.py [--steps 5] [--workloads slab1024_f64,slab1024_f64_32] [--only substr]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import mpifft4py_b200 as m  # noqa: E402
from mpifft4py_b200 import _lib  # noqa: E402
from mpifft4py_b200.comm import SelfComm  # noqa: E402

# name -> plan attributes
CONFIGS = [("yblock (default)", {}), ("natural", {"layout": "natural"})]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--workloads", default="slab1024_f64,slab1024_f64_32")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    L = _lib.lib()
    peak = 6554.9
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:  # noqa: BLE001
        pass
    for name in args.workloads.split(","):
        kind, N, prec, dealias, kw = bench.WORKLOADS[name]
        F0 = bench.make_transform(m, SelfComm(), name)
        rshape = tuple(int(s) for s in (F0.real_shape_padded() if dealias == "3/2-rule" else F0.real_shape()))
        cshape = tuple(int(s) for s in F0.complex_shape())
        rdt = torch.float64 if prec == "double" else torch.float32
        cdt = torch.complex128 if prec == "double" else torch.complex64
        u = torch.rand(rshape, dtype=rdt, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1234))
        fu = torch.empty(cshape, dtype=cdt, device="cuda")
        u2 = torch.empty_like(u)
        ref = None
        flops = bench.flops_roundtrip(tuple(int(1.5 * n) if dealias == "3/2-rule" else n for n in N))
        del F0
        for cname, attrs in CONFIGS:
            if args.only and args.only not in cname and not cname.endswith("(default)"):
                continue
            F = bench.make_transform(m, SelfComm(), name)
            for k, v in attrs.items():
                if not k.startswith("_"):
                    setattr(F, k, v)
            try:
                fwd_, inv_ = (F.fft2, F.ifft2) if kind == "line" else (F.fftn, F.ifftn)
                for _ in range(2):
                    fwd_(u, fu, dealias)
                    inv_(fu, u2, dealias)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.steps):
                    fwd_(u, fu, dealias)
                    inv_(fu, u2, dealias)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / args.steps
                rt = float(torch.linalg.vector_norm(u2 - u) / torch.linalg.vector_norm(u))
                fwd_(u, fu, dealias)
                if ref is None:
                    ref = fu.clone()
                dev = float(torch.linalg.vector_norm(fu - ref) / torch.linalg.vector_norm(ref))
                F.set_timing(True)
                acc = {}
                for _ in range(2):
                    for direction, call in (("fwd", lambda: fwd_(u, fu, dealias)), ("inv", lambda: inv_(fu, u2, dealias))):
                        call()
                        torch.cuda.synchronize()
                        for s in F.last_steps():
                            a = acc.setdefault((direction, s[4], s[0]), [0.0, 0.0])
                            a[0] += s[1] / 2
                            a[1] = max(a[1], 0) + s[2] / 2
                passes = [{"dir": d, "pass": i, "type": t, "ms": round(v[0], 3), "GBps": round(v[1] / v[0] / 1e6, 0) if v[0] > 0 else None}
                          for (d, i, t), v in sorted(acc.items())]
                k, _ = F.last_launches()
                out = {"workload": name, "config": cname, "attrs": attrs, "ms_per_round_trip": round(ms, 3),
                       "GFLOPs": round(flops / ms / 1e6), "roundtrip_rel_l2": rt, "forward_vs_default_rel_l2": dev,
                       "kernels_last_transform": k, "passes": passes, "hbm_peak": peak}
            except Exception as e:  # noqa: BLE001
                out = {"workload": name, "config": cname, "error": repr(e)[:300]}
            print(json.dumps(out), flush=True)
            if "error" in out:
                sys.stderr.write("%-18s %-40s ERROR %s\n" % (name, cname, out["error"]))
            else:
                sys.stderr.write("%-18s %-40s %8.3f ms  rt %.1e  dev %.1e  %s\n" % (
                    name, cname, ms, rt, dev, " ".join("%s%d:%.2f" % (p["dir"][0], p["pass"], p["ms"]) for p in passes)))
            del F
            torch.cuda.empty_cache()
        del u, fu, u2, ref
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
