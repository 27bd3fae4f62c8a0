#!/usr/bin/env python
"""Print the per-pass table of a bench.py JSON line."""
import json
import sys
for ln in open(sys.argv[1]):
    ln = ln.strip()
    if not ln.startswith("{"):
        continue
    d = json.loads(ln)
    r = d.get("roofline", {})
    print("%s: %.2f ms/step, %.0f %s, rt err %.1e, peak %.0f" % (d["config"]["name"], d["ms_per_step"], d["value"], d["unit"],
                                                        d.get("roundtrip_rel_l2", -1), r.get("peak", 0)))
    for p in r.get("passes", []):
        print("   %s %d %-8s n=%-5d %8.3f ms  %7.1f GB/s  %.3f" % (p["dir"], p["step"], p["type"], p["len"], p["ms"], p["GBps"] or 0,
                                                              (p["GBps"] or 0) / r["peak"]))
