#!/usr/bin/env python
"""Markdown table of a directory of bench.py JSON lines (one file per run): ms per round trip, GFLOP/s, the
slowest FFT pass and its fraction of the HBM peak, per-direction pass times, exchange GB/s.

    python scripts/summarize_benches.py gpurun_out/r02_single [more dirs...]
"""
import glob
import json
import os
import sys


def load(path):
    for ln in open(path):
        ln = ln.strip()
        if ln.startswith("{"):
            try:
                d = json.loads(ln)
            except ValueError:
                continue
            if "ms_per_step" in d:
                return d
    return None


def main():
    rows = []
    for d in sys.argv[1:] or ["gpurun_out"]:
        for f in sorted(glob.glob(os.path.join(d, "*.json"))):
            j = load(f)
            if j is None:
                continue
            r = j.get("roofline") or {}
            passes = r.get("passes", [])
            fft = " ".join("%s%d:%.2f" % (p["dir"][0], p["step"], p["ms"]) for p in passes if p["type"] != "exchange")
            ex = [p for p in passes if p["type"] == "exchange" and p["ms"] > 0]
            exs = "%.2f ms" % sum(p["ms"] for p in ex) if ex else "-"
            nv = (r.get("nvlink") or {}).get("achieved_GBps_per_direction")
            x = j["config"].get("exchange") or {}
            rows.append((os.path.basename(f)[:-5], j["config"]["name"], j["n_gpus"], "%.3f" % j["ms_per_step"], "%.0f" % j["value"],
                         "%s %.2f" % (r.get("kernel", "?"), r.get("frac", 0)), fft, exs, "%s" % (nv if nv else "-"),
                         "%s/%s" % (x.get("transport", "-"), x.get("pipelined_chunks", "-")), "%.1e" % j.get("roundtrip_rel_l2", -1)))
    hdr = ("run", "workload", "GPUs", "ms/round trip", "GFLOP/s", "slowest pass, frac of HBM peak", "FFT passes (ms)", "exchange",
           "GB/s per dir", "transport/chunks", "rt rel L2")
    print("| " + " | ".join(hdr) + " |")
    print("|" + "---|" * len(hdr))
    for r in rows:
        print("| " + " | ".join(str(c) for c in r) + " |")


if __name__ == "__main__":
    main()
