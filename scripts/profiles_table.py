#!/usr/bin/env python
"""Markdown table of a profiles/*_full.csv summary (metrics x kernels, written by scripts/ncu_summary.py).

    python scripts/profiles_table.py profiles/r01c_slab1024_f64_full.csv
"""
import csv
import sys

ROWS = [("gpu__time_duration.sum", "time (ms, under ncu: cold, serialised)", 1.0, "%.2f"),
        ("dram__bytes_read.sum", "DRAM read (GB)", 1.0, "%.2f"),
        ("dram__bytes_write.sum", "DRAM written (GB)", 1.0, "%.2f"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput (% of ncu peak)", 1.0, "%.0f"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "LSU wavefronts (% of peak)", 1.0, "%.0f"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts (% of peak)", 1.0, "%.0f"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active (%)", 1.0, "%.0f"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active (% of peak)", 1.0, "%.0f"),
        ("launch__registers_per_thread", "registers / thread", 1.0, "%.0f"),
        ("launch__block_size", "threads / CTA", 1.0, "%.0f"),
        ("launch__shared_mem_per_block_dynamic", "dynamic shared memory / CTA (KB)", 1.0, "%.0f"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard (warps / issue)", 1.0, "%.2f"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier", 1.0, "%.2f"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall: MIO throttle", 1.0, "%.2f"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard", 1.0, "%.2f")]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr = rows[0]
    data = {r[0]: r for r in rows[1:]}
    names = [h.replace("void fft_kernel<", "").replace(">(Params)", "").replace("double, ", "f64 ").replace("float, ", "f32 ") for h in hdr[2:]]
    print("| metric | " + " | ".join("`%s`" % n for n in names) + " |")
    print("|---|" + "---|" * len(names))
    for key, label, scale, fmt in ROWS:
        if key not in data:
            continue
        vals = []
        for v in data[key][2:]:
            try:
                vals.append(fmt % (float(v.replace(",", "")) * scale))
            except ValueError:
                vals.append(v)
        print("| %s | %s |" % (label, " | ".join(vals)))


if __name__ == "__main__":
    main()
