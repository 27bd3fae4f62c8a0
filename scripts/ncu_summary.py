#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full capture) into a small CSV for profiles/.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_xxx.csv
    python scripts/ncu_summary.py gpurun_out/ncu_raw_xxx.csv profiles/r01_xxx.csv   (raw page exported on the box)
"""
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "lts__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    if rep.endswith(".csv"):  # already exported on the GPU box with `ncu -i x.ncu-rep --page raw --csv`
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + ["%s#%s" % (r[idx["Kernel Name"]].replace("void b200fft::", "")[:70], r[idx["ID"]]) for r in data])
        for m in METRICS:
            if m in idx:
                w.writerow([m, units[idx[m]]] + [r[idx[m]] for r in data])
    print("wrote", out)


if __name__ == "__main__":
    main()
