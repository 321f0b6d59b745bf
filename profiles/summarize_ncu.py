#!/usr/bin/env python
"""Summarise an .ncu-rep (full set) into a small CSV of the metrics DESIGN.md / bench.py cite.
usage: python profiles/summarize_ncu.py gpurun_out/X.ncu-rep profiles/X_summary.csv"""
import csv
import subprocess
import sys

KEEP = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__maximum_warps_per_active_cycle_pct",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg.per_second",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(k) for k in KEEP if k in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[i] for i in idx])


if __name__ == "__main__":
    main()
