#!/usr/bin/env python3
"""Key metrics of one `ncu --set full` capture as a small CSV:  python tools/ncu_summary.py capture.ncu-rep > profiles/x.csv"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fmaheavy_cycles_active.sum.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_active", "smsp__inst_executed.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]


def main(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    w = csv.writer(sys.stdout)
    w.writerow(["kernel", "metric", "unit", "value"])
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")]
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                w.writerow([name, k, units[i], vals[i]])


if __name__ == "__main__":
    main(sys.argv[1])
