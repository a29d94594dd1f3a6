#!/usr/bin/env python
"""Text summary of `ncu --set full` reports (gpurun_out/prof_<kernel>.ncu-rep) for profiles/: the metrics the
roofline discussion in DESIGN.md uses, one block per kernel.  Usage: summarize_ncu_full.py out.txt rep [rep ...]"""
import csv
import io
import os
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size", "launch__cluster_size",
    "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "sm__cycles_elapsed.max",
]


def main():
    out, reps = sys.argv[1], sys.argv[2:]
    lines = []
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            lines.append(f"{rep}: unreadable")
            continue
        head, units, vals = rows[0], rows[1], rows[2]
        name = vals[head.index("Kernel Name")] if "Kernel Name" in head else os.path.basename(rep)
        lines.append(f"kernel: {name}   [{os.path.basename(rep)}; ncu --set full --clock-control none --import-source on]")
        for w in WANT:
            if w in head:
                i = head.index(w)
                lines.append(f"  {w:84s} {units[i]:16s} {vals[i]}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
