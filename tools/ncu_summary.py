#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page) into the handful of metrics DESIGN.md / profiles/ quote.  usage: ncu_summary.py rep [out.md]"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instruction (warp execution efficiency)"),
    ("smsp__issue_active.avg.pct", "issue-slot utilisation %"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "IPC (per SM, elapsed)"),
    ("l1tex__t_sector_hit_rate.pct", "L1 sector hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 sector hit %"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard (warps/issue)"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch_resolving"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
    ("smsp__inst_executed.sum", "warp instructions"),
]


def main():
    rep = sys.argv[1]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        out.append(f"### {d.get('Kernel Name', '?')}  (id {d.get('ID', '?')})")
        for key, label in WANT:
            if key in d and d[key] != "":
                out.append(f"- {label}: {d[key]} {units[hdr.index(key)]}  (`{key}`)")
        out.append("")
    text = "\n".join(out)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
