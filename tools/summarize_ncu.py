#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the few numbers DESIGN.md / bench.py quote: python tools/summarize_ncu.py rep [out.md]"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "launch__grid_size", "launch__block_size",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for row in rows[2:]:
        d = dict(zip(hdr, row))
        k = {"kernel": d.get("Kernel Name", "")}
        for key in KEYS:
            if key in d and d[key] != "":
                k[key] = d[key] + " " + units[hdr.index(key)]
        res.append(k)
    text = []
    for k in res:
        text.append("### " + k["kernel"])
        text.append("")
        text.append("| metric | value |")
        text.append("|---|---|")
        for key in KEYS:
            if key in k:
                text.append("| %s | %s |" % (key, k[key]))
        text.append("")
    s = "\n".join(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(s)
    else:
        print(s)
    return res


if __name__ == "__main__":
    main()
