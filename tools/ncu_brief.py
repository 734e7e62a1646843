#!/usr/bin/env python
"""One-screen summary (time, DRAM bytes, issue rate, occupancy, stall reasons) of every kernel in an .ncu-rep: python tools/ncu_brief.py rep"""
import csv, io, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("##", d["Kernel Name"])
    for k in KEYS:
        if k in d:
            print("   %-70s %s %s" % (k, d[k], units[hdr.index(k)]))
    st = [(k, float(d[k])) for k in hdr if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and d[k] not in ("", "n/a")]
    st.sort(key=lambda x: -x[1])
    print("   stalls/issue: " + ", ".join("%s %.2f" % (k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v) for k, v in st[:7]))
