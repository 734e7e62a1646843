#!/usr/bin/env python
"""Hot SASS instructions of one kernel in an .ncu-rep (source page): python tools/ncu_hot.py rep [kernel_index] [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(out)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}
        blocks.append(cur)
    elif row and row[0] == "Address":
        cur["hdr"] = row
    elif cur is not None and row:
        cur["rows"].append(row)
b = blocks[which]
h = b["hdr"]
ia, isrc, iex, ith, ismp = h.index("Address"), h.index("Source"), h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
rows = b["rows"]
tot = sum(int(r[iex]) for r in rows)
tots = sum(int(r[ismp]) for r in rows)
print(b["name"], "warp-instr", tot, "samples", tots, "sass lines", len(rows))
idx = sorted(range(len(rows)), key=lambda i: -int(rows[i][ismp]))[:top]
for i in sorted(idx):
    r = rows[i]
    print("%5d %6.2f%% smp %6.2f%% ex  %s" % (i, 100.0 * int(r[ismp]) / max(tots, 1), 100.0 * int(r[iex]) / max(tot, 1), r[isrc].strip()))
