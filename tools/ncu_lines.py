#!/usr/bin/env python
"""Warp instructions executed and stall samples per CUDA source line of one kernel in an .ncu-rep (needs --import-source on):
python tools/ncu_lines.py rep [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source=cuda,sass"], capture_output=True, text=True).stdout
cur_file, hdr, lines = None, None, []
for row in csv.reader(io.StringIO(out)):
    if not row:
        continue
    if row[0] == "File Path":
        cur_file = row[1].split("/")[-1]
    elif row[0] == "Line No":
        hdr = row
    elif hdr and row[0] not in ("", "Function Name", "Kernel Name") and row[0].isdigit():
        iex, ismp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        lines.append((cur_file, int(row[0]), row[1].strip(), int(row[iex]) if row[iex].isdigit() else 0, int(row[ismp]) if row[ismp].isdigit() else 0))
tot = sum(l[3] for l in lines)
tots = sum(l[4] for l in lines)
print("warp-instr", tot, "samples", tots)
for f, n, src, ex, smp in sorted(lines, key=lambda l: -l[3])[:top]:
    print("%6.2f%% ex %6.2f%% smp  %s:%d  %s" % (100.0 * ex / max(tot, 1), 100.0 * smp / max(tots, 1), f, n, src[:110]))
