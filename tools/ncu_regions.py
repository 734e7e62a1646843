#!/usr/bin/env python
"""Executed-instruction profile of one kernel by SASS region, from `ncu --page source --csv` output: runs of consecutive
instructions with (nearly) the same execution count are one region. python tools/ncu_regions.py src.csv [min_instr_per_warp]"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    minshow = float(sys.argv[2]) if len(sys.argv) > 2 else 4
    hdr, data = rows[1], rows[2:]
    ia, isrc, isamp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    base = int(data[0][0], 16)
    nw = float(data[0][ia])  # the first instruction runs once per warp
    runs, tot = [], 0
    for r in data:
        off, c, s = int(r[0], 16) - base, int(r[ia]), int(r[isamp])
        tot += c
        tok = r[isrc].split()
        op = tok[1] if tok[0].startswith("@") else tok[0]
        if runs and abs(runs[-1]["c"] - c) <= 0.02 * max(c, 1):
            ru = runs[-1]
            ru["n"] += 1; ru["end"] = off; ru["samp"] += s; ru["ops"].append(op); ru["tot"] += c
        else:
            runs.append(dict(start=off, end=off, c=c, n=1, samp=s, ops=[op], tot=c))
    print("warps %d, instructions per warp %.1f" % (nw, tot / nw))
    for ru in runs:
        if ru["tot"] / nw >= minshow:
            oc = " ".join("%s:%d" % kv for kv in collections.Counter(ru["ops"]).most_common(7))
            print("%05x-%05x n=%4d x%5.2f = %7.1f /warp  samples %6d  %s" % (ru["start"], ru["end"], ru["n"], ru["c"] / nw, ru["tot"] / nw, ru["samp"], oc))


if __name__ == "__main__":
    main()
