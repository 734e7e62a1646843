#!/usr/bin/env python
"""C3 through ONE process driving several GPUs (mg_group_h_dense_window): end-to-end ms per step for groups of 1 .. all devices,
host rows in pinned memory, host columns out. python tools/bench_group.py [--steps 5] [--intervals 10000000]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from misha_b200 import gpu, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--intervals", type=int, default=10_000_000)
    args = ap.parse_args()
    import torch
    ndev = torch.cuda.device_count()
    chroms = synth.genome()
    sizes = [s for _, s in chroms]
    counts = synth.interval_counts(chroms, args.intervals)
    bins = [synth.dense_bins(i, sizes[i], 20) for i in range(len(sizes))]
    ch, st, en = [], [], []
    for i in range(len(sizes)):
        s, e = synth.intervals(i, sizes[i], int(counts[i]))
        ch.append(np.full(len(s), i, np.int32)); st.append(s); en.append(e)
    ch, st, en = np.concatenate(ch), np.concatenate(st), np.concatenate(en)
    funcs = ["avg", "max", ("quantile", 0.9)]
    ref = None
    for n in sorted({1, 2, 4, 8} & set(range(1, ndev + 1))):
        grp = gpu.Group(sizes, list(range(n)))
        t = gpu.GroupDenseTrack(grp, 20)
        for i, b in enumerate(bins):
            t.stage(i, b)
        ctx0 = gpu.Context.__new__(gpu.Context)  # pinned buffers through member 0's context
        ctx0.lib, ctx0.h = grp.lib, grp.lib.mg_group_ctx(grp.h, 0)
        hc, hs, he = ctx0.pinned(len(ch), np.int32), ctx0.pinned(len(ch), np.int64), ctx0.pinned(len(ch), np.int64)
        hc[:], hs[:], he[:] = ch, st, en
        out = [ctx0.pinned(len(ch), np.float64) for _ in funcs]
        for _ in range(2):
            t.window_host(hc, hs, he, funcs, -5000, 5000, out=out)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            t.window_host(hc, hs, he, funcs, -5000, 5000, out=out)
        ms = (time.perf_counter() - t0) / args.steps * 1e3
        cols = [o.copy() for o in out]
        if ref is None:
            ref = cols
        same = all(np.array_equal(a, b, equal_nan=True) for a, b in zip(cols, ref))
        load = [sum(sizes[i] for i in range(len(sizes)) if grp.owner(i) == k) for k in range(n)]
        print(json.dumps({"devices": n, "ms_per_step_e2e": round(ms, 3), "intervals_per_s": args.intervals / ms * 1e3, "equal_to_one_device": bool(same),
                          "bins_owned_max_over_mean": round(max(load) / (sum(load) / n), 3)}), flush=True)
        t.free(); grp.close()


if __name__ == "__main__":
    main()
