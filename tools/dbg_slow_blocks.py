#!/usr/bin/env python
"""Development aid (needs a library built with -DMG_DBG_BLOCKTIME, MISHAGPU_LIB=...): which thread blocks of the C3 step live longest on one
chromosome, and what their rows look like.   python tools/dbg_slow_blocks.py <chromid>"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from misha_b200 import gpu, synth

ci = int(sys.argv[1]) if len(sys.argv) > 1 else 3
BIN, SS, ES = 20, -5000, 5000
chroms = synth.genome(1.0)
sizes = [s for _, s in chroms]
counts = synth.interval_counts(chroms, 10_000_000)
ctx = gpu.Context(sizes, device=0)
dense = gpu.DenseTrack(ctx, BIN)
bins = synth.dense_bins(ci, sizes[ci], BIN)
dense.stage(ci, bins)
s, e = synth.intervals(ci, sizes[ci], int(counts[ci]))
n = len(s)
rows = ctx.rows_upload(np.full(n, ci, np.int32), s, e)
funcs = ["avg", "max", ("quantile", 0.9)]
out = [ctx.alloc(n, np.float64) for _ in funcs]
for _ in range(3):
    dense.window_dev(rows, funcs, SS, ES, out=out)
ctx.sync()
lib = ctx.lib
nblk = (n + 255) // 256
buf = (ctypes.c_ulonglong * nblk)()
lib.mg_dbg_blocktimes(None, 0, 1)
dense.window_dev(rows, funcs, SS, ES, out=out)
ctx.sync()
lib.mg_dbg_blocktimes(buf, nblk, 0)
t = np.array(buf[:], dtype=np.float64) / 1e3
print("blocks %d  median %.1f us  p99 %.1f  max %.1f" % (nblk, np.median(t), np.percentile(t, 99), t.max()))
v = bins.copy()
v[np.isinf(v)] = np.nan
for b in np.argsort(-t)[:8]:
    r0, r1 = b * 256, min(n, b * 256 + 256)
    ws = np.maximum(s[r0:r1] + SS, 0) // BIN
    we = (np.minimum(e[r0:r1] + ES, sizes[ci]) + BIN - 1) // BIN
    nn = np.array([np.count_nonzero(~np.isnan(v[a:c])) for a, c in zip(ws, we)])
    sums = np.array([np.nansum(v[a:c].astype(np.float64)) for a, c in zip(ws, we)])
    print("block %6d  %.1f us  span %d bins  window bins %d-%d  non-NaN per window min %d max %d  min |sum| %.3g  rows with n<=5: %d" % (
        b, t[b], we.max() - ws.min(), (we - ws).min(), (we - ws).max(), nn.min(), nn.max(), np.abs(sums).min(), np.count_nonzero(nn <= 5)))
