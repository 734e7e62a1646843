import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from misha_b200 import gpu, synth
chroms = synth.genome(); sizes = [s for _, s in chroms]
ctx = gpu.Context(sizes); dense = gpu.DenseTrack(ctx, 20)
for i in range(len(chroms)): dense.stage(i, synth.dense_bins(i, sizes[i], 20))
counts = synth.interval_counts(chroms, 10_000_000)
n = int(counts.sum())
hc, hs, he = ctx.pinned(n, np.int32), ctx.pinned(n, np.int64), ctx.pinned(n, np.int64)
off = 0
for i in range(len(chroms)):
    s, e = synth.intervals(i, sizes[i], int(counts[i])); hc[off:off+len(s)] = i; hs[off:off+len(s)] = s; he[off:off+len(s)] = e; off += len(s)
out = [ctx.pinned(n, np.float64) for _ in range(3)]
F = ["avg", "max", ("quantile", 0.9)]
for chunk in (1 << 19,):
    ctx.set_option("pipe_chunk_rows", chunk)
    for _ in range(2): dense.window_host(hc, hs, he, F, -5000, 5000, out=out)
    t0 = time.perf_counter()
    for _ in range(5): dense.window_host(hc, hs, he, F, -5000, 5000, out=out)
    dt = (time.perf_counter() - t0) / 5
    print("chunk", chunk, "ms", round(dt * 1e3, 3), "G rows/s", round(n / dt / 1e9, 3), flush=True)
# raw PCIe: one big H2D + one big D2H
d = ctx.alloc(n * 3, np.float64)
import ctypes
t0 = time.perf_counter(); ctx.lib.mg_copy_d2h(ctx.h, out[0].ctypes.data_as(ctypes.c_void_p), ctypes.c_void_p(d.ptr), ctypes.c_int64(n * 8), None); ctx.sync(); print("D2H 80MB GB/s", n * 8 / (time.perf_counter() - t0) / 1e9)
t0 = time.perf_counter(); ctx.lib.mg_copy_h2d(ctx.h, ctypes.c_void_p(d.ptr), hs.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(n * 8), None); ctx.sync(); print("H2D 80MB GB/s", n * 8 / (time.perf_counter() - t0) / 1e9)
for zc in (1, 0, 1):
    ctx.set_option("zero_copy", zc); ctx.set_option("pipe_chunk_rows", 1 << 19)
    for _ in range(2): dense.window_host(hc, hs, he, F, -5000, 5000, out=out)
    t0 = time.perf_counter()
    for _ in range(5): dense.window_host(hc, hs, he, F, -5000, 5000, out=out)
    dt = (time.perf_counter() - t0) / 5
    print("zero_copy", zc, "ms", round(dt * 1e3, 3), "G rows/s", round(n / dt / 1e9, 3), flush=True)
    res = [o.copy() for o in out]
    if zc == 0: ref0 = res
print("equal:", all(np.array_equal(a, b, equal_nan=True) for a, b in zip(res, ref0)))
