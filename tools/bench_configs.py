#!/usr/bin/env python
"""Kernel timings of the other BASELINE.json configs (C2, C4, C5) at full size -- not the driver's bench line, but the
numbers DESIGN.md / profiles quote for those rows. CUDA events on the launching stream, inputs resident in HBM.

    python tools/bench_configs.py [--configs C2,C4,C5] [--steps 5] [--scale 1.0]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from misha_b200 import gpu, synth  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timed(torch, fn, steps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def report(name, ms, alg_bytes, rows, extra=None):
    d = {"config": name, "ms": round(ms, 4), "rows": int(rows), "rows_per_s": rows / ms * 1e3, "alg_bytes": int(alg_bytes),
         "alg_gbs": alg_bytes / ms / 1e6, "frac_of_measured_hbm_peak": alg_bytes / ms / 1e6 / PEAK}
    if extra:
        d.update(extra)
    print(json.dumps(d), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="C2,C4,C5")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--scale", type=float, default=1.0)
    args = ap.parse_args()
    import torch
    torch.cuda.set_device(0)
    stream = torch.cuda.current_stream().cuda_stream
    chroms = synth.genome(args.scale)
    sizes = [s for _, s in chroms]
    ctx = gpu.Context(sizes)
    for opt in os.environ.get("MG_OPTIONS", "").split(","):  # experiments: MG_OPTIONS=name=value,...
        if "=" in opt:
            ctx.set_option(opt.split("=")[0], int(opt.split("=")[1]))
    want = args.configs.split(",")
    sc = np.arange(len(chroms), dtype=np.int32)
    ss, se = np.zeros(len(chroms), np.int64), np.array(sizes, np.int64)

    if "C2" in want:
        dense = gpu.DenseTrack(ctx, 20)
        nb = 0
        for i in range(len(chroms)):
            b = synth.dense_bins(i, sizes[i], 20)
            nb += len(b)
            dense.stage(i, b)
        rows = ctx.rows_fixedbin(20, sc, ss, se)
        part = gpu.summary_init(ctx)
        ms = timed(torch, lambda: dense.summary_dev(rows, "avg", part, stream=stream), args.steps)
        report("C2 gsummary (dense 20bp, whole genome, track-resolution iterator)", ms, 4 * nb, rows.n)
        counts = ctx.alloc(100, np.uint64).zero()
        br = np.linspace(0, 100, 101)
        ms = timed(torch, lambda: dense.hist_dev(rows, "avg", br, counts, stream=stream), args.steps)
        report("C2 gdist 100 bins (dense 20bp, whole genome)", ms, 4 * nb, rows.n)
        # generic (non-streaming) path: iterator 200 bp = 10 bins per row through the prefix sums
        rows2 = ctx.rows_fixedbin(200, sc, ss, se)
        ms = timed(torch, lambda: dense.summary_dev(rows2, "avg", part, stream=stream), args.steps)
        report("C2b gsummary iterator=200 (prefix path)", ms, 4 * nb, rows2.n)
        dense.free()

    if "C3x" in want:
        # the order-dependent functions (k_dense_ext, warp per row) and stddev at the C3 shape: off the headline path, timed for the record
        dense = gpu.DenseTrack(ctx, 20)
        nb = 0
        for i in range(len(chroms)):
            b = synth.dense_bins(i, sizes[i], 20)
            nb += len(b)
            dense.stage(i, b)
        counts = synth.interval_counts(chroms, int(10_000_000 * args.scale))
        cs, ss3, se3 = [], [], []
        for i in range(len(chroms)):
            s3, e3 = synth.intervals(i, sizes[i], counts[i])
            cs.append(np.full(len(s3), i, np.int32)); ss3.append(s3); se3.append(e3)
        rows3 = ctx.rows_upload(np.concatenate(cs), np.concatenate(ss3), np.concatenate(se3))
        out = ctx.alloc(rows3.n, np.float64)
        for name, f in (("first", "first"), ("last.pos", "last.pos"), ("max.pos", "max.pos"), ("lse", "lse"), ("stddev", "stddev")):
            ms = timed(torch, lambda: dense.window_dev(rows3, [f], -5000, 5000, out=[out], stream=stream), args.steps)
            report("C3x %s, +-5 kb, 10 M intervals" % name, ms, 4 * nb + rows3.n * 28, rows3.n)
        del dense

    if "C4" in want:
        sp = gpu.SparseTrack(ctx)
        iv = gpu.IntervalSet(ctx)
        cnt = synth.interval_counts(chroms, int(5_000_000 * args.scale))
        nrec = 0
        for i in range(len(chroms)):
            s, e, v = synth.sparse_records(i, sizes[i], int(cnt[i]))
            sp.stage(i, s, e, v)
            iv.stage(i, s, e)
            nrec += len(s)
        rows = ctx.rows_fixedbin(20, sc, ss, se)
        out = [ctx.alloc(rows.n, np.float64)]
        ms1 = timed(torch, lambda: sp.window_dev(rows, ["nearest"], out=out, stream=stream), args.steps)
        report("C4 sparse nearest @ 20bp iterator", ms1, 20 * nrec + 8 * rows.n, rows.n)
        out2 = ctx.alloc(rows.n, np.float64)
        ms2 = timed(torch, lambda: iv.coverage_dev(rows, out=out2, stream=stream), args.steps)
        report("C4 coverage @ 20bp iterator", ms2, 16 * nrec + 8 * rows.n, rows.n)
        report("C4 nearest + coverage (two launches)", ms1 + ms2, 20 * nrec + 16 * rows.n, rows.n)
        # the whole gscreen: both columns, the comparison predicate on the device, the run merge (synchronises: returns the count)
        flag = ctx.alloc(rows.n, np.int8)
        scratch = (ctx.alloc(rows.n, np.int32), ctx.alloc(rows.n, np.int64), ctx.alloc(rows.n, np.int64))

        def screen():
            sp.window_dev(rows, ["nearest"], out=out, stream=stream)
            iv.coverage_dev(rows, out=out2, stream=stream)
            gpu.screen_compare(ctx, [out[0], out2], [">", ">"], [0.5, 0.25], False, rows.n, dflag=flag, stream=stream)
            return gpu.screen_merge(ctx, rows, flag, stream=stream, scratch=scratch)
        t0 = time.perf_counter()
        res = screen()
        for _ in range(args.steps - 1):
            screen()
        wall = (time.perf_counter() - t0) / args.steps * 1e3
        report("C4 gscreen(near > 0.5 & cov > 0.25) end to end: columns + device predicate + merge + intervals to host (wall clock)", wall,
               20 * nrec + 16 * rows.n, rows.n, {"intervals_out": int(len(res[0]))})
        sp.free(); iv.free()

    if "C5" in want:
        seq = gpu.Sequence(ctx)
        t0 = time.perf_counter()
        for i in range(len(chroms)):
            seq.stage(i, synth.sequence(i, sizes[i]))
        stage_s = time.perf_counter() - t0
        n = int(1_000_000 * args.scale)
        cnt = synth.interval_counts(chroms, n)
        ch, st, en = [], [], []
        for i in range(len(chroms)):
            s, e = synth.intervals(i, sizes[i], int(cnt[i]), 500, 500)
            ch.append(np.full(len(s), i, np.int32)); st.append(s); en.append(e)
        rows = ctx.rows_upload(np.concatenate(ch), np.concatenate(st), np.concatenate(en))
        logp = gpu.pssm_logp(synth.pssm(20), 0.01)
        out = ctx.alloc(rows.n, np.float64)
        for mode in ("pwm", "pwm.max"):
            ms = timed(torch, lambda: seq.pwm_dev(rows, logp, bidirect=True, extend=True, mode=mode, out=out, stream=stream), args.steps)
            anchors = rows.n * 500
            report("C5 %s 20-mer bidirect, 1M x 500bp" % mode, ms, rows.n * (130 + 8), rows.n,
                   {"anchors_per_s": anchors / ms * 1e3, "strand_scores_per_s": 2 * anchors / ms * 1e3, "seq_staging_s": round(stage_s, 2)})
        for kmer, strand in (("TTAGGG", 0), ("GGCCAGGCTGGTCTCGAACTCC", 1)):
            ms = timed(torch, lambda: seq.kmer_dev(rows, kmer, strand=strand, out=out, stream=stream), args.steps)
            report("C5-shaped kmer.count %d-mer strand %d, 1M x 500bp" % (len(kmer), strand), ms, rows.n * (130 + 8), rows.n,
                   {"starts_per_s": rows.n * 500 / ms * 1e3})
        seq.free()
    if "Q" in want:
        # gintervals.quantiles shape: a 20 M row value column in HBM cut into segments (one per scope interval), 5 quantiles each.
        # The call is host-synchronous (offsets up, quantile columns down), so it is timed on the host clock.
        rng = np.random.default_rng(5)
        n = int(20_000_000 * args.scale)
        col = np.where(rng.random(n) < 0.05, np.nan, rng.lognormal(size=n))
        d = ctx.alloc(n, np.float64).from_host(col)
        p = np.array([0.1, 0.25, 0.5, 0.75, 0.9])
        for name, lo, hi in (("4-36 rows", 4, 37), ("100-400 rows", 100, 401), ("2000-12000 rows", 2000, 12001)):
            sz = rng.integers(lo, hi, int(n / ((lo + hi) / 2)) + 1)
            sf = np.concatenate([[0], np.cumsum(sz)])
            sf = sf[sf <= n].astype(np.int64)
            for _ in range(2):
                gpu.quantiles_segmented(ctx, d, sf, p)
            t0 = time.perf_counter()
            for _ in range(args.steps):
                gpu.quantiles_segmented(ctx, d, sf, p)
            ms = (time.perf_counter() - t0) * 1e3 / args.steps
            t0 = time.perf_counter()
            for s in range(200):  # the loop of whole-device mg_quantiles calls this replaced, on the first 200 segments
                gpu.quantiles(ctx, d, int(sf[s + 1] - sf[s]), p, first=int(sf[s]))
            loop_ms = (time.perf_counter() - t0) * 1e3 / 200 * (len(sf) - 1)
            report("Q gintervals.quantiles, segments of %s, 5 quantiles (host clock, copies included)" % name, ms, 8 * int(sf[-1]), int(sf[-1]),
                   {"segments": len(sf) - 1, "per_segment_loop_ms_extrapolated": round(loop_ms, 1)})
    if "B" in want:
        # gbins.summary / gbins.quantiles / gintervals.summary shape: 50 M rows, value and bin columns in HBM, 100 bins (host clock:
        # the calls return host arrays)
        rng = np.random.default_rng(6)
        n = int(50_000_000 * args.scale)
        dv = ctx.alloc(n, np.float64).from_host(np.where(rng.random(n) < 0.05, np.nan, rng.lognormal(size=n)).astype(np.float32))
        db_ = ctx.alloc(n, np.float64).from_host(np.clip(rng.lognormal(size=n) * 10, 0, 100).astype(np.float32))
        br = np.linspace(0, 100, 101)
        p = np.array([0.1, 0.5, 0.9])
        sz = rng.integers(100, 401, n // 250 + 1)
        sf = np.concatenate([[0], np.cumsum(sz)])
        sf = sf[sf <= n].astype(np.int64)
        for name, fn, byts in (("gbins.summary, 100 bins", lambda: gpu.bins_summary(ctx, dv, [db_], n, [br]), 16 * n),
                               ("gbins.quantiles, 100 bins, 3 quantiles", lambda: gpu.bins_quantiles(ctx, dv, [db_], n, [br], p), 16 * n),
                               ("gintervals.summary, %d segments of 100-400 rows" % (len(sf) - 1), lambda: gpu.summary_segmented(ctx, dv, sf), 8 * n)):
            for _ in range(2):
                fn()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                fn()
            report("B " + name + " (host clock)", (time.perf_counter() - t0) * 1e3 / args.steps, byts, n)
    ctx.close()


if __name__ == "__main__":
    main()
