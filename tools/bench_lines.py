"""bench.py --config C2 | C4 | C5: the JSON line of bench.py's contract for the other BASELINE.json configurations.

  C2  gsummary + gdist (100 breaks) of the 20 bp dense track over the whole 3.09 Gbp genome, track-resolution iterator
  C4  5 M-record sparse track / interval set: nearest + coverage vtracks and gscreen(near > 0.5 & cov > 0.25) over the 20 bp iterator
  C5  pwm vtrack (20-mer PSSM, both strands, extend) over 1 M x 500 bp intervals on the packed 3.09 Gbp sequence

`value` = rows of the iterator per second with everything resident in HBM (CUDA events on the launching stream, the K steps back to
back); `e2e` = the same step through the host-facing calls (host rows in / host results out where the step has any); `roofline` =
SURVEY.md 8(d)'s algorithmic bytes of the dominant launch over its own event-timed duration; `cpu_baseline` = the reference's compiled
objects (oracle/_ref) or the C restatement on a bounded sample, forked over floor(0.7 x cores) workers.
N > 1 (torchrun): the rows are cut into contiguous 1/N slices (scope intervals for the fixed-bin iterator), a rank stages what its
slice touches; C2's partials are combined with one NCCL all-reduce per result inside the timed step.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from misha_b200 import synth  # noqa: E402

METRICS = {"C2": "iterator intervals/sec (gsummary + gdist 100 breaks, 20 bp dense track, whole genome)",
           "C4": "iterator intervals/sec (sparse nearest + coverage vtracks + gscreen, 20 bp iterator)",
           "C5": "iterator intervals/sec (pwm vtrack, 20-mer PSSM bidirect, 500 bp intervals)"}


def slice_scope(sizes, rank, world, binsize=20):
    """Contiguous 1/world slice of the genome's fixed-bin rows as scope intervals (cuts fall on bin boundaries)."""
    rows = [-(-s // binsize) for s in sizes]
    total = sum(rows)
    lo, hi = rank * total // world, (rank + 1) * total // world
    c, s, e, off = [], [], [], 0
    for i, r in enumerate(rows):
        a, b = max(lo, off), min(hi, off + r)
        if a < b:
            c.append(i); s.append((a - off) * binsize); e.append(min((b - off) * binsize, sizes[i]))
        off += r
    return np.array(c, np.int32), np.array(s, np.int64), np.array(e, np.int64)


def fork_run(tasks, procs):
    """Runs the callables over `procs` forked workers (greedy by the given weights); returns wall seconds."""
    procs = max(1, min(procs, len(tasks)))
    order = sorted(range(len(tasks)), key=lambda i: -tasks[i][0])
    load, assign = [0] * procs, [[] for _ in range(procs)]
    for i in order:
        k = min(range(procs), key=lambda j: load[j])
        assign[k].append(i); load[k] += tasks[i][0]
    t0 = time.perf_counter()
    pids = []
    for k in range(procs):
        pid = os.fork()
        if pid == 0:
            rc = 0
            try:
                for i in assign[k]:
                    tasks[i][1]()
            except BaseException:  # noqa: BLE001
                rc = 1
            os._exit(rc)
        pids.append(pid)
    ok = True
    for pid in pids:
        _, st = os.waitpid(pid, 0)
        ok = ok and st == 0
    if not ok:
        raise RuntimeError("a CPU baseline worker failed")
    return time.perf_counter() - t0, procs


class Env:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from bench import ClockSampler, bind_to_gpu_cpus, peaks
        from misha_b200 import gpu
        self.torch, self.dist, self.gpu = torch, dist, gpu
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.numa = bind_to_gpu_cpus(self.local)
        self.chroms = synth.genome(args.scale)
        self.sizes = [s for _, s in self.chroms]
        self.ctx = gpu.Context(self.sizes, device=self.local)
        for opt in os.environ.get("MG_OPTIONS", "").split(","):
            if "=" in opt:
                self.ctx.set_option(opt.split("=")[0], int(opt.split("=")[1]))
        self.stream = torch.cuda.current_stream().cuda_stream
        self.sampler = ClockSampler(self.local) if self.rank == 0 else None
        self.peak, self.peak_src = peaks()
        self.args = args

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def timed(self, fn, steps, warm=3, events=True):
        """ms per step, max over ranks: CUDA events on the launching stream (events) or the host clock (host-synchronous calls)."""
        torch = self.torch
        for _ in range(max(warm, 3)):
            fn()
        self.barrier()
        if self.sampler:
            self.sampler.mark(True)
        if events:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(steps):
                fn()
            b.record()
            self.barrier()
            ms = a.elapsed_time(b) / steps
        else:
            t0 = time.perf_counter()
            for _ in range(steps):
                fn()
            self.barrier()
            ms = (time.perf_counter() - t0) * 1e3 / steps
        if self.sampler:
            self.sampler.mark(False)
        return self.max_over_ranks(ms)

    def finish(self, line):
        if self.rank == 0:
            line["clocks"] = self.sampler.stop() if self.sampler else None
            print(json.dumps(line))
        if self.world > 1:
            self.dist.destroy_process_group()


def base_line(env, cfg, rows_total, ms_step, workload, extra_cfg):
    a = env.args
    return {"metric": METRICS[cfg], "value": rows_total / (ms_step / 1e3), "unit": "intervals/s", "n_gpus": env.world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "data": "synthetic",
            "config": dict({"workload": workload, "sharding": "contiguous 1/N slices of the rows; a rank stages what its slice touches",
                            "l2_policy": "inputs larger than L2 (126 MB)", "host_numa": env.numa}, **extra_cfg)}


# ---- C2 ----------------------------------------------------------------------------------------------------------------------------
def run_c2(args):
    env = Env(args)
    gpu, ctx, sizes = env.gpu, env.ctx, env.sizes
    from misha_b200 import dist as mdist
    sc, ss, se = slice_scope(sizes, env.rank, env.world)
    dense = gpu.DenseTrack(ctx, 20)
    host_bins, nb_mine = {}, 0
    for i in sorted(set(sc.tolist())):
        b = synth.dense_bins(i, sizes[i], 20)
        dense.stage(i, b)
        if env.rank == 0 and env.world == 1:
            host_bins[i] = b
    rows = ctx.rows_fixedbin(20, sc, ss, se)
    nb_mine = rows.n
    rows_total = int(env.sum_over_ranks(float(rows.n)))
    br = np.linspace(0, 100, 101)
    part = gpu.summary_init(ctx)
    counts = ctx.alloc(100, np.uint64)
    result = {}

    def step_dev():
        ctx._check(ctx.lib.mg_summary_init(ctx.h, gpu.c_vp(part.ptr), gpu.c_vp(env.stream)))
        counts.zero(env.stream)
        dense.summary_dev(rows, "avg", part, stream=env.stream)
        dense.hist_dev(rows, "avg", br, counts, include_lowest=True, stream=env.stream)

    def step_e2e():  # what gsummary() + gdist() return to the caller: the two results on the host, combined over the ranks
        step_dev()
        p, h = part.to_host(env.stream), counts.to_host(env.stream)
        if env.world > 1:
            p, h = mdist.reduce_summary(p), mdist.reduce_hist(h)
        result["summary"], result["hist"] = gpu.summary_finalize(p), h

    ms_step = env.timed(step_dev, args.steps)
    ms_e2e = env.timed(step_e2e, args.steps, events=False)
    ms_sum = env.timed(lambda: dense.summary_dev(rows, "avg", part, stream=env.stream), args.steps)
    ms_hist = env.timed(lambda: dense.hist_dev(rows, "avg", br, counts, include_lowest=True, stream=env.stream), args.steps)
    step_e2e()
    assert int(result["hist"].sum()) == int(result["summary"][0] - result["summary"][1]), "gdist total != non-NaN count of gsummary"
    alg = 4 * nb_mine
    line = base_line(env, "C2", rows_total, ms_step, "C2: gsummary + gdist (100 breaks, include.lowest) of a 20 bp dense track over the whole synthetic %.2f Gbp "
                     "genome, track-resolution iterator (%d rows)" % (sum(sizes) / 1e9, rows_total), {"rows": rows_total, "bin_size": 20})
    line["dtype"] = "f32 data / f64 accumulate, u64 counts"
    line["e2e"] = {"value": rows_total / (ms_e2e / 1e3), "unit": "intervals/s", "h2d_bytes_per_step": 808, "d2h_bytes_per_step": 848,
                   "ms_per_step": ms_e2e, "api": "mg_dense_summary + mg_dense_hist, results to the host" + (" + NCCL all-reduce of the partials" if env.world > 1 else "")}
    line["gpu_launches"] = int(ctx.launch_count())
    line["roofline"] = {"bound": "hbm", "achieved": alg / ms_hist / 1e6, "peak": env.peak, "unit": "GB/s", "frac": alg / ms_hist / 1e6 / env.peak,
                        "traffic": None, "kernel": "k_segs_hist (gdist)", "alg_bytes_per_launch": alg, "peak_source": env.peak_src,
                        "per_kernel": [{"kernel": "k_segs_summary (gsummary)", "ms": round(ms_sum, 4), "alg_bytes": alg, "gbs": round(alg / ms_sum / 1e6, 1)},
                                       {"kernel": "k_segs_hist (gdist)", "ms": round(ms_hist, 4), "alg_bytes": alg, "gbs": round(alg / ms_hist / 1e6, 1)}]}
    if env.rank == 0 and env.world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_c2(env, host_bins)
    env.finish(line)


def cpu_c2(env, host_bins, target_rows=40_000_000):
    """The reference's GenomeTrackFixedBin::read_interval per 20 bp row (oracle/_ref) + IntervalSummary + BinFinder restatements."""
    from oracle import port, ref
    from oracle.cpu_baseline import default_procs
    import tempfile
    import shutil
    procs = default_procs()
    kind = "reference" if ref.available() else "port"
    d = tempfile.mkdtemp(prefix="misha_cpu_c2_")
    br = np.linspace(0, 100, 101)
    per = target_rows // procs
    tasks, n = [], 0
    try:
        for i in sorted(host_bins)[:procs]:
            b = host_bins[i][:per]
            path = os.path.join(d, env.chroms[i][0])
            with open(path, "wb") as f:
                f.write(np.uint32(20).tobytes()); f.write(b.tobytes())
            s = np.arange(len(b), dtype=np.int64) * 20
            e = np.minimum(s + 20, env.sizes[i])

            def work(path=path, i=i, s=s, e=e, b=b):
                v = ref.dense_eval(path, i, s, e, ("avg",))["avg"] if kind == "reference" else port.dense_window(b, 20, env.sizes[i], s, e, 0, 0, 0.5, ("avg",))["avg"]
                port.summary(v); port.hist(v, [br], True)
            tasks.append((len(b), work)); n += len(b)
        (ref.lib() if kind == "reference" else port.lib())
        dt, used = fork_run(tasks, procs)
    finally:
        shutil.rmtree(d, ignore_errors=True)
    return {"value": n / dt, "unit": "intervals/s", "cores": used, "kind": kind,
            "sample": "the first %d bins of each of %d chromosomes (%d rows), read row by row through the reference's reader, then summary + histogram; "
                      "forked workers, floor(0.7 x %d cores)" % (per, used, n, os.cpu_count() or 1)}


# ---- C4 ----------------------------------------------------------------------------------------------------------------------------
def run_c4(args):
    env = Env(args)
    gpu, ctx, sizes = env.gpu, env.ctx, env.sizes
    sc, ss, se = slice_scope(sizes, env.rank, env.world)
    sp, iv = gpu.SparseTrack(ctx), gpu.IntervalSet(ctx)
    cnt = synth.interval_counts(env.chroms, int(5_000_000 * args.scale))
    recs, nrec = {}, 0
    for i in sorted(set(sc.tolist())):
        s, e, v = synth.sparse_records(i, sizes[i], int(cnt[i]))
        sp.stage(i, s, e, v); iv.stage(i, s, e)
        nrec += len(s)
        if env.rank == 0 and env.world == 1:
            recs[i] = (s, e, v)
    rows = ctx.rows_fixedbin(20, sc, ss, se)
    rows_total = int(env.sum_over_ranks(float(rows.n)))
    near, cov = ctx.alloc(rows.n, np.float64), ctx.alloc(rows.n, np.float64)
    flag = ctx.alloc(rows.n, np.int8)
    cap = 16_000_000  # intervals the pinned host arrays can take (the step produces ~ 2 M)
    host_iv = (ctx.pinned(cap, np.int32), ctx.pinned(cap, np.int64), ctx.pinned(cap, np.int64))
    out = {}

    def cols():
        sp.window_dev(rows, ["nearest"], out=[near], stream=env.stream)
        iv.coverage_dev(rows, out=cov, stream=env.stream)

    def step_dev():  # the two vtrack columns + the filter expression, everything on the device
        cols()
        gpu.screen_compare(ctx, [near, cov], [">", ">"], [0.5, 0.25], False, rows.n, dflag=flag, stream=env.stream)

    def step_e2e():  # gscreen as the caller sees it: the merged intervals on the host
        step_dev()
        out["iv"] = gpu.screen_merge_host(ctx, rows, flag, host_iv, stream=env.stream)

    ms_step = env.timed(step_dev, args.steps)
    ms_e2e = env.timed(step_e2e, args.steps, events=False)
    ms_near = env.timed(lambda: sp.window_dev(rows, ["nearest"], out=[near], stream=env.stream), args.steps)
    ms_cov = env.timed(lambda: iv.coverage_dev(rows, out=cov, stream=env.stream), args.steps)
    ms_cmp = env.timed(lambda: gpu.screen_compare(ctx, [near, cov], [">", ">"], [0.5, 0.25], False, rows.n, dflag=flag, stream=env.stream), args.steps)
    ms_merge = env.timed(lambda: gpu.screen_merge_host(ctx, rows, flag, host_iv, stream=env.stream), args.steps, events=False)
    n_out = len(out["iv"][0])
    alg_near, alg_cov = 20 * nrec + 8 * rows.n, 16 * nrec + 8 * rows.n
    line = base_line(env, "C4", rows_total, ms_step, "C4: nearest (5 M-record sparse track) + coverage (the same intervals) vtracks and the gscreen predicate "
                     "near > 0.5 & cov > 0.25 over the 20 bp iterator of the synthetic %.2f Gbp genome (%d implicit rows)" % (sum(sizes) / 1e9, rows_total),
                     {"rows": rows_total, "records": int(sum(cnt)), "iterator": 20})
    line["dtype"] = "f32 values / f64 accumulate, i64 coordinates"
    line["e2e"] = {"value": rows_total / (ms_e2e / 1e3), "unit": "intervals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(n_out * 20),
                   "ms_per_step": ms_e2e, "api": "mg_sparse_window + mg_coverage + mg_screen_compare + mg_h_screen_merge (intervals into pinned host arrays); the iterator's rows "
                   "are implicit: no per-step host input", "intervals_out_rank0": int(n_out),
                   "attribution_ms": {"nearest": round(ms_near, 4), "coverage": round(ms_cov, 4), "predicate": round(ms_cmp, 4),
                                      "run merge + compaction + D2H of the intervals (host clock, synchronises)": round(ms_merge, 4)}}
    line["gpu_launches"] = int(ctx.launch_count())
    line["roofline"] = {"bound": "hbm", "achieved": alg_near / ms_near / 1e6, "peak": env.peak, "unit": "GB/s", "frac": alg_near / ms_near / 1e6 / env.peak,
                        "traffic": None, "kernel": "k_sparse_merge<NEAREST> (nearest)", "alg_bytes_per_launch": alg_near, "peak_source": env.peak_src,
                        "per_kernel": [{"kernel": "k_sparse_merge<NEAREST>", "ms": round(ms_near, 4), "alg_bytes": alg_near, "gbs": round(alg_near / ms_near / 1e6, 1)},
                                       {"kernel": "k_coverage_merge", "ms": round(ms_cov, 4), "alg_bytes": alg_cov, "gbs": round(alg_cov / ms_cov / 1e6, 1)},
                                       {"kernel": "k_screen_compare", "ms": round(ms_cmp, 4), "alg_bytes": 17 * rows.n, "gbs": round(17 * rows.n / ms_cmp / 1e6, 1)}]}
    if env.rank == 0 and env.world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_c4(env, recs)
    env.finish(line)


def cpu_c4(env, recs, target_rows=24_000_000):
    """GenomeTrackSparse::read_interval (compiled reference) for nearest; the coverage / predicate / run-merge restatements."""
    from oracle import port, ref
    from oracle.cpu_baseline import default_procs
    import tempfile
    import shutil
    procs = default_procs()
    kind = "reference" if ref.available() else "port"
    d = tempfile.mkdtemp(prefix="misha_cpu_c4_")
    per = target_rows // procs
    tasks, n = [], 0
    try:
        for i in sorted(recs)[:procs]:
            rs, re, rv = recs[i]
            synth.write_sparse_chrom(d, "sp", env.chroms[i][0], rs, re, rv)
            path = os.path.join(synth.track_dir(d, "sp"), env.chroms[i][0])
            m = min(per, -(-env.sizes[i] // 20))
            s = np.arange(m, dtype=np.int64) * 20
            e = np.minimum(s + 20, env.sizes[i])

            def work(path=path, i=i, s=s, e=e, rs=rs, re=re, rv=rv):
                nr = ref.sparse_eval(path, i, s, e, ("nearest",))["nearest"] if kind == "reference" else port.sparse_window(rs, re, rv, env.sizes[i], s, e, 0, 0, 0.5, ("nearest",))["nearest"]
                cv = port.coverage(rs, re, env.sizes[i], s, e)
                with np.errstate(invalid="ignore"):
                    f = ((nr > 0.5) & (cv > 0.25)).astype(np.int8)
                port.screen_merge(np.full(len(s), i, np.int32), s, e, f)
            tasks.append((m, work)); n += m
        (ref.lib() if kind == "reference" else port.lib())
        dt, used = fork_run(tasks, procs)
    finally:
        shutil.rmtree(d, ignore_errors=True)
    return {"value": n / dt, "unit": "intervals/s", "cores": used, "kind": kind,
            "sample": "the first %d rows of each of %d chromosomes (%d rows): nearest through the reference's sparse reader, coverage / predicate / run merge "
                      "through the C restatement; forked workers, floor(0.7 x %d cores)" % (per, used, n, os.cpu_count() or 1)}


# ---- C5 ----------------------------------------------------------------------------------------------------------------------------
def run_c5(args):
    env = Env(args)
    gpu, ctx, sizes = env.gpu, env.ctx, env.sizes
    n_total = int(1_000_000 * args.scale)
    cnt = synth.interval_counts(env.chroms, n_total)
    lo, hi = env.rank * n_total // env.world, (env.rank + 1) * n_total // env.world
    seq = gpu.Sequence(ctx)
    ch, st, en, off, kept = [], [], [], 0, {}
    stage_s = 0.0
    for i in range(len(sizes)):
        a, b = max(lo, off), min(hi, off + int(cnt[i]))
        if a < b:
            bases = synth.sequence(i, sizes[i])  # synthesis is not staging
            t0 = time.perf_counter()
            seq.stage(i, bases)
            ctx.sync()
            stage_s += time.perf_counter() - t0
            if env.rank == 0 and env.world == 1 and len(kept) < 4:
                kept[i] = bases
            s, e = synth.intervals(i, sizes[i], int(cnt[i]), 500, 500)
            ch.append(np.full(b - a, i, np.int32)); st.append(s[a - off:b - off]); en.append(e[a - off:b - off])
        off += int(cnt[i])
    ch, st, en = np.concatenate(ch), np.concatenate(st), np.concatenate(en)
    n_mine = len(ch)
    h_chrom, h_start, h_end, h_out = ctx.pinned(n_mine, np.int32), ctx.pinned(n_mine, np.int64), ctx.pinned(n_mine, np.int64), ctx.pinned(n_mine, np.float64)
    h_chrom[:], h_start[:], h_end[:] = ch, st, en
    rows = ctx.rows_upload(h_chrom, h_start, h_end)
    pssm = synth.pssm(20)
    logp = gpu.pssm_logp(pssm, 0.01)
    out = ctx.alloc(n_mine, np.float64)
    ms_step = env.timed(lambda: seq.pwm_dev(rows, logp, bidirect=True, extend=True, mode="pwm", out=out, stream=env.stream), args.steps)
    ms_max = env.timed(lambda: seq.pwm_dev(rows, logp, bidirect=True, extend=True, mode="pwm.max", out=out, stream=env.stream), args.steps)
    ms_e2e = env.timed(lambda: seq.pwm_host(h_chrom, h_start, h_end, logp, bidirect=True, extend=True, mode="pwm", out=h_out), args.steps, events=False)
    dev = seq.pwm(rows, logp, bidirect=True, extend=True, mode="pwm")
    if not np.array_equal(dev, h_out, equal_nan=True):
        raise SystemExit("bench C5: host-buffer path and device path disagree")
    alg = n_mine * (130 + 8)
    anchors = n_mine * 500
    line = base_line(env, "C5", n_total, ms_step, "C5: pwm vtrack (20-mer PSSM from Dirichlet(0.5) rows, bidirect, prior 0.01, extend) over %d x 500 bp intervals on the "
                     "2-bit packed synthetic %.2f Gbp sequence" % (n_total, sum(sizes) / 1e9), {"intervals": n_total, "motif_length": 20, "seq_staging_s_one_time": round(stage_s, 2)})
    line["dtype"] = "f32 log-probabilities, f32 log-sum-exp (pwm_strict: f64 exp / log)"
    line["e2e"] = {"value": n_total / (ms_e2e / 1e3), "unit": "intervals/s", "h2d_bytes_per_step": int(n_mine * 20), "d2h_bytes_per_step": int(n_mine * 8),
                   "ms_per_step": ms_e2e, "api": "mg_h_pwm (pinned host arrays in, host column out)"}
    line["gpu_launches"] = int(ctx.launch_count())
    line["roofline"] = {"bound": "hbm", "achieved": alg / ms_step / 1e6, "peak": env.peak, "unit": "GB/s", "frac": alg / ms_step / 1e6 / env.peak, "traffic": None,
                        "kernel": "k_pwm<BIDIRECT, TOTAL>", "alg_bytes_per_launch": alg, "peak_source": env.peak_src,
                        "note": "not HBM-bound (SURVEY.md 8(d)): 2 x 20 shared-memory table adds per anchor and one expf / logf pair per strand score; "
                                "the figure that describes it is strand scores per second",
                        "strand_scores_per_s": 2 * anchors / (ms_step / 1e3) * env.world if env.world == 1 else None,
                        "per_kernel": [{"kernel": "k_pwm<BIDIRECT, TOTAL> (pwm)", "ms": round(ms_step, 4), "alg_bytes": alg, "anchors_per_s": anchors / ms_step * 1e3},
                                       {"kernel": "k_pwm<BIDIRECT, MAX> (pwm.max)", "ms": round(ms_max, 4), "alg_bytes": alg, "anchors_per_s": anchors / ms_max * 1e3}]}
    if env.rank == 0 and env.world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_c5(env, kept, pssm, ch, st, en)
    env.finish(line)


def cpu_c5(env, kept, pssm, ch, st, en, target_rows=40_000):
    """PWMScorer::score_interval of the compiled reference over a sample of the intervals (sequence files in the reference's format)."""
    from oracle import port, ref
    from oracle.cpu_baseline import default_procs
    import tempfile
    import shutil
    procs = default_procs()
    kind = "reference" if ref.available() else "port"
    d = tempfile.mkdtemp(prefix="misha_cpu_c5_")
    names = [n for n, _ in env.chroms]
    logp = port.pssm_logp(pssm, 0.01)
    tasks, n = [], 0
    try:
        ids = sorted(kept)
        per = max(1, target_rows // max(len(ids), 1))
        for i in ids:
            synth.write_seq_chrom(d, names[i], kept[i])
        for i in ids:
            m = np.flatnonzero(ch == i)[:per]
            pieces = np.array_split(m, max(1, procs // len(ids)))
            for pc in pieces:
                if not len(pc):
                    continue

                def work(i=i, pc=pc):
                    if kind == "reference":
                        ref.pwm_eval(d, names, env.sizes, pssm, np.full(len(pc), i, np.int32), st[pc], en[pc], bidirect=True, prior=0.01, extend=True, mode="pwm")
                    else:
                        port.pwm(kept[i], logp, st[pc], en[pc], bidirect=True, extend=True, mode="pwm")
                tasks.append((len(pc), work)); n += len(pc)
        (ref.lib() if kind == "reference" else port.lib())
        dt, used = fork_run(tasks, procs)
    finally:
        shutil.rmtree(d, ignore_errors=True)
    return {"value": n / dt, "unit": "intervals/s", "cores": used, "kind": kind,
            "sample": "%d of the intervals (the first of %d chromosomes), PWMScorer::score_interval per interval; forked workers, floor(0.7 x %d cores)"
                      % (n, len(kept), os.cpu_count() or 1)}


def run_config(args):
    {"C2": run_c2, "C4": run_c4, "C5": run_c5}[args.config](args)


class CpuEnv:
    """What the cpu_* legs need of Env, without a GPU (bench.py --impl reference --config ...)."""

    def __init__(self, args):
        self.chroms = synth.genome(args.scale)
        self.sizes = [s for _, s in self.chroms]
        self.args = args


def run_reference_config(args):
    """The CPU arm of a config on its own: the same bounded sample the cpu_baseline leg times, printed as the reference arm's line."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle.cpu_baseline import default_procs
    env = CpuEnv(args)
    procs = default_procs()
    t0 = time.perf_counter()
    if args.config == "C2":
        bins = {i: synth.dense_bins(i, env.sizes[i], 20) for i in range(min(procs, len(env.sizes)))}
        cpu = cpu_c2(env, bins)
    elif args.config == "C4":
        cnt = synth.interval_counts(env.chroms, int(5_000_000 * args.scale))
        recs = {i: synth.sparse_records(i, env.sizes[i], int(cnt[i])) for i in range(min(procs, len(env.sizes)))}
        cpu = cpu_c4(env, recs)
    else:
        cnt = synth.interval_counts(env.chroms, int(1_000_000 * args.scale))
        ids = sorted(range(len(env.sizes)), key=lambda i: env.sizes[i])[:4]  # the four shortest chromosomes: synthesis is not what is timed
        kept = {i: synth.sequence(i, env.sizes[i]) for i in ids}
        ch, st, en = [], [], []
        for i in ids:
            s, e = synth.intervals(i, env.sizes[i], int(cnt[i]), 500, 500)
            ch.append(np.full(len(s), i, np.int32)); st.append(s); en.append(e)
        cpu = cpu_c5(env, kept, synth.pssm(20), np.concatenate(ch), np.concatenate(st), np.concatenate(en))
    line = {"impl": "reference", "metric": METRICS[args.config], "value": cpu["value"], "unit": "intervals/s", "n_gpus": args.gpus, "steps": 1, "warmup": 0,
            "ms_per_step": None, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "data": "synthetic",
            "config": {"workload": args.config + " (bounded sample, see cpu_baseline.sample)", "setup_s": round(time.perf_counter() - t0, 1)},
            "cpu_baseline": cpu, "e2e": {"value": cpu["value"], "unit": "intervals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))
