#!/usr/bin/env python
"""bench.py -- iterator intervals/sec of gextract over dense-track window vtracks at hg38 scale (BASELINE.json metric).

Workload (configs[2] of BASELINE.json, the one the metric is quoted on; SURVEY.md 8(d) "C3"): synthetic 3.1 Gbp genome
(24 hg38-length chromosomes), 20 bp dense track (155 M float32, log-normal, 5 % NaN runs, 0.01 % inf), 10 M sorted
non-overlapping intervals (100-300 bp), three vtracks avg / max / quantile(0.9) with sshift = -5000, eshift = +5000.
One "step" = one gextract pass producing the three value columns for every interval.

  python bench.py [--gpus N] [--steps K] [--warmup W]            this repo's CUDA path (libmishagpu.so)
  python bench.py --impl reference ...                           the reference's own CPU arithmetic (oracle/_ref),
                                                                 all host threads it would use, bounded sample per step

N > 1 is launched by torch.distributed.run, one rank per GPU; every rank takes one contiguous 1/N slice of the
chromosome-major rows and stages the chromosomes its slice touches, there is no data-path collective (each rank owns a
disjoint row range) and the total job is fixed whatever N -> reported as "strong".
`value`: the K timed steps are captured once in a CUDA graph and replayed between two CUDA events on the launching stream
(barrier + synchronize on both sides, max over ranks). `e2e`: the same step through mg_h_dense_window with pinned host arrays.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from misha_b200 import synth  # noqa: E402

BIN = 20
SSHIFT, ESHIFT = -5000, 5000
FUNCS = ["avg", "max", ("quantile", 0.9)]
N_INTERVALS = 10_000_000
METRIC = "iterator intervals/sec (gextract vtrack avg, hg38-scale)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3", choices=["C2", "C3", "C4", "C5"],
                    help="BASELINE.json config to run; C3 (default) is the one the metric is quoted on, the others print the same kind of line "
                         "for their own workload (tools/bench_lines.py)")
    ap.add_argument("--intervals", type=int, default=N_INTERVALS, help=argparse.SUPPRESS)  # smaller = NOT the named config
    ap.add_argument("--scale", type=float, default=1.0, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true", help=argparse.SUPPRESS)
    # development aid: one process on one GPU takes the rows rank R of W would take (where does a rank's time go?); NOT a bench line
    ap.add_argument("--emulate-slice", default="", help=argparse.SUPPRESS)
    return ap.parse_args()


class ClockSampler:
    """SM clock and throttle reasons sampled every ~3 ms through NVML while the timed regions run (nvidia-smi's own
    polling is too coarse for regions that last tens of milliseconds). mark(True/False) brackets the timed regions; only
    samples taken inside them count."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, gpu_index):
        self.samples = []
        self.active = False
        self.stop_flag = False
        self.err = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].isdigit() else gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception as e:  # noqa: BLE001 -- no NVML: report it, never fail the bench
            self.err = "nvml unavailable: %s" % e

    def _pump(self):
        nv = self.nv
        while not self.stop_flag:
            if self.active:
                try:
                    sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    try:
                        rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                    except Exception:  # noqa: BLE001 -- older bindings
                        rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                    self.samples.append((sm, rs))
                except Exception as e:  # noqa: BLE001
                    self.err = str(e)
                    return
            time.sleep(0.003)

    def mark(self, on):
        self.active = on

    def stop(self):
        if self.err and not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.err], "samples": 0}
        self.stop_flag = True
        self.t.join(timeout=1)
        sm = [x for x, _ in self.samples]
        reasons = sorted({name for _, rs in self.samples for name, bit in self.REASONS if rs & bit})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_sm, "reasons": reasons, "samples": len(sm),
                "how": "NVML, ~3 ms period, inside the timed regions only"}


def build_workload(args, rank, world):
    """Rows are chromosome-major and every rank takes one contiguous 1/world slice of them (the reference's own
    multitasking also cuts chromosomes by range, src/rdbinterval.cpp:335-421): equal rows per rank whatever the chromosome
    lengths. A rank stages the chromosomes its slice touches. mine = [(chromid, first interval, last interval + 1)]."""
    chroms = synth.genome(args.scale)
    counts = synth.interval_counts(chroms, args.intervals)
    total = int(sum(counts))
    lo, hi = rank * total // world, (rank + 1) * total // world
    mine, off = [], 0
    for i, c in enumerate(counts):
        a, b = max(lo, off), min(hi, off + int(c))
        if a < b:
            mine.append((i, a - off, b - off))
        off += int(c)
    return chroms, counts, mine


def bind_to_gpu_cpus(gpu_index):
    """Pins this process to the CPUs next to its GPU (NVML's ideal affinity) BEFORE any pinned host memory is allocated, so that the
    row / result buffers are first touched on the GPU's own NUMA node. Returns what was done, for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].isdigit() else gpu_index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w in range(words) for b in range(64) if (int(mask[w]) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return "no NVML affinity within the allowed CPUs"
        os.sched_setaffinity(0, cpus)
        return "bound to the GPU's %d local CPUs (NVML affinity)" % len(cpus)
    except Exception as e:  # noqa: BLE001 -- binding is an optimisation, never a failure
        return "not bound (%s)" % e


def copy_bound(torch, ctx, h_in, h_out, barrier, steps):
    """The hardware bound of the end-to-end step: the same bytes, H2D from and D2H into the same pinned buffers, as plain
    cudaMemcpyAsync on two streams (both directions at once), every rank at the same time. ms per step on this rank."""
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    d_in = [ctx.alloc(a.size, a.dtype) for a in h_in]
    d_out = [ctx.alloc(a.size, a.dtype) for a in h_out]

    def once():
        for d, a in zip(d_in, h_in):
            ctx._check(ctx.lib.mg_copy_h2d(ctx.h, ctypes.c_void_p(d.ptr), ctypes.c_void_p(a.ctypes.data), ctypes.c_int64(a.nbytes), ctypes.c_void_p(s_in.cuda_stream)))
        for d, a in zip(d_out, h_out):
            ctx._check(ctx.lib.mg_copy_d2h(ctx.h, ctypes.c_void_p(a.ctypes.data), ctypes.c_void_p(d.ptr), ctypes.c_int64(a.nbytes), ctypes.c_void_p(s_out.cuda_stream)))
    once()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        once()
    s_in.synchronize()
    s_out.synchronize()
    dt = (time.perf_counter() - t0) / steps
    barrier()
    for d in d_in + d_out:
        d.free()
    return dt * 1e3


def copy_bound_chunked(torch, ctx, h_in, h_out, barrier, steps, chunk_rows=3 << 17, slots=3):
    """The same bytes the way a pipeline has to move them: in chunks of the library's default size, a chunk's three uploads and then its three
    downloads on the chunk's slot stream (three slots in turn) -- mg_h_pipeline's copies without its kernels. ms per step on this rank."""
    streams = [torch.cuda.Stream() for _ in range(slots)]
    d_in = [ctx.alloc(a.size, a.dtype) for a in h_in]
    d_out = [ctx.alloc(a.size, a.dtype) for a in h_out]
    n = h_in[0].size
    vp, i64 = ctypes.c_void_p, ctypes.c_int64

    def once():
        k = 0
        for a0 in range(0, n, chunk_rows):
            m = min(chunk_rows, n - a0)
            st = vp(streams[k % slots].cuda_stream)
            k += 1
            for d, a in zip(d_in, h_in):
                ctx._check(ctx.lib.mg_copy_h2d(ctx.h, vp(d.ptr + a0 * a.itemsize), vp(a.ctypes.data + a0 * a.itemsize), i64(m * a.itemsize), st))
            for d, a in zip(d_out, h_out):
                ctx._check(ctx.lib.mg_copy_d2h(ctx.h, vp(a.ctypes.data + a0 * a.itemsize), vp(d.ptr + a0 * a.itemsize), i64(m * a.itemsize), st))
    once()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        once()
    for st in streams:
        st.synchronize()
    dt = (time.perf_counter() - t0) / steps
    barrier()
    for d in d_in + d_out:
        d.free()
    return dt * 1e3


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def run_ours(args):
    import torch
    import torch.distributed as dist
    from misha_b200 import gpu

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    numa = bind_to_gpu_cpus(local)
    if args.emulate_slice:
        if args.emulate_slice.startswith("c:"):  # one whole chromosome
            chroms, counts, _ = build_workload(args, 0, 1)
            ci = int(args.emulate_slice[2:])
            mine = [(ci, 0, int(counts[ci]))]
        else:
            er, ew = (int(x) for x in args.emulate_slice.split("/"))
            chroms, counts, mine = build_workload(args, er, ew)
    else:
        chroms, counts, mine = build_workload(args, rank, world)
    sizes = [s for _, s in chroms]
    ctx = gpu.Context(sizes, device=local)
    stream = torch.cuda.current_stream().cuda_stream
    for opt in os.environ.get("MG_OPTIONS", "").split(","):  # experiments: MG_OPTIONS=name=value,...
        if "=" in opt:
            ctx.set_option(opt.split("=")[0], int(opt.split("=")[1]))

    # ---- stage the track once (one-time cost, reported separately) ----
    dense = gpu.DenseTrack(ctx, BIN)
    nbins_mine = 0
    host_bins = {}
    stage_s = 0.0
    for i, _, _ in mine:
        b = synth.dense_bins(i, sizes[i], BIN)  # synthesis is not staging
        t0 = time.perf_counter()
        dense.stage(i, b)
        ctx.sync()
        stage_s += time.perf_counter() - t0
        if rank == 0 and world == 1:
            host_bins[i] = b  # the CPU baseline reads the same data

    # ---- intervals in pinned host memory ----
    n_mine = int(sum(b - a for _, a, b in mine))
    h_chrom, h_start, h_end = ctx.pinned(n_mine, np.int32), ctx.pinned(n_mine, np.int64), ctx.pinned(n_mine, np.int64)
    off = 0
    for i, a, b in mine:
        s, e = synth.intervals(i, sizes[i], int(counts[i]))
        s, e = s[a:b], e[a:b]
        h_chrom[off:off + len(s)] = i
        h_start[off:off + len(s)] = s
        h_end[off:off + len(s)] = e
        # bins this slice's windows span: the track bytes of this rank's algorithmic traffic
        nbins_mine += (min(int(e[-1]) + ESHIFT, sizes[i]) - max(int(s[0]) + SSHIFT, 0) + BIN - 1) // BIN
        off += len(s)
    h_out = [ctx.pinned(n_mine, np.float64) for _ in FUNCS]

    rows = ctx.rows_upload(h_chrom, h_start, h_end)
    d_out = [ctx.alloc(n_mine, np.float64) for _ in FUNCS]

    def step(st=None):
        dense.window_dev(rows, FUNCS, SSHIFT, ESHIFT, out=d_out, stream=stream if st is None else st)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local) if rank == 0 else None

    # ---- kernel-resident throughput: inputs already in HBM ----
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    # The K steps are captured once in a CUDA graph and replayed inside the timed region: at N = 8 a step is an 80 us kernel and
    # the Python / ctypes cost of enqueueing it would otherwise be part of what is timed. Falls back to plain launches.
    graph = None
    try:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for _ in range(args.steps):
                step(torch.cuda.current_stream().cuda_stream)
        graph.replay()  # one untimed replay
        barrier()
    except Exception as e:  # noqa: BLE001 -- any capture problem: time the plain launches
        sys.stderr.write("bench: CUDA graph capture of the step failed (%s); timing plain launches\n" % e)
        graph = None
        barrier()
    launches0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if sampler:
        sampler.mark(True)
    ev0.record()
    if graph is not None:
        graph.replay()
    else:
        for _ in range(args.steps):
            step()
    ev1.record()
    barrier()
    if sampler:
        sampler.mark(False)
    ms_mine = ev0.elapsed_time(ev1) / args.steps
    ms_step = max_over_ranks(ms_mine)
    ms_ranks = [ms_mine]
    if world > 1:  # every rank's own step time: the slowest one sets the job's
        t = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(t, torch.tensor([ms_mine], dtype=torch.float64, device="cuda"))
        ms_ranks = [float(x.item()) for x in t]
    launches = args.steps if graph is not None else ctx.launch_count() - launches0

    # ---- the host link's bound for the end-to-end step (before it: the copies overwrite the result buffers) ----
    bound_ms = max_over_ranks(copy_bound(torch, ctx, [h_chrom, h_start, h_end], h_out, barrier, args.steps))
    bound_chunked_ms = max_over_ranks(copy_bound_chunked(torch, ctx, [h_chrom, h_start, h_end], h_out, barrier, args.steps))
    # ---- end to end through the host-buffer C-ABI call: pinned host arrays in, host columns out ----
    for _ in range(2):
        dense.window_host(h_chrom, h_start, h_end, FUNCS, SSHIFT, ESHIFT, out=h_out)
    barrier()
    if sampler:
        sampler.mark(True)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dense.window_host(h_chrom, h_start, h_end, FUNCS, SSHIFT, ESHIFT, out=h_out)
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / args.steps)
    if sampler:
        sampler.mark(False)
    launches = args.steps + (ctx.launch_count() - launches0 - (0 if graph is not None else args.steps))  # timed steps + end-to-end chunks

    # ---- per-kernel timing for the roofline (each function group alone, same stream, CUDA events) ----
    kernels = []
    for name, funcs in (("k_dense_prefix[avg]", ["avg"]), ("k_dense_rowquery[max]", ["max"]), ("k_dense_thin[quantile(0.9)]", [("quantile", 0.9)]),
                        ("k_dense_thin[avg+max+quantile(0.9)]", FUNCS)):
        outs = d_out[:len(funcs)]
        for _ in range(3):
            dense.window_dev(rows, funcs, SSHIFT, ESHIFT, out=outs, stream=stream)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if sampler:
            sampler.mark(True)
        a.record()
        for _ in range(args.steps):
            dense.window_dev(rows, funcs, SSHIFT, ESHIFT, out=outs, stream=stream)
        b.record()
        torch.cuda.synchronize()
        if sampler:
            sampler.mark(False)
        ms = a.elapsed_time(b) / args.steps
        alg = 4 * nbins_mine + n_mine * (20 + 8 * len(funcs))  # unique track bytes + interval in + value columns out
        kernels.append({"kernel": name, "ms": ms, "alg_bytes": alg, "gbs": alg / ms / 1e6})
    clocks = sampler.stop() if sampler else None

    # parity spot check of what was just timed (size-independent property: host path == device path, bitwise)
    dev_cols = [d.to_host() for d in d_out]
    dense.window_dev(rows, FUNCS, SSHIFT, ESHIFT, out=d_out, stream=stream)
    torch.cuda.synchronize()
    dev_cols = [d.to_host() for d in d_out]
    for a, b in zip(dev_cols, h_out):
        if not np.array_equal(a, b, equal_nan=True):
            raise SystemExit("bench: host-buffer path and device path disagree")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    n_total = args.intervals
    # the step is one fused launch today; the dominant kernel is the step's kernel
    # ... and the timed region IS K launches of it: its average launch duration is this rank's step time (CUDA events around the graph
    # replay); the same launches enqueued one by one from Python, as the single-function kernels above are, stay in ms_plain_launches
    dom = dict(kernels[-1])
    dom["ms_plain_launches"] = dom["ms"]
    dom["ms"] = ms_mine
    dom["gbs"] = dom["alg_bytes"] / ms_mine / 1e6
    kernels[-1] = dom
    alg_total_launch = dom["alg_bytes"]
    traffic = None  # dram__bytes_read.sum + dram__bytes_write.sum of the step's kernel, from the committed ncu --set full capture of this kernel
    # (profiles/traffic.json names the capture; a kernel renamed or replaced since finds no entry and reports null instead of a stale figure)
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if world == 1 and args.intervals == N_INTERVALS and args.scale == 1.0 and os.path.exists(tp):
        traffic = json.load(open(tp)).get("k_dense_thin[avg+max+quantile(0.9)]")
    roofline = {"bound": "hbm", "achieved": dom["gbs"], "peak": peak, "unit": "GB/s", "frac": dom["gbs"] / peak, "traffic": traffic,
                "kernel": dom["kernel"], "alg_bytes_per_launch": alg_total_launch, "peak_source": peak_src,
                "per_kernel": [{k: (round(v, 4) if isinstance(v, float) else v) for k, v in kk.items()} for kk in kernels]}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(chroms, counts, host_bins, sizes)

    line = {
        "metric": METRIC, "value": n_total / (ms_step / 1e3), "unit": "intervals/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 data / f64 accumulate", "data": "synthetic",
        "config": {"workload": "C3: gextract avg/max/quantile(0.9) vtracks, sshift/eshift +-5kb, %d sorted intervals (100-300 bp), "
                               "20 bp dense track on a synthetic %.2f Gbp genome (24 hg38-length chromosomes)" % (n_total, sum(sizes) / 1e9),
                   "intervals": n_total, "bin_size": BIN, "funcs": ["avg", "max", "quantile(0.9)"], "sharding": "contiguous 1/N slices of the chromosome-major rows; a rank stages the chromosomes its slice touches",
                   "step_launch": "the K steps replayed from one CUDA graph" if graph is not None else "K plain launches",
                   "ms_per_step_by_rank": [round(x, 4) for x in ms_ranks],
                   "l2_policy": "inputs larger than L2 (track 0.62 GB + index structures 3 GB touched + intervals 0.2 GB vs 126 MB L2)",
                   "staging_s_one_time": round(stage_s, 3), "hbm_bytes_staged_rank0": dense.bytes()},
        "e2e": {"value": n_total / e2e_s, "unit": "intervals/s", "h2d_bytes_per_step": int(n_mine * 20), "d2h_bytes_per_step": int(n_mine * 8 * len(FUNCS)),
                "ms_per_step": e2e_s * 1e3, "api": "mg_h_dense_window (pinned host arrays in, host columns out)",
                "copy_bound_ms": bound_ms, "copy_bound": "the step's H2D + D2H bytes alone, plain cudaMemcpyAsync on two streams from / into the "
                "same pinned buffers, all ranks at once (max over ranks): what the host link allows",
                "copy_bound_chunked_ms": bound_chunked_ms, "copy_bound_chunked": "the same bytes as a pipeline has to move them: chunks of 3 * 2^17 rows, a chunk's "
                "three uploads then its three downloads on one of three streams in turn -- the library's copies without its kernels", "host_numa": numa},
        "gpu_launches": int(launches), "roofline": roofline, "clocks": clocks,
    }
    if cpu:
        line["cpu_baseline"] = cpu
    if args.emulate_slice:  # development aid: the rows one rank of a larger job would take -- NOT the named workload
        line["config"]["emulated_slice"] = args.emulate_slice
        line["config"]["workload"] = "NOT THE NAMED WORKLOAD (--emulate-slice %s of): " % args.emulate_slice + line["config"]["workload"]
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def make_cpu_sample(chroms, counts, host_bins, sizes, n_sample, procs, full=False):
    """full: every interval of every chromosome (the whole C3 job), each chromosome cut into pieces of about total / (4 procs)
    intervals so that the workers end together (prepare4multitasking cuts by range too). Otherwise a bounded sample: the first
    n_sample / procs intervals of `procs` chromosomes."""
    from oracle.cpu_baseline import DenseSample
    sample = DenseSample(BIN, SSHIFT, ESHIFT, ("avg", "max", "quantile"), 0.9)
    if full:
        piece = max(1, int(sum(counts)) // (4 * max(procs, 1)))
        for i in range(len(chroms)):
            bins = host_bins[i] if i in host_bins else synth.dense_bins(i, sizes[i], BIN)
            s, e = synth.intervals(i, sizes[i], int(counts[i]))
            sample.add(i, chroms[i][0], sizes[i], bins, s, e, pieces=-(-len(s) // piece))
        return sample
    # the reference shards by chromosome: take the first intervals of `procs` chromosomes
    ids = sorted(host_bins)[:max(procs, 1)] if host_bins else list(range(min(procs, len(chroms))))
    per = max(1, n_sample // len(ids))
    for i in ids:
        bins = host_bins[i] if i in host_bins else synth.dense_bins(i, sizes[i], BIN)
        s, e = synth.intervals(i, sizes[i], int(counts[i]))
        sample.add(i, chroms[i][0], sizes[i], bins, s[:per], e[:per])
    return sample


def cpu_baseline(chroms, counts, host_bins, sizes, target_cpu_s=20.0):
    from oracle.cpu_baseline import default_procs
    procs = default_procs()
    pilot = make_cpu_sample(chroms, counts, host_bins, sizes, 20000, 1)
    dt, n, kind, _ = pilot.run(1)
    pilot.close()
    per_row = dt / max(n, 1)
    n_sample = int(min(max(target_cpu_s / per_row, 100000), 4_000_000))
    sample = make_cpu_sample(chroms, counts, host_bins, sizes, n_sample, procs)
    dt, n, kind, used = sample.run(procs)
    sample.close()
    return {"value": n / dt, "unit": "intervals/s", "cores": used, "kind": kind,
            "sample": "%d of the 10 M intervals (first %d of each of %d chromosomes), same track; forked workers sharded by chromosome as "
                      "gextract_multitask does, floor(0.7 x %d cores) = %d processes; single-thread cost %.2f us/interval; R assembly omitted"
                      % (n, n // max(used, 1), used, os.cpu_count() or 1, procs, per_row * 1e6)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.cpu_baseline import default_procs
    chroms = synth.genome(args.scale)
    counts = synth.interval_counts(chroms, args.intervals)
    sizes = [s for _, s in chroms]
    procs = default_procs()
    pilot = make_cpu_sample(chroms, counts, {}, sizes, 20000, 1)
    dt, n, kind, _ = pilot.run(1)
    pilot.close()
    per_row = dt / max(n, 1)
    steps, warm = args.steps, max(args.warmup, 1)
    budget_s = 240.0  # whole run: the full job per step when K + W passes over it fit, else a bounded sample of it
    full_s = per_row * args.intervals / procs * 1.15  # what one pass over the whole job should take with `procs` workers
    full = full_s * (steps + warm) <= budget_s
    if full:
        sample = make_cpu_sample(chroms, counts, {}, sizes, 0, procs, full=True)
    else:
        n_sample = int(min(max(budget_s / (steps + warm) * procs / per_row * 0.8, 50000), 3_000_000))
        sample = make_cpu_sample(chroms, counts, {}, sizes, n_sample, procs)
    for _ in range(warm):
        sample.run(procs)
    t0 = time.perf_counter()
    for _ in range(steps):
        _, n, kind, used = sample.run(procs)
    dt = (time.perf_counter() - t0) / steps
    sample.close()
    val = n / dt
    if full:
        desc = ("all %d intervals per step (the whole job, same config as the GPU arm); forked workers over %d range shards, "
                "floor(0.7 x %d cores) = %d processes; R data-frame assembly omitted (optimistic for the CPU)" % (n, len(sample.shards), os.cpu_count() or 1, procs))
    else:
        desc = ("%d of the %d intervals per step (first %d of each of %d chromosomes: the whole job would take %.0f s a step); forked workers "
                "sharded by chromosome, floor(0.7 x %d cores) = %d processes; R data-frame assembly omitted (optimistic for the CPU)"
                % (n, args.intervals, n // used, used, full_s, os.cpu_count() or 1, procs))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "intervals/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 data / f64 accumulate",
        "data": "synthetic",
        "config": {"workload": "C3: gextract avg/max/quantile(0.9) vtracks, sshift/eshift +-5kb, sorted intervals (100-300 bp), 20 bp dense "
                               "track on a synthetic %.2f Gbp genome; %s" % (sum(sizes) / 1e9, "the whole job per step" if full else "bounded sample per step"),
                   "intervals": args.intervals, "bin_size": BIN, "funcs": ["avg", "max", "quantile(0.9)"]},
        "cpu_baseline": {"value": val, "unit": "intervals/s", "cores": used, "kind": kind, "sample": desc},
        "e2e": {"value": val, "unit": "intervals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


if __name__ == "__main__":
    a = parse_args()
    if a.config != "C3":  # the other BASELINE.json configurations: the same kind of line, tools/bench_lines.py
        from tools import bench_lines
        if a.impl == "reference":
            bench_lines.run_reference_config(a)
        else:
            bench_lines.run_config(a)
    elif a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
