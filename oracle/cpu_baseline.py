"""CPU baseline for bench.py -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Times the reference's OWN dense-track arithmetic (oracle/_ref/libmisha_ref.so = unmodified GenomeTrackFixedBin +
StreamPercentiler compiled from /root/reference/src) on a bounded sample of the bench workload, sharded by
chromosome over forked worker processes exactly as gextract_multitask shards its scope
(src/rdbinterval.cpp:446-564; default process count floor(0.7 x cores), R/zzz.R:22), with every worker writing its
rows into one shared anonymous mapping (src/GenomeTrackExtract.cpp:1461-1474). It omits R vector assembly and the
R interpreter, so it flatters the CPU. If the compiled reference is absent the plain-C port (oracle.c) is timed
instead and the result is labelled kind="port".
"""
import mmap
import os
import shutil
import tempfile
import time

import numpy as np

from . import port, ref


def default_procs():
    return max(1, int(0.7 * (os.cpu_count() or 1)))


def _window(start, end, sshift, eshift, chromsize):
    s = np.maximum(start + sshift, 0)
    e = np.minimum(end + eshift, chromsize)
    return s, e


class DenseSample:
    """A bounded sample of the C3 workload: some chromosomes' dense files (truncated to what the sample touches) on disk
    in the reference's format, plus the sampled intervals."""

    def __init__(self, bin_size, sshift, eshift, funcs, percentile):
        self.bin_size, self.sshift, self.eshift, self.funcs, self.percentile = bin_size, sshift, eshift, funcs, percentile
        self.dir = tempfile.mkdtemp(prefix="misha_cpu_baseline_")
        self.shards = []  # (chromid, file, chromsize, start, end, bins)

    def add(self, chromid, name, chromsize, bins, start, end, pieces=1):
        """One chromosome's intervals; pieces > 1 cuts them into that many shards of consecutive intervals (the reference's
        prepare4multitasking cuts big chromosomes by range so that the workers end together, src/rdbinterval.cpp:446-564); the
        shards read the same file."""
        if len(start) == 0:
            return
        last_bin = min(len(bins), int((int(end[-1]) + self.eshift) // self.bin_size + 2))
        last_bin = max(last_bin, 1)
        path = os.path.join(self.dir, name)
        with open(path, "wb") as f:
            f.write(np.uint32(self.bin_size).tobytes())
            f.write(np.ascontiguousarray(bins[:last_bin], dtype=np.float32).tobytes())
        b = np.ascontiguousarray(bins[:last_bin], np.float32)
        cuts = np.linspace(0, len(start), max(1, min(int(pieces), len(start))) + 1).astype(np.int64)
        for a, z in zip(cuts[:-1], cuts[1:]):
            self.shards.append((chromid, path, chromsize, np.ascontiguousarray(start[a:z], np.int64), np.ascontiguousarray(end[a:z], np.int64), b))

    def n(self):
        return sum(len(s[3]) for s in self.shards)

    def close(self):
        shutil.rmtree(self.dir, ignore_errors=True)

    def _run_shard(self, shard, kind):
        chromid, path, chromsize, start, end, bins = shard
        if kind == "reference":
            ws, we = _window(start, end, self.sshift, self.eshift, chromsize)
            out = ref.dense_eval(path, chromid, ws, we, self.funcs, self.percentile)
        else:
            out = port.dense_window(bins, self.bin_size, chromsize, start, end, self.sshift, self.eshift, self.percentile, self.funcs)
        return np.stack([out[f] for f in self.funcs])

    def run(self, procs, kind=None):
        """Returns (wall seconds, rows, kind, procs used). Shards are distributed greedily over `procs` forked workers."""
        kind = kind or ("reference" if ref.available() else "port")
        (ref.lib() if kind == "reference" else port.lib())  # load before forking
        n = self.n()
        nf = len(self.funcs)
        shm = mmap.mmap(-1, max(n * nf * 8, 8))  # anonymous MAP_SHARED, like the reference's result arena
        res = np.frombuffer(shm, dtype=np.float64, count=n * nf).reshape(nf, n) if n else None
        order = sorted(range(len(self.shards)), key=lambda i: -len(self.shards[i][3]))
        procs = max(1, min(procs, len(self.shards)))
        load = [0] * procs
        assign = [[] for _ in range(procs)]
        for i in order:
            k = min(range(procs), key=lambda j: load[j])
            assign[k].append(i)
            load[k] += len(self.shards[i][3])
        offs = np.concatenate([[0], np.cumsum([len(s[3]) for s in self.shards])])
        t0 = time.perf_counter()
        pids = []
        for k in range(procs):
            pid = os.fork()
            if pid == 0:
                rc = 0
                try:
                    for i in assign[k]:
                        res[:, offs[i]:offs[i + 1]] = self._run_shard(self.shards[i], kind)
                except BaseException:  # noqa: BLE001 - a worker must never return into the parent's stack
                    rc = 1
                os._exit(rc)
            pids.append(pid)
        ok = True
        for pid in pids:
            _, status = os.waitpid(pid, 0)
            ok = ok and status == 0
        dt = time.perf_counter() - t0
        if not ok:
            raise RuntimeError("a CPU baseline worker failed")
        self.last_result = np.array(res) if n else None
        return dt, n, kind, procs
