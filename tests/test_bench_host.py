"""Host logic of bench.py (no GPU): how the N ranks of a torchrun job cut the chromosome-major rows -- contiguous, disjoint, covering,
equal to within one row (the reference's own multitasking also cuts by range, src/rdbinterval.cpp:335-421) -- and the synthetic
workload's invariants the timed step relies on (sorted, non-overlapping intervals inside their chromosome)."""
import argparse
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from misha_b200 import synth  # noqa: E402


def _args(intervals, scale=1.0):
    return argparse.Namespace(intervals=intervals, scale=scale)


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_rank_slices_partition_the_rows(world):
    args = _args(1_000_003)
    total, seen = 0, []
    per_rank = []
    for rank in range(world):
        chroms, counts, mine = bench.build_workload(args, rank, world)
        assert int(sum(counts)) == args.intervals
        n = 0
        for cid, a, b in mine:
            assert 0 <= a < b <= int(counts[cid])
            seen.append((cid, a, b))
            n += b - a
        per_rank.append(n)
        total += n
    assert total == args.intervals
    assert max(per_rank) - min(per_rank) <= 1
    # contiguous in chromosome-major order, no row twice, none missing
    seen.sort()
    off = {cid: 0 for cid, _, _ in seen}
    for cid, a, b in seen:
        assert a == off[cid]
        off[cid] = b
    chroms, counts, _ = bench.build_workload(args, 0, 1)
    for cid, c in enumerate(counts):
        assert off.get(cid, 0) == int(c)


def test_synthetic_intervals_are_sorted_disjoint_and_inside():
    chroms = synth.genome(0.01)
    counts = synth.interval_counts(chroms, 20000)
    for cid, (_, size) in enumerate(chroms):
        s, e = synth.intervals(cid, size, int(counts[cid]))
        assert len(s) == counts[cid]
        if len(s) == 0:
            continue
        assert s[0] >= 0 and e[-1] <= size
        assert np.all(e > s) and np.all(e - s >= 100) and np.all(e - s <= 300)
        assert np.all(s[1:] > e[:-1])  # at least one base between neighbours


def test_dense_bins_hold_nan_runs_and_no_inf_after_cleaning():
    v = synth.dense_bins(3, 2_000_000, 20)
    assert v.dtype == np.float32 and len(v) == 100000
    nan = np.isnan(v)
    assert 0.01 < nan.mean() < 0.15           # ~5 % in runs
    assert np.isinf(v).sum() > 0              # the +-inf the staging pass has to turn into NaN (src/GenomeTrackFixedBin.cpp:1066-1069)
    runs = np.diff(np.flatnonzero(np.diff(np.concatenate([[0], nan.astype(np.int8), [0]]))))[::2]
    assert runs.max() > 50                     # long runs exist: the windows over them are the quantile kernel's hard case
