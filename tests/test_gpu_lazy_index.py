"""Index structures are built on first use and dropped least-recently-used under the context's HBM budget (mg_dense_ensure,
option hbm_budget_mb): an avg-only virtual track holds <= 8 bytes per bin, and 30 hg38-scale dense tracks -- what the reference's own
performance tests extract from (tests/testthat/test-perf-regression.R:62-73, 160-173) -- live on one GPU together."""
import numpy as np
import pytest

from conftest import assert_close, assert_same
from misha_b200 import gpu, synth
from oracle import port

pytestmark = pytest.mark.gpu
BIN = 20


def test_groups_appear_with_the_functions_that_need_them():
    sizes = [4_000_000, 1_500_000]
    ctx = gpu.Context(sizes)
    try:
        t = gpu.DenseTrack(ctx, BIN)
        bins = [synth.dense_bins(i, s, BIN) for i, s in enumerate(sizes)]
        for i, b in enumerate(bins):
            t.stage(i, b)
        nb = sum(len(b) for b in bins)
        core = t.bytes()
        assert core <= 8 * nb + (4 << 20), "the core (bins + group-of-8 prefix) is 6 bytes per bin"
        s, e = synth.intervals(0, sizes[0], 5000)
        rows = ctx.rows_upload(np.zeros(len(s), np.int32), s, e)
        exp = port.dense_window(bins[0], BIN, sizes[0], s, e, -3000, 3000, 0.9)
        grew = []
        for funcs, names in ((["avg", "sum", "size"], ("avg", "sum", "size")), (["max", "min"], ("max", "min")), (["stddev"], ("stddev",)),
                             ([("quantile", 0.9)], ("quantile",))):
            before = t.bytes()
            got = t.window(rows, funcs, -3000, 3000)
            grew.append(t.bytes() - before)
            for col, name in zip(got, names):
                if name in ("avg", "sum", "stddev"):
                    assert_close(col, exp[name], 1e-6, name)
                else:
                    assert_same(col, exp[name], name)
        assert grew[0] == 0, "avg / sum / size need nothing beyond the core"
        assert all(g > 0 for g in grew[1:]), "min / max, stddev, quantile each bring their own structures"
        # the fixed-bin iterator bounds its windows: the chromosome-wide matrix is not built for them
        t2 = gpu.DenseTrack(ctx, BIN)
        for i, b in enumerate(bins):
            t2.stage(i, b)
        frows = ctx.rows_fixedbin(200, [0], [0], [sizes[0]])
        b0 = t2.bytes()
        q = t2.window(frows, [("quantile", 0.5)], -1000, 1000)[0]
        block_only = t2.bytes() - b0
        fs = np.arange(0, sizes[0], 200, dtype=np.int64)
        assert_same(q, port.dense_window(bins[0], BIN, sizes[0], fs, np.minimum(fs + 200, sizes[0]), -1000, 1000, 0.5, ("quantile",))["quantile"], "fixed-bin quantile")
        assert 0 < block_only < grew[3], "block-local matrices only"
        # a chromosome staged later gets the groups its track already holds
    finally:
        ctx.close()


def test_budget_drops_least_recently_used_groups():
    sizes = [6_000_000]
    ctx = gpu.Context(sizes)
    try:
        bins = synth.dense_bins(0, sizes[0], BIN)
        nb = len(bins)
        tracks = []
        for _ in range(3):
            t = gpu.DenseTrack(ctx, BIN)
            t.stage(0, bins)
            tracks.append(t)
        s, e = synth.intervals(0, sizes[0], 4000)
        rows = ctx.rows_upload(np.zeros(len(s), np.int32), s, e)
        exp = port.dense_window(bins, BIN, sizes[0], s, e, -5000, 5000, 0.9)
        full = 56 * nb  # what one track's min / max + quantile structures are budgeted at
        budget_mb = (3 * 8 * nb + full + full // 2) >> 20  # the three cores + one and a half tracks' worth of structures
        ctx.set_option("hbm_budget_mb", max(budget_mb, 1))
        for rnd in range(2):
            for k, t in enumerate(tracks):
                got = t.window(rows, ["max", ("quantile", 0.9), "avg"], -5000, 5000)
                assert_same(got[0], exp["max"], "round %d track %d max" % (rnd, k))
                assert_same(got[1], exp["quantile"], "round %d track %d quantile" % (rnd, k))
                held = sum(x.bytes() for x in tracks)
                assert held <= (max(budget_mb, 1) << 20) + (8 << 20), "the budget holds"
        assert min(x.bytes() for x in tracks) <= 8 * nb + (4 << 20), "some track was stripped back to its core"
        ctx.set_option("hbm_budget_mb", 1)
        with pytest.raises(gpu.MishaGpuError):
            tracks[0].window(rows, [("quantile", 0.9)], -5000, 5000)  # not even one track's structures fit: an error, not a wrong answer
        ctx.set_option("hbm_budget_mb", 0)
    finally:
        ctx.close()


def test_thirty_hg38_scale_tracks_on_one_gpu():
    chroms = synth.genome()
    sizes = [s for _, s in chroms]
    ctx = gpu.Context(sizes)
    try:
        bins = [synth.dense_bins(i, sizes[i], BIN) for i in range(len(sizes))]
        nb = sum(len(b) for b in bins)
        tracks = []
        for _ in range(30):
            t = gpu.DenseTrack(ctx, BIN)
            for i, b in enumerate(bins):
                t.stage(i, b)
            tracks.append(t)
        held = sum(t.bytes() for t in tracks)
        assert held <= 30 * 8 * nb, "30 tracks x 155 M bins at <= 8 bytes per bin (%.1f GB held)" % (held / 1e9)
        counts = synth.interval_counts(chroms, 2_000_000)
        ch, st, en = [], [], []
        for i in range(len(sizes)):
            s, e = synth.intervals(i, sizes[i], int(counts[i]))
            ch.append(np.full(len(s), i, np.int32)); st.append(s); en.append(e)
        ch, st, en = np.concatenate(ch), np.concatenate(st), np.concatenate(en)
        rows = ctx.rows_upload(ch, st, en)
        first = tracks[0].window(rows, ["avg"], -5000, 5000)[0]
        m = ch == 13
        assert_close(first[m], port.dense_window(bins[13], BIN, sizes[13], st[m], en[m], -5000, 5000, 0.5, ("avg",))["avg"], 1e-6, "avg vs oracle")
        out = ctx.alloc(rows.n, np.float64)
        for k in range(1, 30):
            tracks[k].window_dev(rows, ["avg"], -5000, 5000, out=[out])
            if k % 7 == 0:
                assert_same(out.to_host(), first, "track %d" % k)
        assert sum(t.bytes() for t in tracks) == held, "avg alone built nothing"
        # quantiles on three of them under a budget that holds one track's structures at a time
        ctx.set_option("hbm_budget_mb", (held + 60 * nb) >> 20)  # one track's min / max + quantile structures (~50 B/bin) fit, two do not
        q0 = tracks[3].window(rows, [("quantile", 0.9), "max"], -5000, 5000)
        for k in (11, 29):
            q = tracks[k].window(rows, [("quantile", 0.9), "max"], -5000, 5000)
            assert_same(q[0], q0[0], "quantile, track %d" % k)
            assert_same(q[1], q0[1], "max, track %d" % k)
        assert_same(q0[0][m], port.dense_window(bins[13], BIN, sizes[13], st[m], en[m], -5000, 5000, 0.9, ("quantile",))["quantile"], "quantile vs oracle")
        assert sum(t.bytes() for t in tracks) <= ((held + 60 * nb) >> 20 << 20)
    finally:
        ctx.close()
