"""Parity of the CUDA dense-track window functions (mg_dense_window / mg_h_dense_window) with the reference.

Bar: bit-exact for min / max / quantile / size, <= 1e-6 relative for avg / sum / stddev (north_star); the golden
vectors come from the reference's own compiled GenomeTrackFixedBin (tests/golden/make_fixtures.py).
"""
import numpy as np
import pytest

from conftest import CHROMS, assert_close, assert_same, read_dense
from misha_b200 import gpu
from oracle import port

pytestmark = pytest.mark.gpu

RTOL = 1e-6


def test_golden_all_functions(gpu_ctx, dense_testdb, golden):
    for cid, (c, size) in enumerate(CHROMS):
        s, e = golden[f"dense_{c}_start"], golden[f"dense_{c}_end"]
        rows = gpu_ctx.rows_upload(np.full(len(s), cid), s, e)
        funcs = ["avg", "sum", "min", "max", "stddev", ("quantile", 0.9), ("quantile", 0.5), ("quantile", 0.0), ("quantile", 1.0)]
        out = dense_testdb.window(rows, funcs)
        assert_close(out[0], golden[f"dense_{c}_avg"], RTOL, f"{c} avg")
        assert_close(out[1], golden[f"dense_{c}_sum"], RTOL, f"{c} sum")
        assert_same(out[2], golden[f"dense_{c}_min"], f"{c} min")
        assert_same(out[3], golden[f"dense_{c}_max"], f"{c} max")
        assert_close(out[4], golden[f"dense_{c}_stddev"], RTOL, f"{c} stddev")
        assert_same(out[5], golden[f"dense_{c}_quantile"], f"{c} q0.9")
        assert_same(out[6], golden[f"dense_{c}_quantile_0.5"], f"{c} q0.5")
        assert_same(out[7], golden[f"dense_{c}_quantile_0.0"], f"{c} q0")
        assert_same(out[8], golden[f"dense_{c}_quantile_1.0"], f"{c} q1")


def test_prefix_path_avg_sum(gpu_ctx, dense_testdb, golden):
    """avg/sum alone go through the O(1) prefix kernel; also the reference's sliding (Kahan) scan."""
    for cid, (c, size) in enumerate(CHROMS):
        for key in ("", "slide_"):
            s, e = golden[f"dense_{c}_{key}start"], golden[f"dense_{c}_{key}end"]
            rows = gpu_ctx.rows_upload(np.full(len(s), cid), s, e)
            avg, sm = dense_testdb.window(rows, ["avg", "sum"])
            assert_close(avg, golden[f"dense_{c}_{key}avg"], RTOL, f"{c} {key}avg")
            assert_close(sm, golden[f"dense_{c}_{key}sum"], RTOL, f"{c} {key}sum")


def test_shifts_and_clamp_vs_oracle(gpu_ctx, dense_testdb):
    rng = np.random.default_rng(7)
    for cid, (c, size) in enumerate(CHROMS):
        bs, bins = read_dense("dense_track", c)
        s = np.sort(rng.integers(0, size - 1, 3000))
        e = np.minimum(s + rng.integers(1, 400, 3000), size)
        rows = gpu_ctx.rows_upload(np.full(len(s), cid), s, e)
        for sshift, eshift in ((-5000, 5000), (-100, -50), (300, 100), (0, 0), (-10**7, 10**7), (2000, -2000)):
            funcs = ["avg", "sum", "min", "max", "stddev", ("quantile", 0.9), "size"]
            got = dense_testdb.window(rows, funcs, sshift, eshift)
            exp = port.dense_window(bins, bs, size, s, e, sshift, eshift, 0.9)
            for k, name in enumerate(("avg", "sum", "min", "max", "stddev", "quantile", "size")):
                if name in ("avg", "sum", "stddev"):
                    assert_close(got[k], exp[name], RTOL, f"{c} {name} shift {sshift},{eshift}")
                else:
                    assert_same(got[k], exp[name], f"{c} {name} shift {sshift},{eshift}")


def test_host_entry_point_matches_device_path(gpu_ctx, dense_testdb):
    rng = np.random.default_rng(11)
    chrom = np.sort(rng.integers(0, 3, 5000)).astype(np.int32)
    sizes = np.array([z for _, z in CHROMS])[chrom]
    start = (rng.random(5000) * (sizes - 2)).astype(np.int64)
    end = np.minimum(start + rng.integers(1, 3000, 5000), sizes)
    funcs = ["avg", "max", ("quantile", 0.9)]
    host = dense_testdb.window_host(chrom, start, end, funcs, -500, 500)
    rows = gpu_ctx.rows_upload(chrom, start, end)
    dev = dense_testdb.window(rows, funcs, -500, 500)
    for a, b in zip(host, dev):
        assert_same(a, b, "host vs device path")


def test_empty_and_ragged(gpu_ctx, dense_testdb):
    rows = gpu_ctx.rows_upload(np.zeros(0, np.int32), np.zeros(0, np.int64), np.zeros(0, np.int64))
    out = dense_testdb.window(rows, ["avg", "max"])
    assert len(out[0]) == 0 and len(out[1]) == 0
    # one row; windows wider than the warp buffer (quantile radix-select path); whole chromosome
    bs, bins = read_dense("dense_track", "chr1")
    s = np.array([0, 100, 250000, 0]); e = np.array([500000, 400100, 250001, 1])
    rows = gpu_ctx.rows_upload(np.zeros(4, np.int32), s, e)
    got = dense_testdb.window(rows, ["avg", "min", "max", ("quantile", 0.9), ("quantile", 0.123), "stddev"])
    exp = port.dense_window(bins, bs, 500000, s, e, 0, 0, 0.9)
    exp2 = port.dense_window(bins, bs, 500000, s, e, 0, 0, 0.123, ("quantile",))
    assert_close(got[0], exp["avg"], RTOL, "avg")
    assert_same(got[1], exp["min"], "min")
    assert_same(got[2], exp["max"], "max")
    assert_same(got[3], exp["quantile"], "q.9 wide")
    assert_same(got[4], exp2["quantile"], "q.123 wide")
    assert_close(got[5], exp["stddev"], RTOL, "stddev")


def test_cancellation_guard(gpu_ctx):
    """A huge value early in the chromosome must not poison windows after it (prefix cancellation guard)."""
    from misha_b200 import gpu
    size = 200000
    ctx2 = gpu.Context([size])
    try:
        rng = np.random.default_rng(3)
        bins = rng.random(size // 20).astype(np.float32) * 1e-3
        bins[10] = 1e30
        bins[500:600] = 0
        t = gpu.DenseTrack(ctx2, 20)
        t.stage(0, bins)
        s = np.arange(0, size - 4000, 777, dtype=np.int64)
        e = s + 4000
        rows = ctx2.rows_upload(np.zeros(len(s), np.int32), s, e)
        avg, sm = t.window(rows, ["avg", "sum"])
        exp = port.dense_window(bins, 20, size, s, e, 0, 0, 0.5, ("avg", "sum"))
        assert_close(avg, exp["avg"], RTOL, "avg")
        assert_close(sm, exp["sum"], RTOL, "sum")
    finally:
        ctx2.close()


def test_sweep_and_index_paths_agree(gpu_ctx, dense_testdb):
    """min / max / quantile through the index structures (pyramids, wavelet matrix) == the warp-per-row sweep, bitwise."""
    rng = np.random.default_rng(21)
    for cid, (c, size) in enumerate(CHROMS):
        s = rng.integers(0, size - 1, 4000)  # unsorted on purpose
        e = np.minimum(s + rng.choice([1, 19, 20, 21, 160, 161, 1000, 9999, 123456], 4000), size)
        rows = gpu_ctx.rows_upload(np.full(len(s), cid), s, e)
        funcs = ["avg", "min", "max", ("quantile", 0.9), ("quantile", 0.5), ("quantile", 0.0), ("quantile", 1.0), ("quantile", 0.3333)]
        idx = dense_testdb.window(rows, funcs, -300, 700)
        gpu_ctx.set_option("force_sweep", 1)
        try:
            swp = dense_testdb.window(rows, funcs, -300, 700)
        finally:
            gpu_ctx.set_option("force_sweep", 0)
        assert_close(idx[0], swp[0], RTOL, f"{c} avg")
        for k in range(1, len(funcs)):
            assert_same(idx[k], swp[k], f"{c} {funcs[k]}")


def test_stddev_prefix_squares_guard():
    """stddev comes from prefix sums of v and v^2; low-variance / constant / tiny windows must fall back to the exact
    in-order Welford update instead of returning a cancelled difference."""
    from misha_b200 import gpu
    size = 20 * 60000
    ctx2 = gpu.Context([size])
    try:
        rng = np.random.default_rng(12)
        bins = (1000 + rng.normal(0, 1e-3, size // 20)).astype(np.float32)  # mean / sd = 1e6
        bins[20000:30000] = 7.25                                              # constant stretch: sd == 0 exactly
        bins[30000:40000] = rng.lognormal(size=10000).astype(np.float32)      # ordinary data
        bins[35000:35100] = np.nan
        bins[50000] = 3e18                                                    # one huge value poisons the running sum of squares
        t = gpu.DenseTrack(ctx2, 20)
        t.stage(0, bins)
        s = np.sort(rng.integers(0, size - 20, 6000))
        e = np.minimum(s + rng.choice([20, 40, 200, 2000, 20000, 400000], 6000), size)
        rows = ctx2.rows_upload(np.zeros(len(s), np.int32), s, e)
        got = t.window(rows, ["stddev", "avg"])
        exp = port.dense_window(bins, 20, size, s, e, 0, 0, 0.5, ("stddev", "avg"))
        assert_close(got[0], exp["stddev"], RTOL, "stddev")
        assert_close(got[1], exp["avg"], RTOL, "avg")
        const = (s >= 20000 * 20) & (e <= 30000 * 20) & (e - s > 20)
        assert const.any() and np.all(got[0][const] == 0.0)
    finally:
        ctx2.close()


def test_block_wavelet_matches_global_and_oracle():
    """Quantiles of windows <= 1024 bins come from the block-local wavelet matrices, larger ones from the chromosome-wide
    matrix: both must equal the oracle bit for bit -- NaN runs, ties, negative values, block-boundary windows."""
    from misha_b200 import gpu
    nb = 2048 * 5 + 777
    size = nb * 20 - 7
    ctx2 = gpu.Context([size])
    try:
        rng = np.random.default_rng(99)
        bins = rng.normal(0, 3, nb).astype(np.float32)
        bins[rng.integers(0, nb, 600)] = np.float32(1.5)           # ties
        bins[1000:1100] = np.nan
        bins[2040:2060] = np.nan                                  # a NaN run across a block boundary
        bins[4096:5200] = np.nan                                  # windows with no values at all
        bins[rng.integers(0, nb, 300)] = np.nan
        bins[7000:7300] = np.round(bins[7000:7300])                # many duplicates
        t = gpu.DenseTrack(ctx2, 20)
        t.stage(0, bins)
        n = 20000
        s = rng.integers(0, size - 1, n)
        wl = rng.choice([1, 20, 399, 5000, 10240, 20479, 20480, 20481, 20500, 40000], n)
        # windows that start / end exactly on block (1024-bin) boundaries
        s[:2000] = (rng.integers(0, nb // 1024, 2000) * 1024 + rng.integers(-1, 2, 2000)).clip(0) * 20
        e = np.minimum(s + wl, size)
        rows = ctx2.rows_upload(np.zeros(n, np.int32), s, e)
        funcs = [("quantile", 0.9), ("quantile", 0.5), ("quantile", 0.0), ("quantile", 1.0), ("quantile", 0.777), "max", "avg"]
        blk = t.window(rows, funcs)
        ctx2.set_option("block_wm", 0)
        try:
            glb = t.window(rows, funcs)
        finally:
            ctx2.set_option("block_wm", 1)
        for k in range(len(funcs)):
            if funcs[k] == "avg":
                assert_close(blk[k], glb[k], RTOL, "avg")
            else:
                assert_same(blk[k], glb[k], f"block vs global {funcs[k]}")
        for k, p in enumerate((0.9, 0.5, 0.0, 1.0, 0.777)):
            exp = port.dense_window(bins, 20, size, s, e, 0, 0, p, ("quantile",))
            assert_same(blk[k], exp["quantile"], f"block wavelet q{p} vs oracle")
    finally:
        ctx2.close()


def test_staged_block_records_any_row_order():
    """The row-query kernel stages the block records of its 256 rows in shared memory (bulk copy + mbarrier) when they are
    neighbours; rows of another chromosome, rows far apart and unsorted rows take the L1 path inside the same thread block.
    Whatever the mix: identical to stage_slots = 0 and to the oracle, bit for bit."""
    from misha_b200 import gpu
    nbs = [2048 * 6 + 300, 2048 * 3 + 11]
    sizes = [nb * 20 - 3 for nb in nbs]
    ctx2 = gpu.Context(sizes)
    try:
        rng = np.random.default_rng(4242)
        tracks = []
        t = gpu.DenseTrack(ctx2, 20)
        for c, nb in enumerate(nbs):
            bins = rng.normal(0, 2, nb).astype(np.float32)
            bins[rng.integers(0, nb, nb // 20)] = np.nan
            bins[rng.integers(0, nb, nb // 30)] = np.float32(0.25)
            tracks.append(bins)
            t.stage(c, bins)

        def check(chrom, s, e, sshift, eshift, label):
            rows = ctx2.rows_upload(chrom.astype(np.int32), s, e)
            funcs = [("quantile", 0.9), ("quantile", 0.31), "avg", "max"]
            outs = {}
            for slots in (5, 0, 1, 12):
                ctx2.set_option("stage_slots", slots)
                outs[slots] = t.window(rows, funcs, sshift=sshift, eshift=eshift)
                outs[(slots, "q")] = t.window(rows, [("quantile", 0.9)], sshift=sshift, eshift=eshift)  # count from the staged NaN level
            ctx2.set_option("stage_slots", 5)
            for slots in (0, 1, 12):
                for k in range(len(funcs)):
                    assert_same(outs[5][k], outs[slots][k], f"{label}: stage_slots 5 vs {slots}, {funcs[k]}")
                assert_same(outs[(5, "q")][0], outs[(slots, "q")][0], f"{label}: quantile-only, stage_slots 5 vs {slots}")
            assert_same(outs[(5, "q")][0], outs[5][0], f"{label}: quantile-only vs fused")
            for c in range(len(nbs)):
                m = chrom == c
                if m.any():
                    for k, p in ((0, 0.9), (1, 0.31)):
                        exp = port.dense_window(tracks[c], 20, sizes[c], s[m], e[m], sshift, eshift, p, ("quantile",))
                        assert_same(outs[5][k][m], exp["quantile"], f"{label}: q{p} vs oracle, chrom {c}")

        # dense sorted rows over both chromosomes (thread blocks that straddle the chromosome change)
        per = [3000, 1500]
        chrom = np.repeat(np.arange(2), per)
        s = np.concatenate([np.sort(rng.integers(0, sizes[c] - 400, per[c])) for c in range(2)])
        e = s + rng.integers(100, 300, len(s))
        check(chrom, s, e, -5000, 5000, "sorted")
        # the same rows shuffled
        perm = rng.permutation(len(s))
        check(chrom[perm], s[perm], e[perm], -5000, 5000, "shuffled")
        # few rows, far apart (more blocks than slots), windows of every size up to the block limit and beyond
        n = 700
        chrom = np.sort(rng.integers(0, 2, n))
        s = np.array([rng.integers(0, sizes[c] - 30000) for c in chrom])
        order = np.lexsort((s, chrom))
        chrom, s = chrom[order], s[order]
        e = s + rng.choice([1, 20, 2000, 20480, 20500, 29000], n)
        check(chrom, s, e, 0, 0, "sparse")
        # partial last thread block, every row in one block record
        s = np.arange(0, 300) * 20 + 40000
        check(np.zeros(300, np.int64), s, s + 1000, -200, 200, "one-record")
    finally:
        ctx2.close()


EXT = ("lse", "first", "last", "first.pos", "last.pos", "min.pos", "max.pos")


def test_order_dependent_functions_golden_and_oracle(gpu_ctx, dense_testdb):
    """lse / first / last / first.pos / last.pos / min.pos / max.pos (k_dense_ext) against the reference's compiled reader
    (ref_vectors_ext.npz) and the oracle: values and positions bit for bit -- the single-bin assignment quirk, ties of the
    extremum (first bin wins), NaN runs, windows clamped by the chromosome end; lse equals the reference's reducer path
    (double max-trick) to float rounding and its generic path (pairwise float) within 1e-6."""
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "ref_vectors_ext.npz"))
    for cid, (c, size) in enumerate(CHROMS):
        s, e = g[f"{c}_start"], g[f"{c}_end"]
        rows = gpu_ctx.rows_upload(np.full(len(s), cid), s, e)
        got = dense_testdb.window(rows, list(EXT))
        for k, name in enumerate(EXT):
            if name == "lse":
                np.testing.assert_allclose(got[k], g[f"dense_{c}_lse"], rtol=1e-6, atol=1e-6, equal_nan=True)
                np.testing.assert_allclose(got[k], g[f"dense_{c}_lse_reducer"], rtol=2e-7, atol=0, equal_nan=True)
                exact = (got[k] == g[f"dense_{c}_lse_reducer"]) | np.isnan(got[k])
                assert exact.mean() > 0.999, exact.mean()
            else:
                assert_same(got[k], g[f"dense_{c}_{name}"], f"{c} {name}")
        # shifted windows, relative positions, mixed with the index-structure functions in one call
        bs, bins = read_dense("dense_track", c)
        rng = np.random.default_rng(31 + cid)
        s = np.sort(rng.integers(0, size - 1, 3000))
        e = np.minimum(s + rng.choice([1, 10, 50, 120, 3000], 3000), size)
        rows = gpu_ctx.rows_upload(np.full(len(s), cid), s, e)
        for sshift, eshift in ((0, 0), (-700, 900), (40, -10), (-10**7, 10**7)):
            funcs = ["max", ("first.pos", 1), ("last.pos", 1), ("min.pos", 1), ("max.pos", 1), "first", "last", "avg"]
            got = dense_testdb.window(rows, funcs, sshift, eshift) + dense_testdb.window(rows, ["min.pos"], sshift, eshift)
            exp = port.dense_window_ext(bins, bs, size, s, e, sshift, eshift, 0)
            base = port.dense_window(bins, bs, size, s, e, sshift, eshift, 0.5, ("max", "avg"))
            ws = np.maximum(s + sshift, 0).astype(np.float64)
            assert_same(got[0], base["max"], f"{c} max next to ext {sshift}")
            assert_close(got[7], base["avg"], RTOL, f"{c} avg next to ext")
            for k, name in ((1, "first.pos"), (2, "last.pos"), (3, "min.pos"), (4, "max.pos")):
                assert_same(got[k], exp[name] - ws, f"{c} {name} relative {sshift},{eshift}")
            assert_same(got[5], exp["first"], f"{c} first {sshift}")
            assert_same(got[6], exp["last"], f"{c} last {sshift}")
            assert_same(got[8], exp["min.pos"], f"{c} min.pos abs {sshift}")


def test_order_dependent_functions_synthetic_ties_and_nans():
    from misha_b200 import gpu
    nb = 5000
    size = nb * 10 - 4
    ctx2 = gpu.Context([size])
    try:
        rng = np.random.default_rng(8)
        bins = np.round(rng.normal(0, 2, nb)).astype(np.float32)  # few distinct values: extremum ties everywhere
        bins[rng.integers(0, nb, 800)] = np.nan
        bins[1000:1400] = np.nan
        bins[3000] = np.inf
        t = gpu.DenseTrack(ctx2, 10)
        t.stage(0, bins)
        s = rng.integers(0, size - 1, 8000)
        e = np.minimum(s + rng.choice([1, 9, 10, 11, 100, 333, 2500, 6000], 8000), size)
        rows = ctx2.rows_upload(np.zeros(len(s), np.int32), s, e)
        got = t.window(rows, list(EXT), sshift=-15, eshift=25)
        exp = port.dense_window_ext(bins, 10, size, s, e, -15, 25, 0)
        for k, name in enumerate(EXT):
            if name == "lse":
                np.testing.assert_allclose(got[k], exp[name], rtol=2e-7, atol=0, equal_nan=True)
            else:
                assert_same(got[k], exp[name], name)
        # first / last alone take the thread-per-row kernel; force_sweep keeps them on the warp-per-row one
        fl = ["first", "last", "first.pos", ("last.pos", 1), "min.pos", ("max.pos", 1), "lse"]  # argext: sparse-table descent; lse: pyramid
        ws = np.maximum(s - 15, 0).astype(np.float64)
        for sweep in (0, 1):
            ctx2.set_option("force_sweep", sweep)
            try:
                got = t.window(rows, fl, sshift=-15, eshift=25)
            finally:
                ctx2.set_option("force_sweep", 0)
            assert_same(got[0], exp["first"], f"first (sweep={sweep})")
            assert_same(got[1], exp["last"], f"last (sweep={sweep})")
            assert_same(got[2], exp["first.pos"], f"first.pos (sweep={sweep})")
            assert_same(got[3], exp["last.pos"] - ws, f"last.pos relative (sweep={sweep})")
            assert_same(got[4], exp["min.pos"], f"min.pos (sweep={sweep})")
            assert_same(got[5], exp["max.pos"] - ws, f"max.pos relative (sweep={sweep})")
            np.testing.assert_allclose(got[6], exp["lse"], rtol=2e-7, atol=0, equal_nan=True)
            assert np.mean((got[6] == exp["lse"]) | np.isnan(got[6])) > 0.999
    finally:
        ctx2.close()


def test_reducer_and_generic_path_counts(gpu_ctx):
    """size / exists / lse as the reference's reducer path and as its generic path report them (MG_FUNC_GENERIC): golden vectors made
    from the compiled reference with the respective function sets registered (tests/golden/path_vectors.npz), and the restatement
    under a window shift."""
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "path_vectors.npz"))
    bins, bs, size, s, e = g["bins"], int(g["bin_size"]), int(g["chromsize"]), g["start"], g["end"]
    ctx2 = gpu.Context([size])
    try:
        t = gpu.DenseTrack(ctx2, bs)
        t.stage(0, bins)
        rows = ctx2.rows_upload(np.zeros(len(s), np.int32), s, e)
        red = t.window(rows, ["size", "exists", "lse"])
        gen = t.window(rows, ["size:generic", "exists:generic", "lse:generic", "avg", "max"])
        for k, name in enumerate(("size", "exists")):
            assert_same(red[k], g["reducer_" + name], "reducer path " + name)
            assert_same(gen[k], g["generic_" + name], "generic path " + name)
        np.testing.assert_allclose(red[2], g["reducer_lse"], rtol=1e-6, atol=1e-6, equal_nan=True)
        np.testing.assert_allclose(gen[2], g["generic_lse"], rtol=1e-6, atol=1e-6, equal_nan=True)
        assert np.sum(red[0] != gen[0]) == np.sum(g["reducer_size"] != g["generic_size"]) > 100
        # one flagged function next to an unflagged one, under shifts, through every kernel family
        for ss, es in ((-7, 3), (-300, 300)):
            exp_g = port.dense_counts(bins, bs, size, s, e, ss, es, generic=True)
            exp_r = port.dense_counts(bins, bs, size, s, e, ss, es, generic=False)
            for tk in (0, 1, 3, 4):
                ctx2.set_option("tile_kernel", tk)
                got = t.window(rows, ["size:generic", "exists", "avg", ("quantile", 0.5)], ss, es)
                assert_same(got[0], exp_g["size"], "generic size, shift %d, kernels %d" % (ss, tk))
                assert_same(got[1], exp_r["exists"], "reducer exists, shift %d, kernels %d" % (ss, tk))
            ctx2.set_option("tile_kernel", 3)
        with pytest.raises(gpu.MishaGpuError):
            t.window(rows, ["avg:generic"])
    finally:
        ctx2.close()
