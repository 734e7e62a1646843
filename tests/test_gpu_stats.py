"""gsummary / gdist / gscreen kernels against the oracle and the reference's known answers."""
import numpy as np
import pytest

from conftest import CHROMS, assert_close, assert_same, read_dense
from misha_b200 import gpu
from oracle import port

pytestmark = pytest.mark.gpu


def _scope_all():
    return np.arange(3, dtype=np.int32), np.zeros(3, np.int64), np.array([z for _, z in CHROMS], dtype=np.int64)


def _all_bins():
    return np.concatenate([read_dense("dense_track", c)[1] for c, _ in CHROMS])


def _clean(v):
    v = v.astype(np.float64)
    v[np.isinf(v)] = np.nan
    return v


def test_gdist_known_answers(gpu_ctx, dense_testdb):
    """tests/testthat/test-gdist.R:69-72: sum(gdist(dense_track, c(0,.2,.5,1))) == 6546 over chr1, 901 over chr1:0-50000."""
    for end, expect in ((500000, 6546), (50000, 901)):
        rows = gpu_ctx.rows_fixedbin(50, [0], [0], [end])
        counts = gpu_ctx.alloc(3, np.uint64).zero()
        dense_testdb.hist_dev(rows, "avg", [0, 0.2, 0.5, 1.0], counts)
        got = counts.to_host()
        assert int(got.sum()) == expect
        bins = _clean(read_dense("dense_track", "chr1")[1][: end // 50])
        assert np.array_equal(got, port.hist(bins, [[0, 0.2, 0.5, 1.0]]))


def test_summary_stream_and_generic(gpu_ctx, dense_testdb):
    sc, ss, se = _scope_all()
    vals = _clean(_all_bins())
    exp = port.summary(vals)
    for binsize in (50,):  # track resolution: streaming kernel
        rows = gpu_ctx.rows_fixedbin(binsize, sc, ss, se)
        part = gpu.summary_init(gpu_ctx)
        dense_testdb.summary_dev(rows, "avg", part)
        got = gpu.summary_finalize(part.to_host())
        assert_same(got[:4], exp[:4], "counts/min/max")
        assert_close(got[4:], exp[4:], 1e-9, "sum/mean/sd")
    # coarser iterator: prefix path, then sweep path (max) through a device column
    for func in ("avg", "sum", "max", ("quantile", 0.5)):
        rows = gpu_ctx.rows_fixedbin(1000, sc, ss, se)
        part = gpu.summary_init(gpu_ctx)
        dense_testdb.summary_dev(rows, func, part)
        got = gpu.summary_finalize(part.to_host())
        c, s, e, _ = rows.coords()
        name = func if isinstance(func, str) else func[0]
        cols = []
        for cid, (cn, size) in enumerate(CHROMS):
            bs, bins = read_dense("dense_track", cn)
            m = c == cid
            cols.append(port.dense_window(bins, bs, size, s[m], e[m], 0, 0, 0.5, (name,))[name])
        exp2 = port.summary(np.concatenate(cols))
        assert_same(got[:2], exp2[:2], f"{name} counts")
        assert_close(got[2:], exp2[2:], 1e-6, f"{name} moments")


def test_summary_values_and_accumulate(gpu_ctx):
    rng = np.random.default_rng(2)
    v = rng.lognormal(size=300001)
    v[rng.random(v.size) < 0.05] = np.nan
    d = gpu_ctx.alloc(v.size, np.float64).from_host(v)
    part = gpu.summary_init(gpu_ctx)
    half = 150000
    gpu.summary_add(gpu_ctx, d.ptr, half, part)
    gpu.summary_add(gpu_ctx, d.ptr + 8 * half, v.size - half, part)  # accumulation across calls (chromosome shards)
    got = gpu.summary_finalize(part.to_host())
    exp = port.summary(v)
    assert_same(got[:4], exp[:4], "counts/min/max")
    assert_close(got[4:], exp[4:], 1e-12, "moments")
    # all-NaN and empty
    d2 = gpu_ctx.alloc(10, np.float64).from_host(np.full(10, np.nan))
    part = gpu.summary_init(gpu_ctx)
    gpu.summary_add(gpu_ctx, d2.ptr, 10, part)
    got = gpu.summary_finalize(part.to_host())
    assert got[0] == 10 and got[1] == 10 and np.all(np.isnan(got[2:]))


def test_hist_bins_golden(gpu_ctx, golden):
    vals = golden["bins_vals"]
    d = gpu_ctx.alloc(vals.size, np.float64).from_host(vals)
    for name in ("uniform", "ragged", "uniform_f32"):
        br = golden[f"bins_{name}_breaks"]
        for il in (0, 1):
            ref_bins = golden[f"bins_{name}_il{il}"]
            exp = np.bincount(ref_bins[ref_bins >= 0], minlength=len(br) - 1).astype(np.uint64)
            counts = gpu_ctx.alloc(len(br) - 1, np.uint64).zero()
            gpu.hist_dev(gpu_ctx, [d], vals.size, [br], counts, include_lowest=bool(il))
            assert np.array_equal(counts.to_host(), exp), f"{name} il={il}"


def test_hist_2d_and_wide(gpu_ctx):
    rng = np.random.default_rng(4)
    n = 200000
    a, b = rng.normal(size=n), rng.random(n)
    a[::97] = np.nan
    da, db = gpu_ctx.alloc(n, np.float64).from_host(a), gpu_ctx.alloc(n, np.float64).from_host(b)
    br = [np.linspace(-3, 3, 13), np.array([0, 0.1, 0.5, 0.9, 1.0])]
    counts = gpu_ctx.alloc(12 * 4, np.uint64).zero()
    gpu.hist_dev(gpu_ctx, [da, db], n, br, counts, include_lowest=True)
    exp = port.hist(np.stack([a, b], axis=1), br, True)
    assert np.array_equal(counts.to_host().reshape(exp.shape, order="F"), exp)
    # 1-D, more bins than the private-column kernel takes
    br1 = np.linspace(-4, 4, 1001)
    counts = gpu_ctx.alloc(1000, np.uint64).zero()
    gpu.hist_dev(gpu_ctx, [da], n, [br1], counts)
    assert np.array_equal(counts.to_host(), port.hist(a, [br1]))


def test_dense_hist_paths(gpu_ctx, dense_testdb):
    sc, ss, se = _scope_all()
    br = np.linspace(0, 1, 101)
    exp = port.hist(_clean(_all_bins()), [br])
    rows = gpu_ctx.rows_fixedbin(50, sc, ss, se)
    counts = gpu_ctx.alloc(100, np.uint64).zero()
    dense_testdb.hist_dev(rows, "avg", br, counts)
    assert np.array_equal(counts.to_host(), exp)
    # windowed avg over a coarser iterator (prefix kernel + private histogram) and max (column + histogram)
    for func in ("avg", "max"):
        rows = gpu_ctx.rows_fixedbin(300, sc, ss, se)
        counts = gpu_ctx.alloc(100, np.uint64).zero()
        dense_testdb.hist_dev(rows, func, br, counts, sshift=-200, eshift=200)
        c, s, e, _ = rows.coords()
        cols = []
        for cid, (cn, size) in enumerate(CHROMS):
            bs, bins = read_dense("dense_track", cn)
            m = c == cid
            cols.append(port.dense_window(bins, bs, size, s[m], e[m], -200, 200, 0.5, (func,))[func])
        col = np.concatenate(cols)
        got, exp2 = counts.to_host(), port.hist(col, [br])
        if func == "max":
            assert np.array_equal(got, exp2)
        else:  # avg is within 1 float ulp of the reference: a value sitting on a break may move one bin
            assert got.sum() == exp2.sum() and np.abs(got.astype(np.int64) - exp2.astype(np.int64)).sum() <= 4


def test_screen_merge(gpu_ctx):
    rng = np.random.default_rng(8)
    sc, ss, se = _scope_all()
    rows = gpu_ctx.rows_fixedbin(20, sc, ss, se)
    c, s, e, _ = rows.coords()
    for p in (0.5, 0.95, 0.02, 1.0, 0.0):
        flag = (rng.random(rows.n) < p).astype(np.int8)
        flag[rng.random(rows.n) < 0.01] = -128  # NA rows end a run like FALSE
        dflag = gpu_ctx.alloc(rows.n, np.int8).from_host(flag)
        gc, gs, ge = gpu.screen_merge(gpu_ctx, rows, dflag)
        ec, es, ee = port.screen_merge(c, s, e, flag)
        assert np.array_equal(gc, ec) and np.array_equal(gs, es) and np.array_equal(ge, ee), f"p={p}"
        # the host-array entry point (mg_h_screen_merge): the same intervals; arrays one entry short get the count and nothing else
        host = (gpu_ctx.pinned(len(ec) + 3, np.int32), gpu_ctx.pinned(len(ec) + 3, np.int64), gpu_ctx.pinned(len(ec) + 3, np.int64))
        hc, hs, he = gpu.screen_merge_host(gpu_ctx, rows, dflag, host)
        assert np.array_equal(hc, ec) and np.array_equal(hs, es) and np.array_equal(he, ee), f"host arrays, p={p}"
        if len(ec):
            short = tuple(h[:len(ec) - 1] for h in host)
            assert gpu.screen_merge_host(gpu_ctx, rows, dflag, short) is None and gpu.screen_merge_host.last_count == len(ec)
    # scope intervals that touch (the run goes on across them), leave a gap, or change chromosome; sub-ranges that begin and end anywhere
    # in a 16-row chunk (the fixed-bin kernels read 16 flags per thread)
    rows = gpu_ctx.rows_fixedbin(20, np.array([0, 0, 0, 1, 1], np.int32), np.array([0, 1000, 5000, 40, 1040]), np.array([1000, 3010, 9000, 1040, 2000]))
    c, s, e, _ = rows.coords()
    for p in (1.0, 0.9, 0.5, 0.1):
        flag = (rng.random(rows.n) < p).astype(np.int8)
        for first, count in ((0, rows.n), (7, rows.n - 20), (33, 100), (48, 17), (0, 1), (rows.n - 1, 1), (49, 150)):
            sub = np.ascontiguousarray(flag[first:first + count])
            dflag = gpu_ctx.alloc(count, np.int8).from_host(sub)
            gc, gs, ge = gpu.screen_merge(gpu_ctx, rows, dflag, first, count)
            ec, es, ee = port.screen_merge(c[first:first + count], s[first:first + count], e[first:first + count], sub)
            assert np.array_equal(gc, ec) and np.array_equal(gs, es) and np.array_equal(ge, ee), f"scope seams p={p} first={first} count={count}"
    # explicit rows with gaps: touching rows merge, gapped rows do not, chromosome change splits
    ch = np.array([0, 0, 0, 0, 1, 1]); st = np.array([0, 10, 30, 40, 40, 50]); en = np.array([10, 20, 40, 50, 50, 60])
    rows = gpu_ctx.rows_upload(ch, st, en)
    dflag = gpu_ctx.alloc(6, np.int8).from_host(np.ones(6, np.int8))
    gc, gs, ge = gpu.screen_merge(gpu_ctx, rows, dflag)
    assert gc.tolist() == [0, 0, 1] and gs.tolist() == [0, 30, 40] and ge.tolist() == [20, 50, 60]


def test_stream_hist_values_on_breaks():
    """Exact float thresholds of the streaming gdist kernel: values sitting exactly on breaks, one ulp either side,
    include.lowest on and off, uniform (ceil formula) and ragged (binary search) breaks."""
    size = 20 * 40000
    ctx2 = gpu.Context([size])
    try:
        rng = np.random.default_rng(6)
        brs = [np.linspace(0, 1, 101), np.arange(11) * 0.1, np.array([0, 0.2, 0.5, 1.0]), np.linspace(-2, 7, 38)]
        pool = np.concatenate([b.astype(np.float32) for b in brs] + [np.nextafter(b.astype(np.float32), np.float32(9)) for b in brs] +
                              [np.nextafter(b.astype(np.float32), np.float32(-9)) for b in brs] + [rng.normal(0.5, 1, 5000).astype(np.float32)])
        bins = rng.choice(pool, size // 20).astype(np.float32)
        bins[::101] = np.nan
        t = gpu.DenseTrack(ctx2, 20)
        t.stage(0, bins)
        rows = ctx2.rows_fixedbin(20, [0], [0], [size])
        v = bins.astype(np.float64)
        for br in brs:
            for il in (False, True):
                counts = ctx2.alloc(len(br) - 1, np.uint64).zero()
                t.hist_dev(rows, "avg", br, counts, include_lowest=il)
                assert np.array_equal(counts.to_host(), port.hist(v, [br], il)), (len(br), il)
        part = gpu.summary_init(ctx2)
        t.summary_dev(rows, "avg", part)
        got = gpu.summary_finalize(part.to_host())
        exp = port.summary(v)
        assert_same(got[:4], exp[:4], "counts/min/max")
        assert_close(got[4:], exp[4:], 1e-12, "moments")
    finally:
        ctx2.close()


def test_screen_compare_matches_r_semantics(gpu_ctx):
    """Device-side comparison predicates: TRUE exactly where R's `&` / `|` of comparisons is TRUE (NA never is)."""
    from misha_b200 import gpu
    rng = np.random.default_rng(31)
    for n in (0, 1, 3, 4, 1001, 40000):
        a = rng.normal(size=n); b = rng.normal(size=n); c = rng.integers(0, 3, n).astype(np.float64)
        for col in (a, b, c):
            col[rng.random(n) < 0.1] = np.nan
        da, db_, dc = (gpu_ctx.alloc(n, np.float64).from_host(x) for x in (a, b, c))
        with np.errstate(invalid="ignore"):
            cases = [
                ([da], [">"], [0.3], False, a > 0.3),
                ([da, db_], [">", "<="], [0.0, 0.5], False, (a > 0.0) & (b <= 0.5)),
                ([da, db_, dc], ["<", ">=", "=="], [-0.2, 1.0, 2.0], True, (a < -0.2) | (b >= 1.0) | (c == 2.0)),
                ([dc, da], ["!=", ">"], [1.0, 10.0], True, ((c < 1.0) | (c > 1.0)) | (a > 10.0)),
            ]
        for cols, ops, thr, is_or, exp in cases:
            if n == 0:
                continue
            flag = gpu.screen_compare(gpu_ctx, cols, ops, thr, is_or, n).to_host()
            assert np.array_equal(flag.astype(bool), exp), (n, ops, is_or)


def test_quantiles_golden_and_oracle(gpu_ctx):
    """mg_quantiles (gquantiles' exact regime) against the reference's own vectors and the oracle, bit for bit."""
    import os
    from misha_b200 import gpu
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "gquantiles_vectors.npz"))
    pcts = g["pcts"]
    ok = (pcts >= 0) & (pcts <= 1)  # the entry point rejects the rest, as C_gquantiles does
    for key in g.files:
        if not key.startswith("vals_"):
            continue
        v = g[key]
        d = gpu_ctx.alloc(len(v), np.float64).from_host(v)
        got, n = gpu.quantiles(gpu_ctx, d, len(v), pcts[ok])
        assert np.array_equal(got, g["q_" + key[5:]][ok], equal_nan=True), key
        assert n == int(np.sum(~np.isnan(v.astype(np.float32))))
    with pytest.raises(gpu.MishaGpuError, match="not in"):
        gpu.quantiles(gpu_ctx, d, len(v), [1.5])
    rng = np.random.default_rng(41)
    for n in (1, 2, 1000, 300001):
        for gen in (lambda m: rng.normal(size=m), lambda m: np.round(rng.normal(size=m) * 2), lambda m: -rng.lognormal(size=m) * 1e-30,
                    lambda m: np.where(rng.random(m) < 0.5, np.nan, rng.random(m))):
            v = gen(n).astype(np.float64)
            p = np.concatenate([[0, 1, 0.5], rng.random(20)])
            d = gpu_ctx.alloc(n, np.float64).from_host(v)
            got, nk = gpu.quantiles(gpu_ctx, d, n, p)
            exp, en = port.gquantiles(v, p)
            assert nk == en and np.array_equal(got, exp, equal_nan=True), n


def test_quantiles_segmented(gpu_ctx):
    """mg_quantiles_segmented (gintervals.quantiles: one percentiler per scope interval, src/GenomeTrackQuantiles.cpp:726-741):
    every segment bit for bit against the oracle's StreamPercentiler<double> restatement and against the golden vectors --
    empty, one-row, all-NaN segments, the four tiers (warp sort, thread block sort, thread block select, multi-block select) and their borders."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "gquantiles_vectors.npz"))
    pcts = g["pcts"]
    pcts = pcts[(pcts >= 0) & (pcts <= 1)]
    keys = [k for k in g.files if k.startswith("vals_")]
    col = np.concatenate([g[k] for k in keys]).astype(np.float64)
    sf = np.concatenate([[0], np.cumsum([len(g[k]) for k in keys])])
    d = gpu_ctx.alloc(max(len(col), 1), np.float64).from_host(col)
    got, nk = gpu.quantiles_segmented(gpu_ctx, d, sf, pcts)
    for s, k in enumerate(keys):
        assert np.array_equal(got[:, s], g["q_" + k[5:]][(g["pcts"] >= 0) & (g["pcts"] <= 1)], equal_nan=True), k
        assert nk[s] == int(np.sum(~np.isnan(g[k].astype(np.float32))))

    rng = np.random.default_rng(97)
    sizes = [0, 1, 2, 31, 32, 33, 255, 256, 257, 0, 1000, 5000, 16383, 16384, 16385, 40000, 7, 0, 300001, 20000, 270000, 262144, 262145] + list(rng.integers(0, 600, 300))
    sf = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    n = int(sf[-1])
    col = np.where(rng.random(n) < 0.2, np.nan, np.round(rng.normal(size=n) * 50) / 8 + rng.normal(size=n) * (rng.random(n) < 0.5))
    col[rng.integers(0, n, 50)] = np.inf
    col[rng.integers(0, n, 50)] = -np.inf
    col[sf[10]:sf[11]] = np.nan      # a 1000-row segment without values
    col[sf[16]:sf[17]] = np.nan
    col[sf[19]:sf[20]] = np.nan      # a select-tier segment without values
    col[sf[20]:sf[21]] = np.nan      # and one of the multi-block tier
    col[sf[18]:sf[18] + 250000] = 1.5  # long runs of one value in a select-tier segment
    d = gpu_ctx.alloc(n, np.float64).from_host(col)
    p = np.concatenate([[0, 1, 0.5, 0.9], rng.random(9)])
    got, nk = gpu.quantiles_segmented(gpu_ctx, d, sf, p)
    assert got.shape == (len(p), len(sizes))
    p64 = np.concatenate([[1, 0], rng.random(62)])  # the most quantiles a call takes: 128 target ranks in the select tier
    got64, _ = gpu.quantiles_segmented(gpu_ctx, d, sf[15:24], p64)
    for s in range(15, 23):
        assert np.array_equal(got64[:, s - 15], port.gquantiles(col[sf[s]:sf[s + 1]], p64)[0], equal_nan=True), s
    for s in range(len(sizes)):
        exp, en = port.gquantiles(col[sf[s]:sf[s + 1]], p)
        assert nk[s] == en, s
        assert np.array_equal(got[:, s], exp, equal_nan=True), (s, sizes[s])
    got0, nk0 = gpu.quantiles_segmented(gpu_ctx, d, [0], p)  # no segment at all
    assert got0.shape == (len(p), 0) and len(nk0) == 0
    with pytest.raises(gpu.MishaGpuError, match="not in"):
        gpu.quantiles_segmented(gpu_ctx, d, sf, [1.5])
    with pytest.raises(gpu.MishaGpuError, match="ascend"):
        gpu.quantiles_segmented(gpu_ctx, d, [0, 10, 5], [0.5])


def _exp_summary_rows(groups):
    return np.array([port.summary(np.asarray(g, dtype=np.float32).astype(np.float64)) for g in groups]).reshape(len(groups), 7)


def _check_summary_rows(got, exp, what):
    assert_same(got[:, :4], exp[:, :4], what + " counts/min/max")  # exact
    assert_close(got[:, 4:6], exp[:, 4:6], 1e-11, what + " sum/mean")
    assert_close(got[:, 6], exp[:, 6], 1e-7, what + " sd")       # the reference's one-pass formula cancels: sd carries sqrt of the sums' last bits


def test_summary_segmented(gpu_ctx):
    """mg_summary_segmented (gintervals.summary: one IntervalSummary per scope interval, src/GenomeTrackSummary.cpp:248-431) against
    the oracle's IntervalSummary restatement per segment: empty / one-row / all-NaN segments and one beyond the warp tier."""
    rng = np.random.default_rng(12)
    sizes = [0, 1, 2, 33, 0, 1000, 7, 300000, 5] + list(rng.integers(0, 200, 500))
    sf = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    n = int(sf[-1])
    # doubles that are not floats (a host expression like track * 0.1): every segment, the 300000-row one that goes through the whole-device
    # kernels included, reads its values through float as the reference does (src/GenomeTrackSummary.cpp:314)
    col = np.where(rng.random(n) < 0.1, np.nan, rng.lognormal(size=n) * 0.1)
    col[sf[6]:sf[7]] = np.nan
    d = gpu_ctx.alloc(n, np.float64).from_host(col)
    got = gpu.summary_segmented(gpu_ctx, d, sf)
    exp = _exp_summary_rows([col[sf[s]:sf[s + 1]] for s in range(len(sizes))])
    _check_summary_rows(got, exp, "segments")
    assert gpu.summary_segmented(gpu_ctx, d, [0]).shape == (0, 7)
    with pytest.raises(gpu.MishaGpuError, match="ascend"):
        gpu.summary_segmented(gpu_ctx, d, [0, 10, 5])


def _bin_index(cols, breaks, include_lowest):
    idx = np.zeros(len(cols[0]), dtype=np.int64)
    ok = np.ones(len(cols[0]), dtype=bool)
    mult = 1
    for c, b in zip(cols, breaks):
        k = port.val2bin(b, c, include_lowest).astype(np.int64)
        ok &= k >= 0
        idx += np.where(k >= 0, k, 0) * mult
        mult *= len(b) - 1
    return np.where(ok, idx, -1), mult


@pytest.mark.parametrize("case", ["1d_few", "2d", "1d_many", "include_lowest"])
def test_bins_summary_and_quantiles(gpu_ctx, case):
    """mg_bins_summary / mg_bins_quantiles (gbins.summary, src/GenomeTrackSummary.cpp:433-506; gbins.quantiles,
    src/GenomeTrackQuantiles.cpp:1199-1330): rows grouped by BinsManager::vals2idx of their bin columns -- expected from the oracle's
    val2bin, IntervalSummary and StreamPercentiler<double> restatements. Counts / min / max / quantiles exact, sums to the last bits."""
    rng = np.random.default_rng({"1d_few": 1, "2d": 2, "1d_many": 3, "include_lowest": 4}[case])
    n = 40001
    vals = np.where(rng.random(n) < 0.15, np.nan, rng.lognormal(size=n) * 3)  # one sign: no cancellation in the sums
    a = np.where(rng.random(n) < 0.05, np.nan, np.round(rng.random(n) * 40) / 40)  # many values exactly on breaks
    b = rng.normal(size=n)
    if case == "1d_few":
        bcols, breaks, il = [a], [np.linspace(0, 1, 11)], False
    elif case == "include_lowest":
        bcols, breaks, il = [a], [np.array([0, 0.1, 0.35, 0.5, 0.9])], True
    elif case == "2d":
        bcols, breaks, il = [a, b], [np.linspace(0, 1, 6), np.array([-10, -1, 0, 0.5, 3])], False
    else:  # more bins than a thread block's shared memory holds records for
        bcols, breaks, il = [b], [np.linspace(-3, 3, 1501)], False
    idx, total = _bin_index(bcols, breaks, il)
    dv = gpu_ctx.alloc(n, np.float64).from_host(vals)
    dc = [gpu_ctx.alloc(n, np.float64).from_host(c) for c in bcols]
    got = gpu.bins_summary(gpu_ctx, dv, dc, n, breaks, il)
    exp = _exp_summary_rows([vals[idx == k] for k in range(total)])
    _check_summary_rows(got, exp, case)
    p = [0, 0.1, 0.5, 0.99, 1]
    q, nk = gpu.bins_quantiles(gpu_ctx, dv, dc, n, breaks, p, il)
    assert q.shape == (len(p), total)
    for k in range(total):
        e, en = port.gquantiles(vals[idx == k], p)
        assert nk[k] == en and np.array_equal(q[:, k], e, equal_nan=True), (case, k)
    q0, nk0 = gpu.bins_quantiles(gpu_ctx, dv, dc, 0, breaks, p, il)  # no rows: every bin empty
    assert np.isnan(q0).all() and not nk0.any()
    assert gpu.bins_summary(gpu_ctx, dv, dc, 0, breaks, il)[:, 0].sum() == 0
    with pytest.raises(gpu.MishaGpuError, match="not sorted|not unique"):
        gpu.bins_summary(gpu_ctx, dv, dc[:1], n, [[0, 0.5, 0.2]], il)


def test_percentile_lookup(gpu_ctx):
    """mg_percentile_lookup (global.percentile*, src/TrackVarProcessor.cpp:98-110) against the oracle's val2bin: values on breaks,
    below the first, above the last, NaN; evenly spaced breaks (the ceil shortcut) and ragged ones (bisection)."""
    rng = np.random.default_rng(8)
    for breaks in (np.linspace(0, 2, 201), np.sort(rng.random(50)) * 3):
        bins = np.sort(rng.random(len(breaks)))
        v = np.concatenate([rng.normal(size=5000) + 1, breaks, breaks - 1e-9, breaks + 1e-9, [np.nan, -np.inf, np.inf, breaks[0], breaks[-1]]])
        v = v.astype(np.float32).astype(np.float64)
        d = gpu_ctx.alloc(len(v), np.float64).from_host(v)
        gpu.percentile_lookup(gpu_ctx, d, len(v), breaks, bins)
        gpu_ctx.sync()
        got = d.to_host()
        b = port.val2bin(breaks, np.where(np.isnan(v), 0, v))
        exp = np.where(b >= 0, bins[np.maximum(b, 0)], np.where(v <= breaks[0], bins[0], 1.0))
        exp[np.isnan(v)] = np.nan
        assert np.array_equal(got, exp, equal_nan=True)
    with pytest.raises(gpu.MishaGpuError, match="not sorted|not unique"):
        gpu.percentile_lookup(gpu_ctx, d, 1, [0, 2, 1], [0, 0.5, 1])
