"""k_dense_tile (thread block per tile of rows, everything from shared memory; misha_b200/csrc/dense_tile.cuh) against the
oracle and against the per-row index kernels it replaces (option tile_kernel = 0).

Bar: avg / sum / nearest / size / exists / min / max / quantile BIT-EXACT against the oracle's fresh-window arithmetic
(src/GenomeTrackFixedBin.cpp:881-970, src/StreamPercentiler.h:191-255) -- the tile path re-sums in bin order whenever its
prefix difference could round to another float than the sequential sum. Row sets are chosen to hit the tile path (sorted, dense),
its fallbacks (rows far apart, unsorted rows, tiles across chromosomes, windows longer than a block record) and the seams between.
"""
import numpy as np
import pytest

from conftest import assert_close, assert_same
from misha_b200 import gpu, synth
from oracle import port

pytestmark = pytest.mark.gpu

SIZES = [4_000_000, 2_500_000, 90_000, 1_700_000]
BIN = 20
FUNCS = ["avg", "sum", "min", "max", ("quantile", 0.9), "size", "nearest", "exists", ("quantile", 0.5), ("quantile", 0.0), ("quantile", 1.0)]


def make_bins(kind, cid, size):
    rng = np.random.default_rng([99, cid, ("lognormal", "signed", "nanheavy", "ties").index(kind)])
    n = -(-size // BIN)
    if kind == "lognormal":
        return synth.dense_bins(cid, size, BIN)
    if kind == "signed":  # cancellation: sums near zero against large terms
        v = (rng.standard_normal(n) * 1000).astype(np.float32)
        v[rng.random(n) < 0.02] = np.nan
        v[::977] = np.float32(3e7)
        v[1::977] = np.float32(-3e7)
        return v
    if kind == "nanheavy":
        v = rng.random(n).astype(np.float32)
        v[rng.random(n) < 0.7] = np.nan
        v[1000:4000] = np.nan  # windows without a value
        v[rng.random(n) < 1e-3] = -np.inf
        return v
    if kind == "ties":  # few distinct values: rank ties inside every window
        return rng.integers(0, 4, n).astype(np.float32)
    raise ValueError(kind)


@pytest.fixture(scope="module", params=["lognormal", "signed", "nanheavy", "ties"])
def track(request):
    ctx = gpu.Context(SIZES)
    t = gpu.DenseTrack(ctx, BIN)
    bins = [make_bins(request.param, i, s) for i, s in enumerate(SIZES)]
    for i, b in enumerate(bins):
        t.stage(i, b)
    yield dict(ctx=ctx, t=t, bins=bins, kind=request.param)
    ctx.close()


def oracle_cols(bins, chrom, start, end, ss, es):
    """name -> column over all rows (rows of several chromosomes), from the oracle."""
    n = len(chrom)
    out = {k: np.full(n, np.nan) for k in ("avg", "sum", "min", "max", "size", "q0.9", "q0.5", "q0.0", "q1.0")}
    for cid in np.unique(chrom):
        m = chrom == cid
        e = port.dense_window(bins[cid], BIN, SIZES[cid], start[m], end[m], ss, es, 0.9)
        for k in ("avg", "sum", "min", "max", "size"):
            out[k][m] = e[k]
        out["q0.9"][m] = e["quantile"]
        for p in (0.5, 0.0, 1.0):
            out["q%.1f" % p][m] = port.dense_window(bins[cid], BIN, SIZES[cid], start[m], end[m], ss, es, p, ("quantile",))["quantile"]
    return out


def check(track, chrom, start, end, ss, es, what):
    ctx, t = track["ctx"], track["t"]
    rows = ctx.rows_upload(chrom, start, end)
    ctx.set_option("tile_kernel", 1)  # k_dense_tile: the path with sums to the reference's bits
    got = t.window(rows, FUNCS, ss, es)
    exp = oracle_cols(track["bins"], chrom, start, end, ss, es)
    names = ["avg", "sum", "min", "max", "q0.9", "size", "avg", None, "q0.5", "q0.0", "q1.0"]
    for k, name in enumerate(names):
        if name is None:
            continue
        assert_same(got[k], exp[name], "%s %s [%s] %s" % (track["kind"], what, FUNCS[k], name))
    size = exp["size"]
    okrow = ~np.isnan(got[5])
    assert_same(got[7][okrow], (size[okrow] > 0).astype(float), "exists")
    # the per-row index kernels agree (avg / sum to 1e-6: their prefix difference is not re-summed)
    ctx.set_option("tile_kernel", 0)
    try:
        old = t.window(rows, FUNCS, ss, es)
    finally:
        ctx.set_option("tile_kernel", 1)
    for k in range(len(FUNCS)):
        if FUNCS[k] in ("avg", "sum", "nearest"):
            assert_close(got[k], old[k], 1e-6, "%s old kernel %s" % (what, FUNCS[k]))
        else:
            assert_same(got[k], old[k], "%s old kernel %s" % (what, FUNCS[k]))
    # k_dense_thin (staging only; sums from the chromosome's prefix entries: within 1e-6)
    ctx.set_option("tile_kernel", 4)
    try:
        thin = t.window(rows, FUNCS, ss, es)
        thin_q = t.window(rows, [("quantile", 0.9)], ss, es)[0]
        thin_m = t.window(rows, ["min", "max"], ss, es)
    finally:
        ctx.set_option("tile_kernel", 1)
    for k in range(len(FUNCS)):
        if FUNCS[k] in ("avg", "sum", "nearest"):
            assert_close(thin[k], got[k], 1e-6, "%s thin kernel %s" % (what, FUNCS[k]))
        else:
            assert_same(thin[k], got[k], "%s thin kernel %s" % (what, FUNCS[k]))
    assert_same(thin_q, got[4], "%s thin kernel, quantile alone" % what)
    assert_same(thin_m[0], got[2], "%s thin kernel, min alone" % what)
    assert_same(thin_m[1], got[3], "%s thin kernel, max alone" % what)
    # k_dense_pipe: the same arithmetic in a persistent, multi-buffered thread block (5: two buffers, 6: three): bit for bit
    # k_dense_thin's columns
    for tk in (5, 6):
        ctx.set_option("tile_kernel", tk)
        try:
            pipe = t.window(rows, FUNCS, ss, es)
            pipe_q = t.window(rows, [("quantile", 0.9)], ss, es)[0]
            pipe_m = t.window(rows, ["min", "max"], ss, es)
            pipe_a = t.window(rows, ["avg", "max", ("quantile", 0.9)], ss, es)
        finally:
            ctx.set_option("tile_kernel", 1)
        for k in range(len(FUNCS)):
            assert_same(pipe[k], thin[k], "%s pipe kernel %d %s" % (what, tk, FUNCS[k]))
        assert_same(pipe_q, got[4], "%s pipe kernel %d, quantile alone" % (what, tk))
        assert_same(pipe_m[0], got[2], "%s pipe kernel %d, min alone" % (what, tk))
        assert_same(pipe_m[1], got[3], "%s pipe kernel %d, max alone" % (what, tk))
        assert_same(pipe_a[0], thin[0], "%s pipe kernel %d, avg of avg + max + quantile" % (what, tk))
        assert_same(pipe_a[1], got[3], "%s pipe kernel %d, max of avg + max + quantile" % (what, tk))
        assert_same(pipe_a[2], got[4], "%s pipe kernel %d, quantile of avg + max + quantile" % (what, tk))
    # each function group alone runs another instantiation of the kernel
    for sub in (["avg"], ["max"], ["min", "max"], [("quantile", 0.9)], ["sum", ("quantile", 0.5)], ["avg", "min"]):
        ctx.set_option("tile_kernel", 2)
        try:
            one = t.window(rows, sub, ss, es)
        finally:
            ctx.set_option("tile_kernel", 1)
        for f, col in zip(sub, one):
            assert_same(col, got[FUNCS.index(f)], "%s alone %s" % (what, f))
    ctx.set_option("tile_kernel", 3)  # the default dispatch
    dflt = t.window(rows, FUNCS, ss, es)
    for k in range(len(FUNCS)):
        if FUNCS[k] in ("avg", "sum", "nearest"):
            assert_close(dflt[k], got[k], 1e-6, "%s default dispatch %s" % (what, FUNCS[k]))
        else:
            assert_same(dflt[k], got[k], "%s default dispatch %s" % (what, FUNCS[k]))


def dense_rows(rng, cid, n, lo=100, hi=300):
    s, e = synth.intervals(cid, SIZES[cid], n, lo, hi)
    return np.full(n, cid, np.int32), s, e


def test_c3_shape_sorted_rows(track):
    """The bench's shape: sorted rows ~15 bins apart, +-5 kb windows: every tile takes the shared-memory path."""
    rng = np.random.default_rng(1)
    parts = [dense_rows(rng, cid, SIZES[cid] // 310) for cid in (0, 1, 3)]
    chrom, start, end = (np.concatenate(x) for x in zip(*parts))
    check(track, chrom, start, end, -5000, 5000, "c3 shape")


def test_other_windows(track):
    rng = np.random.default_rng(2)
    chrom, start, end = dense_rows(rng, 0, 6000)
    for ss, es in ((0, 0), (-200, 100), (-10000, 10000), (-20470, 0), (-30000, 30000), (500, 2000), (-60000, 60000)):
        check(track, chrom, start, end, ss, es, "shift %d %d" % (ss, es))


def test_rows_of_every_density(track):
    """From several rows per bin (tiles of a few bins) to rows far apart (tiles that do not fit: the per-row path)."""
    rng = np.random.default_rng(3)
    for n in (200_000, 40_000, 4000, 300):
        s = np.sort(rng.integers(0, SIZES[0] - 500, n))
        e = s + rng.integers(1, 400, n)
        check(track, np.zeros(n, np.int32), s, e, -3000, 3000, "density %d" % n)


def test_unsorted_and_mixed_chromosomes(track):
    rng = np.random.default_rng(4)
    n = 20000
    chrom = rng.integers(0, 4, n).astype(np.int32)
    size = np.array(SIZES)[chrom]
    s = (rng.random(n) * (size - 10)).astype(np.int64)
    e = np.minimum(s + rng.integers(1, 3000, n), size)
    check(track, chrom, s, e, -2000, 2000, "unsorted")
    # sorted by chromosome and start: the tiles at the chromosome seams hold rows of two chromosomes; chromosome 2 is shorter
    # than one tile
    o = np.lexsort((s, chrom))
    check(track, chrom[o], s[o], e[o], -5000, 5000, "sorted, seams")


def test_windows_with_a_handful_of_values():
    """A long run of NaN under the windows, every other bin of the block a number: a window's few values are hundreds of local ranks
    apart, so the (k+1)-th smallest is found by a walk of its own (mg_bw_rank_of_kth), not behind the k-th in pos_of_rank."""
    size = 600_000
    n = -(-size // BIN)
    rng = np.random.default_rng(77)
    v = np.exp(rng.standard_normal(n)).astype(np.float32)
    v[5000:5700] = np.nan   # longer than a +-5 kb window
    v[12000:12490] = np.nan  # a little shorter: windows keep ~ 20 values from either side
    ctx = gpu.Context([size])
    try:
        t = gpu.DenseTrack(ctx, BIN)
        t.stage(0, v)
        s = np.arange(4400 * BIN, 13200 * BIN, 7, dtype=np.int64)  # several rows per bin: sorted, every tile on the shared-memory path
        e = s + 150
        rows = ctx.rows_upload(np.zeros(len(s), np.int32), s, e)
        for q in (0.0, 0.3, 0.5, 0.9, 1.0):
            exp = port.dense_window(v, BIN, size, s, e, -5000, 5000, q, ("quantile", "size"))
            for tk in (3, 1, 0):
                ctx.set_option("tile_kernel", tk)
                got = t.window(rows, [("quantile", q), "size"], -5000, 5000)
                assert_same(got[0], exp["quantile"], "quantile %.1f, tile_kernel %d" % (q, tk))
                assert_same(got[1], exp["size"], "size, tile_kernel %d" % tk)
            few = (exp["size"] > 0) & (exp["size"] < 12)
            assert few.sum() > 40  # the case is there
    finally:
        ctx.close()


def test_edges(track):
    """Windows clamped at either chromosome end, rows without any bin (shifted out of range), one-bin windows."""
    size = SIZES[1]
    s = np.array([0, 0, 10, 19, 20, size - 1, size - 20, size - 21, size - 5000, 100000, 100000, 100020], dtype=np.int64)
    e = np.array([1, 20, 11, 21, 40, size, size, size, size, 100001, 100040, 100021], dtype=np.int64)
    chrom = np.ones(len(s), np.int32)
    for ss, es in ((0, 0), (-5000, 5000), (10**7, 10**7 + 5), (-10**7, -10**7 + 5), (30, -30)):
        check(track, chrom, s, e, ss, es, "edges %d %d" % (ss, es))


def test_fixed_bin_rows(track):
    """The fixed-bin iterator's implicit rows (no row arrays in HBM) under a window."""
    ctx, t = track["ctx"], track["t"]
    rows = ctx.rows_fixedbin(BIN, [0, 1], [1000, 0], [900_000, 500_000])
    ctx.set_option("tile_kernel", 4)
    thin = t.window(rows, ["avg", "max", ("quantile", 0.9), "size"], -2000, 2000)
    ctx.set_option("tile_kernel", 5)
    pipe = t.window(rows, ["avg", "max", ("quantile", 0.9), "size"], -2000, 2000)
    for k in range(4):
        assert_same(pipe[k], thin[k], "fixed-bin rows, pipe kernel column %d" % k)
    ctx.set_option("tile_kernel", 1)
    got = t.window(rows, ["avg", "max", ("quantile", 0.9), "size"], -2000, 2000)
    ctx.set_option("tile_kernel", 3)
    assert_close(thin[0], got[0], 1e-6, "fixed-bin rows, thin kernel avg")
    for k in (1, 2, 3):
        assert_same(thin[k], got[k], "fixed-bin rows, thin kernel column %d" % k)
    ch, st, en = [], [], []
    for cid, a, b in ((0, 1000, 900_000), (1, 0, 500_000)):
        s = np.arange(a // BIN * BIN, b, BIN, dtype=np.int64)
        e = np.minimum(s + BIN, b)
        s = np.maximum(s, a)
        ch.append(np.full(len(s), cid, np.int32)); st.append(s); en.append(e)
    ch, st, en = np.concatenate(ch), np.concatenate(st), np.concatenate(en)
    assert len(ch) == rows.n
    exp = oracle_cols(track["bins"], ch, st, en, -2000, 2000)
    for k, name in enumerate(("avg", "max", "q0.9", "size")):
        assert_same(got[k], exp[name], "fixed-bin rows " + name)
