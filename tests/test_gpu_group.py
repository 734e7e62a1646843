"""Several GPUs behind ONE host thread (mg_group_*): gextract, gsummary, gdist and gquantiles over a group of devices are what one
device returns -- the reference's forked children + parent merge (src/rdbinterval.cpp:446-564, src/GenomeTrackSummary.cpp:205-216,
src/GenomeTrackDistribution.cpp:131-137, src/GenomeTrackQuantiles.cpp:278) without the fork. Runs on a group of 1 on a one-GPU box
(the routing, gather / scatter and merge code is the same) and on 2 devices when the box has them."""
import numpy as np
import pytest

from conftest import assert_close, assert_same
from misha_b200 import gpu, synth
from oracle import port

pytestmark = pytest.mark.gpu
SIZES = [3_000_000, 1_200_000, 40_000, 2_100_000, 900_000, 1_700_000]
BIN = 20


def ndev():
    import torch
    return torch.cuda.device_count()


@pytest.fixture(scope="module", params=[1, 2])
def setup(request):
    n = request.param
    if n > ndev():
        pytest.skip("needs %d GPUs" % n)
    one = gpu.Context(SIZES)
    t1 = gpu.DenseTrack(one, BIN)
    grp = gpu.Group(SIZES, list(range(n)))
    tg = gpu.GroupDenseTrack(grp, BIN)
    bins = []
    for i, size in enumerate(SIZES):
        b = synth.dense_bins(i, size, BIN)
        bins.append(b)
        t1.stage(i, b)
        tg.stage(i, b)
    yield dict(n=n, one=one, t1=t1, grp=grp, tg=tg, bins=bins)
    tg.free(); grp.close(); one.close()


def test_owners_balance_the_bins(setup):
    grp, n = setup["grp"], setup["n"]
    load = np.zeros(n)
    for i, s in enumerate(SIZES):
        k = grp.owner(i)
        assert 0 <= k < n
        load[k] += s
    assert load.max() <= sum(SIZES) / n + max(SIZES)  # longest-processing-time bound
    if n == 2:
        assert setup["tg"].bytes(0) > 0 and setup["tg"].bytes(1) > 0  # either member holds only its own chromosomes
        assert setup["tg"].bytes(0) + setup["tg"].bytes(1) <= setup["t1"].bytes() * 1.05


def rows_sorted(rng, per_chrom=30000):
    ch, st, en = [], [], []
    for i, size in enumerate(SIZES):
        n = min(per_chrom, size // 400)
        s, e = synth.intervals(i, size, n)
        ch.append(np.full(n, i, np.int32)); st.append(s); en.append(e)
    return np.concatenate(ch), np.concatenate(st), np.concatenate(en)


FUNCS = ["avg", "max", ("quantile", 0.9), "min", "sum", "size"]


def test_gextract_sorted_rows(setup):
    ch, st, en = rows_sorted(np.random.default_rng(1))
    a = setup["t1"].window_host(ch, st, en, FUNCS, -5000, 5000)
    b = setup["tg"].window_host(ch, st, en, FUNCS, -5000, 5000)
    for k, f in enumerate(FUNCS):
        assert_same(b[k], a[k], "group gextract %s" % (f,))
    # and against the oracle on one chromosome
    m = ch == 3
    exp = port.dense_window(setup["bins"][3], BIN, SIZES[3], st[m], en[m], -5000, 5000, 0.9)
    assert_same(b[1][m], exp["max"], "max vs oracle")
    assert_same(b[2][m], exp["quantile"], "quantile vs oracle")
    assert_close(b[0][m], exp["avg"], 1e-6, "avg vs oracle")


def test_gextract_rows_in_any_order(setup):
    """Shuffled rows take the gather / scatter route (more owner changes than the run limit); a few long runs take the direct one."""
    rng = np.random.default_rng(2)
    ch, st, en = rows_sorted(rng, 4000)
    o = rng.permutation(len(ch))
    a = setup["t1"].window_host(ch[o], st[o], en[o], FUNCS, -2000, 3000)
    b = setup["tg"].window_host(ch[o], st[o], en[o], FUNCS, -2000, 3000)
    for k, f in enumerate(FUNCS):
        assert_same(b[k], a[k], "group gextract, shuffled rows, %s" % (f,))
    blocks = np.concatenate([np.flatnonzero(ch == c) for c in (4, 0, 5, 2, 0, 3)])
    a = setup["t1"].window_host(ch[blocks], st[blocks], en[blocks], FUNCS)
    b = setup["tg"].window_host(ch[blocks], st[blocks], en[blocks], FUNCS)
    for k, f in enumerate(FUNCS):
        assert_same(b[k], a[k], "group gextract, chromosome blocks out of order, %s" % (f,))
    with pytest.raises(gpu.MishaGpuError):
        setup["tg"].window_host(np.array([0, 99], np.int32), np.array([0, 0]), np.array([10, 10]), ["avg"])
    empty = setup["tg"].window_host(np.zeros(0, np.int32), np.zeros(0, np.int64), np.zeros(0, np.int64), ["avg"])
    assert len(empty[0]) == 0


def test_gsummary_gdist_gquantiles(setup):
    one, t1, tg = setup["one"], setup["t1"], setup["tg"]
    sc = np.arange(len(SIZES), dtype=np.int32)
    ss, se = np.zeros(len(SIZES), np.int64), np.array(SIZES, np.int64)
    rows = one.rows_fixedbin(BIN, sc, ss, se)
    part = gpu.summary_init(one)
    t1.summary_dev(rows, "avg", part)
    a = gpu.summary_finalize(part.to_host())
    b = gpu.summary_finalize(tg.summary(BIN, sc, ss, se))
    assert_same(b[:4], a[:4], "gsummary counts / min / max")
    assert_close(b[4:], a[4:], 1e-12, "gsummary sum / mean / sd (the merge order differs)")
    allv = np.concatenate(setup["bins"]).astype(np.float64)
    allv[np.isinf(allv)] = np.nan
    assert_same(b[:4], port.summary(allv)[:4], "gsummary vs oracle")
    br = np.linspace(0, 100, 101)
    cnt = one.alloc(100, np.uint64).zero()
    t1.hist_dev(rows, "avg", br, cnt, include_lowest=True)
    h = tg.hist(BIN, sc, ss, se, br, include_lowest=True)
    assert np.array_equal(h, cnt.to_host()), "gdist over the group"
    assert np.array_equal(h, port.hist(allv, [br], True)), "gdist vs oracle"
    # windowed function under a coarser iterator: the generic (prefix / column) paths
    part = gpu.summary_init(one)
    rows2 = one.rows_fixedbin(200, sc, ss, se)
    t1.summary_dev(rows2, "max", part, -300, 300)
    a = gpu.summary_finalize(part.to_host())
    b = gpu.summary_finalize(tg.summary(200, sc, ss, se, "max", -300, 300))
    assert_same(b[:4], a[:4], "gsummary(max, windowed) counts / min / max")
    assert_close(b[4:], a[4:], 1e-12, "gsummary(max, windowed) moments")
    # gquantiles: one exact selection over the members' columns
    p = np.array([0.0, 0.1, 0.5, 0.9, 0.999, 1.0])
    col = t1.window_dev(rows, ["avg"])[0]
    qa, na = gpu.quantiles(one, col, rows.n, p)
    qb, nb = tg.quantiles(BIN, sc, ss, se, p)
    assert na == nb == int(np.sum(~np.isnan(allv)))
    assert_same(qb, qa, "gquantiles over the group")
    assert_same(qb, port.gquantiles(allv, p)[0], "gquantiles vs oracle")
    # a scope that leaves a member without rows
    only = np.array([i for i in range(len(SIZES)) if setup["grp"].owner(i) == 0], np.int32)
    qa2, _ = tg.quantiles(BIN, only, ss[only], se[only], p)
    rows3 = one.rows_fixedbin(BIN, only, ss[only], se[only])
    qb2, _ = gpu.quantiles(one, t1.window_dev(rows3, ["avg"])[0], rows3.n, p)
    assert_same(qa2, qb2, "gquantiles, scope on one member")
