"""BASELINE.json's full sizes (config C3: 3.1 Gbp genome, 20 bp dense track, 10 M intervals, +-5 kb windows) checked through
size-independent properties, plus oracle parity on a sample the oracle finishes in seconds."""
import numpy as np
import pytest

from conftest import assert_close, assert_same
from misha_b200 import gpu, synth
from oracle import port

pytestmark = pytest.mark.gpu
N = 10_000_000
SS, ES = -5000, 5000


@pytest.fixture(scope="module")
def full():
    chroms = synth.genome()
    sizes = [s for _, s in chroms]
    ctx = gpu.Context(sizes)
    dense = gpu.DenseTrack(ctx, 20)
    bins = {}
    for i in range(len(chroms)):
        b = synth.dense_bins(i, sizes[i], 20)
        dense.stage(i, b)
        if i in (0, 13, 23):  # chr1, chr21-ish, chrY kept on the host for the oracle
            bins[i] = b
    counts = synth.interval_counts(chroms, N)
    ch, st, en = [], [], []
    for i in range(len(chroms)):
        s, e = synth.intervals(i, sizes[i], int(counts[i]))
        ch.append(np.full(len(s), i, np.int32)); st.append(s); en.append(e)
    ch, st, en = np.concatenate(ch), np.concatenate(st), np.concatenate(en)
    rows = ctx.rows_upload(ch, st, en)
    yield dict(ctx=ctx, dense=dense, sizes=sizes, bins=bins, chrom=ch, start=st, end=en, rows=rows, chroms=chroms)
    ctx.close()


def test_c3_properties_over_all_rows(full):
    d = full["dense"]
    funcs = ["avg", "sum", "min", "max", "stddev", "size", ("quantile", 0.9), ("quantile", 0.0), ("quantile", 1.0)]
    avg, sm, mn, mx, sd, size, q9, q0, q1 = d.window(full["rows"], funcs, SS, ES)
    assert len(avg) == N
    assert_same(q0, mn, "quantile(0) == min")
    assert_same(q1, mx, "quantile(1) == max")
    ok = ~np.isnan(avg)
    assert ok.mean() > 0.99
    assert np.all(mn[ok] <= q9[ok]) and np.all(q9[ok] <= mx[ok])
    assert np.all(mn[ok] <= avg[ok] * (1 + 1e-6)) and np.all(avg[ok] <= mx[ok] * (1 + 1e-6))
    assert_close(sm[ok], (avg[ok] * size[ok]), 2e-6, "sum == avg * size")
    assert np.all(size[ok] >= 1) and np.all(size <= 530) and np.all(sd[ok & (size > 1)] >= 0)
    # the end-to-end host-buffer entry point returns the same bits as the device-resident path
    host = d.window_host(full["chrom"], full["start"], full["end"], ["avg", "max", ("quantile", 0.9)], SS, ES)
    assert_same(host[0], avg, "host avg"); assert_same(host[1], mx, "host max"); assert_same(host[2], q9, "host q0.9")
    full["cols"] = dict(avg=avg, sum=sm, min=mn, max=mx, stddev=sd, size=size, q9=q9)


def test_c3_index_paths_equal_sweep_on_a_sample(full):
    rng = np.random.default_rng(1)
    idx = np.sort(rng.choice(N, 200_000, replace=False))
    ctx, d = full["ctx"], full["dense"]
    rows = ctx.rows_upload(full["chrom"][idx], full["start"][idx], full["end"][idx])
    funcs = ["avg", "min", "max", "stddev", ("quantile", 0.9), ("quantile", 0.5)]
    a = d.window(rows, funcs, SS, ES)
    ctx.set_option("force_sweep", 1)
    try:
        b = d.window(rows, funcs, SS, ES)
    finally:
        ctx.set_option("force_sweep", 0)
    assert_close(a[0], b[0], 1e-6, "avg"); assert_close(a[3], b[3], 1e-6, "stddev")
    for k in (1, 2, 4, 5):
        assert_same(a[k], b[k], str(funcs[k]))
    if "cols" in full:
        assert_same(a[2], full["cols"]["max"][idx], "sample max == full-run max")


def test_c3_oracle_parity_on_a_sample(full):
    rng = np.random.default_rng(2)
    for cid, b in full["bins"].items():
        m = np.flatnonzero(full["chrom"] == cid)
        pick = np.sort(rng.choice(m, 8000, replace=False))
        s, e = full["start"][pick], full["end"][pick]
        rows = full["ctx"].rows_upload(np.full(len(s), cid, np.int32), s, e)
        got = full["dense"].window(rows, ["avg", "sum", "min", "max", "stddev", ("quantile", 0.9)], SS, ES)
        exp = port.dense_window(b, 20, full["sizes"][cid], s, e, SS, ES, 0.9)
        for k, name in enumerate(("avg", "sum", "min", "max", "stddev", "quantile")):
            if name in ("avg", "sum", "stddev"):
                assert_close(got[k], exp[name], 1e-6, f"chrom {cid} {name}")
            else:
                assert_same(got[k], exp[name], f"chrom {cid} {name}")


def test_c2_gsummary_gdist_whole_genome(full):
    ctx, d, sizes = full["ctx"], full["dense"], full["sizes"]
    sc = np.arange(len(sizes), dtype=np.int32)
    rows = ctx.rows_fixedbin(20, sc, np.zeros(len(sizes), np.int64), np.array(sizes, np.int64))
    assert rows.n == sum(-(-s // 20) for s in sizes)
    part = gpu.summary_init(ctx)
    d.summary_dev(rows, "avg", part)
    tot = gpu.summary_finalize(part.to_host())
    br = np.linspace(0, 100, 101)
    counts = ctx.alloc(100, np.uint64).zero()
    d.hist_dev(rows, "avg", br, counts, include_lowest=True)
    h = counts.to_host()
    # checksum of checksums: every non-NaN bin lies in [0, 100] so the histogram total is the non-NaN count
    assert int(h.sum()) == int(tot[0] - tot[1])
    # per-chromosome parity with the oracle for the chromosomes kept on the host, and additivity over chromosomes
    for cid, b in full["bins"].items():
        r1 = ctx.rows_fixedbin(20, [cid], [0], [sizes[cid]])
        p1 = gpu.summary_init(ctx)
        d.summary_dev(r1, "avg", p1)
        got = gpu.summary_finalize(p1.to_host())
        v = b.astype(np.float64); v[np.isinf(v)] = np.nan
        exp = port.summary(v)
        assert_same(got[:4], exp[:4], f"chrom {cid} counts/min/max")
        assert_close(got[4:], exp[4:], 1e-10, f"chrom {cid} moments")
        c1 = ctx.alloc(100, np.uint64).zero()
        d.hist_dev(r1, "avg", br, c1, include_lowest=True)
        assert np.array_equal(c1.to_host(), port.hist(v, [br], True)), f"chrom {cid} gdist"
