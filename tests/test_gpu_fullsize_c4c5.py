"""BASELINE.json configs C4 and C5 at their full sizes.

C4: a 5 M-record sparse track / interval set on the 3.09 Gbp genome under the 20 bp fixed-bin iterator (155 M implicit rows):
    nearest / coverage / the gscreen intervals bit-exact against the oracle on whole sampled chromosomes, and the merge kernels equal
    to the per-row search kernels over ALL rows (src/GenomeTrackSparse.cpp:237-340, src/IntervVarProcessor.cpp:242-325,
    src/GenomeTrackScreener.cpp run merge).
C5: 1 M x 500 bp intervals on the 3.09 Gbp packed sequence, 20-mer PSSM, both strands: pwm / pwm.max against the oracle on a
    sample (<= 1e-6 relative), default (float transcendentals) against pwm_strict over all rows (src/PWMScorer.cpp:742-861).
"""
import numpy as np
import pytest

from conftest import assert_close, assert_same
from misha_b200 import gpu, synth
from oracle import port

pytestmark = pytest.mark.gpu
KEEP = (0, 13, 23)  # chromosomes whose records / bytes stay on the host for the oracle (chr1, chr21, chrY)


def fixedbin_rows_host(size, binsize=20):
    s = np.arange(0, size, binsize, dtype=np.int64)
    return s, np.minimum(s + binsize, size)


@pytest.fixture(scope="module")
def c4():
    chroms = synth.genome()
    sizes = [s for _, s in chroms]
    ctx = gpu.Context(sizes)
    sp, iv = gpu.SparseTrack(ctx), gpu.IntervalSet(ctx)
    cnt = synth.interval_counts(chroms, 5_000_000)
    recs, nrec = {}, 0
    for i in range(len(chroms)):
        s, e, v = synth.sparse_records(i, sizes[i], int(cnt[i]))
        sp.stage(i, s, e, v)
        iv.stage(i, s, e)
        nrec += len(s)
        if i in KEEP:
            recs[i] = (s, e, v)
    assert nrec == 5_000_000
    sc = np.arange(len(chroms), dtype=np.int32)
    rows = ctx.rows_fixedbin(20, sc, np.zeros(len(chroms), np.int64), np.array(sizes, np.int64))
    yield dict(ctx=ctx, sp=sp, iv=iv, sizes=sizes, recs=recs, rows=rows)
    ctx.close()


def test_c4_merge_equals_search_over_all_rows(c4):
    ctx, sp, iv, rows = c4["ctx"], c4["sp"], c4["iv"], c4["rows"]
    assert rows.n == sum(-(-s // 20) for s in c4["sizes"]) and rows.n > 154_000_000
    near = sp.window_dev(rows, ["nearest"])[0]
    cov = iv.coverage_dev(rows)
    a_near, a_cov = near.to_host(), cov.to_host()
    ctx.set_option("force_sweep", 1)
    try:
        sp.window_dev(rows, ["nearest"], out=[near])
        iv.coverage_dev(rows, out=cov)
        b_near, b_cov = near.to_host(), cov.to_host()
    finally:
        ctx.set_option("force_sweep", 0)
    assert_same(a_near, b_near, "nearest: merge kernel vs per-row search, all rows")
    assert_same(a_cov, b_cov, "coverage: merge kernel vs per-row search, all rows")
    assert np.all((a_cov >= 0) & (a_cov <= 1))
    c4["near"], c4["cov"] = a_near, a_cov
    near.free(); cov.free()


def test_c4_oracle_parity_on_whole_chromosomes(c4):
    ctx, sp, iv, sizes = c4["ctx"], c4["sp"], c4["iv"], c4["sizes"]
    for cid, (rs, re, rv) in c4["recs"].items():
        r1 = ctx.rows_fixedbin(20, [cid], [0], [sizes[cid]])
        s, e = fixedbin_rows_host(sizes[cid])
        assert r1.n == len(s)
        got = sp.window(r1, ["nearest", "avg", "size"])
        exp = port.sparse_window(rs, re, rv, sizes[cid], s, e, 0, 0, 0.5, funcs=("nearest", "avg", "size"))
        assert_same(got[0], exp["nearest"], f"chrom {cid} nearest")
        assert_same(got[1], exp["avg"], f"chrom {cid} avg")
        assert_same(got[2], exp["size"], f"chrom {cid} size")
        assert_same(iv.coverage(r1), port.coverage(rs, re, sizes[cid], s, e), f"chrom {cid} coverage")
        if "near" in c4:  # the whole-genome launch produced the same bits for this chromosome's rows
            off = sum(-(-sizes[k] // 20) for k in range(cid))
            assert_same(c4["near"][off:off + len(s)], exp["nearest"], f"chrom {cid} nearest inside the whole-genome launch")


def test_c4_gscreen_intervals(c4):
    """gscreen(nearest > 0.5 & coverage > 0.25): device predicate + run merge over all 155 M rows; the intervals of the sampled
    chromosomes are the oracle's merge of the oracle's columns."""
    ctx, sp, iv, rows, sizes = c4["ctx"], c4["sp"], c4["iv"], c4["rows"], c4["sizes"]
    near = sp.window_dev(rows, ["nearest"])[0]
    cov = iv.coverage_dev(rows)
    flag = gpu.screen_compare(ctx, [near, cov], [">", ">"], [0.5, 0.25], False, rows.n)
    gc, gs, ge = gpu.screen_merge(ctx, rows, flag)
    assert len(gc) > 100_000 and np.all(ge > gs)
    host = (ctx.pinned(len(gc), np.int32), ctx.pinned(len(gc), np.int64), ctx.pinned(len(gc), np.int64))
    hc, hs, he = gpu.screen_merge_host(ctx, rows, flag, host)  # device side allocated at the size of the result
    assert np.array_equal(hc, gc) and np.array_equal(hs, gs) and np.array_equal(he, ge)
    order_ok = (gc[1:] > gc[:-1]) | ((gc[1:] == gc[:-1]) & (gs[1:] > ge[:-1]))  # sorted, runs separated by at least one FALSE row
    assert np.all(order_ok)
    for cid, (rs, re, rv) in c4["recs"].items():
        s, e = fixedbin_rows_host(sizes[cid])
        en = port.sparse_window(rs, re, rv, sizes[cid], s, e, 0, 0, 0.5, funcs=("nearest",))["nearest"]
        ec = port.coverage(rs, re, sizes[cid], s, e)
        with np.errstate(invalid="ignore"):
            f = ((en > 0.5) & (ec > 0.25)).astype(np.int8)
        oc, os_, oe = port.screen_merge(np.full(len(s), cid, np.int32), s, e, f)
        m = gc == cid
        assert np.array_equal(gs[m], os_) and np.array_equal(ge[m], oe), f"chrom {cid} gscreen intervals"
    near.free(); cov.free(); flag.free()


@pytest.fixture(scope="module")
def c5():
    chroms = synth.genome()
    sizes = [s for _, s in chroms]
    ctx = gpu.Context(sizes)
    seq = gpu.Sequence(ctx)
    kept = {}
    for i in range(len(chroms)):
        b = synth.sequence(i, sizes[i])
        seq.stage(i, b)
        if i in KEEP:
            kept[i] = b
        del b
    cnt = synth.interval_counts(chroms, 1_000_000)
    ch, st, en = [], [], []
    for i in range(len(chroms)):
        s, e = synth.intervals(i, sizes[i], int(cnt[i]), 500, 500)
        ch.append(np.full(len(s), i, np.int32)); st.append(s); en.append(e)
    ch, st, en = np.concatenate(ch), np.concatenate(st), np.concatenate(en)
    rows = ctx.rows_upload(ch, st, en)
    logp = gpu.pssm_logp(synth.pssm(20), 0.01)
    yield dict(ctx=ctx, seq=seq, sizes=sizes, kept=kept, chrom=ch, start=st, end=en, rows=rows, logp=logp)
    ctx.close()


def test_c5_pwm_full_size(c5):
    ctx, seq, rows, logp = c5["ctx"], c5["seq"], c5["rows"], c5["logp"]
    assert rows.n == 1_000_000
    tot = seq.pwm(rows, logp, bidirect=True, extend=True, mode="pwm")
    mx = seq.pwm(rows, logp, bidirect=True, extend=True, mode="pwm.max")
    pos = seq.pwm(rows, logp, bidirect=True, extend=True, mode="pwm.max.pos")
    ctx.set_option("pwm_strict", 1)
    try:
        tot_s = seq.pwm(rows, logp, bidirect=True, extend=True, mode="pwm")
        mx_s = seq.pwm(rows, logp, bidirect=True, extend=True, mode="pwm.max")
    finally:
        ctx.set_option("pwm_strict", 0)
    assert_close(tot, tot_s, 1e-6, "pwm: float vs strict transcendentals, all rows")
    assert_close(mx, mx_s, 1e-6, "pwm.max (best of the two strands' log-sum-exp): float vs strict, all rows")
    ok = ~np.isnan(tot)
    assert ok.mean() > 0.9
    assert np.all(mx[ok] <= tot[ok] + 1e-4)  # log-sum-exp over the anchors is at least its largest term
    assert np.array_equal(np.isnan(pos), np.isnan(mx))
    rng = np.random.default_rng(3)
    for cid, b in c5["kept"].items():
        m = np.flatnonzero(c5["chrom"] == cid)
        pick = np.sort(rng.choice(m, min(4000, len(m)), replace=False))
        s, e = c5["start"][pick], c5["end"][pick]
        for mode, got in (("pwm", tot), ("pwm.max", mx)):
            exp = port.pwm(b, logp, s, e, bidirect=True, extend=True, mode=mode)
            assert_close(got[pick], exp, 1e-6, f"chrom {cid} {mode} vs oracle")
        assert_close(tot_s[pick], port.pwm(b, logp, s, e, bidirect=True, extend=True, mode="pwm"), 1e-6, f"chrom {cid} strict pwm vs oracle")
        assert_same(pos[pick], port.pwm(b, logp, s, e, bidirect=True, extend=True, mode="pwm.max.pos"), f"chrom {cid} pwm.max.pos")
