"""Parity of the sparse-track, coverage, PWM and iterator-row kernels with the reference's compiled objects
(golden vectors) and with the oracle on seeded inputs."""
import numpy as np
import pytest

from conftest import CHROMS, assert_close, assert_same, read_seq, read_sparse
from misha_b200 import gpu
from oracle import port

pytestmark = pytest.mark.gpu
RTOL = 1e-6


@pytest.fixture(scope="module")
def sparse_testdb(gpu_ctx):
    t = gpu.SparseTrack(gpu_ctx)
    for cid, (c, _) in enumerate(CHROMS):
        t.stage(cid, *read_sparse("sparse_track", c))
    return t


@pytest.fixture(scope="module")
def seq_testdb(gpu_ctx):
    s = gpu.Sequence(gpu_ctx)
    for cid, (c, _) in enumerate(CHROMS):
        s.stage(cid, read_seq(c))
    return s


def test_sparse_golden(gpu_ctx, sparse_testdb, golden):
    names = ("avg", "min", "max", "stddev", "sum", "nearest")
    for cid, (c, size) in enumerate(CHROMS):
        s, e = golden[f"sparse_{c}_start"], golden[f"sparse_{c}_end"]
        rows = gpu_ctx.rows_upload(np.full(len(s), cid), s, e)
        out = sparse_testdb.window(rows, list(names) + [("quantile", 0.9)])
        for k, name in enumerate(names):
            # record-order accumulation == the reference's order: everything is bit-exact, including avg/sum/stddev
            assert_same(out[k], golden[f"sparse_{c}_{name}"], f"{c} {name}")
        assert_same(out[6], golden[f"sparse_{c}_quantile"], f"{c} quantile")


def test_sparse_shifts_vs_oracle(gpu_ctx, sparse_testdb):
    rng = np.random.default_rng(5)
    for cid, (c, size) in enumerate(CHROMS):
        rs, re, rv = read_sparse("sparse_track", c)
        s = np.sort(rng.integers(0, size - 1, 4000))
        e = np.minimum(s + rng.integers(1, 300, 4000), size)
        rows = gpu_ctx.rows_upload(np.full(len(s), cid), s, e)
        for sshift, eshift in ((0, 0), (-3000, 3000), (500, -100), (-10**7, 10**7)):
            names = ("avg", "sum", "min", "max", "stddev", "nearest", "size")
            got = sparse_testdb.window(rows, list(names) + [("quantile", 0.5)], sshift, eshift)
            exp = port.sparse_window(rs, re, rv, size, s, e, sshift, eshift, 0.5)
            for k, name in enumerate(names):
                assert_same(got[k], exp[name], f"{c} {name} {sshift},{eshift}")
            assert_same(got[7], exp["quantile"], f"{c} quantile {sshift},{eshift}")


def test_sparse_nan_inf_and_empty(gpu_ctx):
    size = 100000
    ctx2 = gpu.Context([size, size])
    try:
        rs = np.array([100, 300, 500, 700, 900]); re = rs + 100
        rv = np.array([1.0, np.nan, np.inf, -np.inf, 2.0], dtype=np.float32)
        t = gpu.SparseTrack(ctx2)
        t.stage(0, rs, re, rv)
        t.stage(1, np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.float32))  # empty chromosome
        s = np.array([0, 150, 250, 350, 450, 0, 650, 950, 2000]); e = np.array([50, 260, 300, 460, 1000, 100000, 700, 960, 3000])
        for chrom in (0, 1):
            rows = ctx2.rows_upload(np.full(len(s), chrom), s, e)
            got = t.window(rows, ["avg", "nearest", "max", "size"])
            if chrom == 0:
                exp = port.sparse_window(rs, re, rv, size, s, e, funcs=("avg", "nearest", "max", "size"))
            else:
                exp = port.sparse_window(rs[:0], re[:0], rv[:0], size, s, e, funcs=("avg", "nearest", "max", "size"))
            for k, name in enumerate(("avg", "nearest", "max", "size")):
                assert_same(got[k], exp[name], f"chrom{chrom} {name}")
        with pytest.raises(gpu.MishaGpuError):
            t2 = gpu.SparseTrack(ctx2)
            t2.stage(0, [10, 5], [20, 8], [1, 2])  # unsorted / overlapping -> reference's load-time error
    finally:
        ctx2.close()


def test_coverage_known_answers(gpu_ctx):
    """tests/testthat/test-vtrack-coverage.R:5-29,48-69 (DB-independent)."""
    iv = gpu.IntervalSet(gpu_ctx)
    # source chr1: [100,200) and [300,400) (already unified)
    iv.stage(0, [100, 300], [200, 400])
    iv.stage(1, [], [])
    iv.stage(2, [], [])

    def cov(starts, ends, chrom=0, **kw):
        rows = gpu_ctx.rows_upload(np.full(len(starts), chrom), starts, ends)
        return iv.coverage(rows, **kw)

    assert_same(cov([0, 100], [100, 200]), [0.0, 1.0], "c(0,1)")
    assert_same(cov([0, 100, 200, 250, 400], [100, 300, 250, 500, 500]), port.coverage([100, 300], [200, 400], 500000, [0, 100, 200, 250, 400], [100, 300, 250, 500, 500]), "mixed")
    assert_same(cov([0, 150, 350], [50, 250, 450]), [0.0, 0.5, 0.5], "half")
    assert_same(cov([0], [500]), [0.4], "0.4")
    assert_same(cov([0], [10], chrom=1), [0.0], "empty chrom")
    # out-of-range shifted window -> NaN (src/IntervVarProcessor.cpp:249-252)
    out = cov([100], [110], sshift=50, eshift=-50)
    assert np.isnan(out[0])


def test_coverage_random_vs_oracle(gpu_ctx):
    rng = np.random.default_rng(9)
    iv = gpu.IntervalSet(gpu_ctx)
    srcs = []
    for cid, (c, size) in enumerate(CHROMS):
        st = rng.integers(0, size - 2000, 3000)
        en = st + rng.integers(1, 1500, 3000)
        _, us, ue = port.unify(np.zeros(3000), st, en, True)
        iv.stage(cid, us, ue)
        srcs.append((us, ue))
    for cid, (c, size) in enumerate(CHROMS):
        s = np.sort(rng.integers(0, size - 1, 5000))
        e = np.minimum(s + rng.integers(1, 5000, 5000), size)
        rows = gpu_ctx.rows_upload(np.full(5000, cid), s, e)
        for sshift, eshift in ((0, 0), (-700, 900), (100, -100)):
            got = iv.coverage(rows, sshift, eshift)
            assert_same(got, port.coverage(srcs[cid][0], srcs[cid][1], size, s, e, sshift, eshift), f"{c} {sshift}")


def test_pwm_golden(gpu_ctx, seq_testdb, golden):
    logp = gpu.pssm_logp(golden["pwm_pssm20"], 0.01)
    assert np.array_equal(logp, golden["pwm_logp20"])  # bit-exact table
    for cid, (c, size) in enumerate(CHROMS):
        s, e = golden[f"pwm_{c}_start"], golden[f"pwm_{c}_end"]
        rows = gpu_ctx.rows_upload(np.full(len(s), cid), s, e)
        for bid in (True, False):
            for ext in (True, False):
                for mode in ("pwm", "pwm.max", "pwm.count"):
                    for strand in (1, -1):
                        if mode == "pwm.max" and strand == -1:
                            continue
                        exp = golden[f"pwm_{c}_{mode}_bid{int(bid)}_ext{int(ext)}_str{strand}"]
                        got = seq_testdb.pwm(rows, logp, bidirect=bid, extend=ext, mode=mode, strand=strand, score_thresh=-25.0)
                        what = f"{c} {mode} bid={bid} ext={ext} strand={strand}"
                        if mode == "pwm.count":
                            # a hit count can move by one when an anchor score sits within 1 ulp of the threshold
                            assert np.array_equal(np.isnan(got), np.isnan(exp)), what
                            d = np.abs(got - exp)[~np.isnan(exp)]
                            assert d.max() <= 1 and (d > 0).mean() < 1e-3, what
                        elif not bid and mode == "pwm.max":
                            assert_same(got, exp, what)  # pure float adds in the reference's order
                        else:
                            assert_close(got, exp, RTOL, what)


def test_pwm_long_rows_and_shift_vs_oracle(gpu_ctx, seq_testdb, golden):
    logp = gpu.pssm_logp(golden["pwm_pssm20"], 0.01)
    seq = read_seq("chr2")
    s = np.array([0, 1000, 250000, 299000, 299990]); e = np.array([5000, 1300, 251000, 300000, 300000])
    rows = gpu_ctx.rows_upload(np.full(len(s), 1), s, e)
    for sshift, eshift in ((0, 0), (-200, 300), (10, -5)):
        for mode in ("pwm", "pwm.max"):
            got = seq_testdb.pwm(rows, logp, mode=mode, sshift=sshift, eshift=eshift)
            exp = port.pwm(seq, logp, s, e, True, True, mode, 1, 0.0, sshift, eshift)
            assert_close(got, exp, RTOL, f"{mode} {sshift},{eshift}")


def test_rows_fixedbin_and_intervals_golden(gpu_ctx, golden):
    # scope is sorted by the entry points before it reaches the iterator
    sc, ss, se = golden["it_scope_chrom"], golden["it_scope_start"], golden["it_scope_end"]
    perm = port.sort_perm(sc, ss)
    for b in (500, 50, 333):
        rows = gpu_ctx.rows_fixedbin(b, sc[perm], ss[perm], se[perm])
        c, s, e, k = rows.coords()
        assert np.array_equal(c, golden[f"it_fixed{b}_chrom"])
        assert np.array_equal(s, golden[f"it_fixed{b}_start"])
        assert np.array_equal(e, golden[f"it_fixed{b}_end"])
        assert np.array_equal(perm[k], golden[f"it_fixed{b}_scope"])  # scope idx = original row (udata)
    ic, is_, ie = port.unify(golden["it_iv_chrom"], golden["it_iv_start"], golden["it_iv_end"], False)
    assert np.array_equal(ic, golden["unify0_chrom"]) and np.array_equal(is_, golden["unify0_start"]) and np.array_equal(ie, golden["unify0_end"])
    sc, ss, se = golden["it_scope2_chrom"], golden["it_scope2_start"], golden["it_scope2_end"]
    perm = port.sort_perm(sc, ss)
    rows = gpu_ctx.rows_intervals(ic, is_, ie, sc[perm], ss[perm], se[perm])
    c, s, e, k = rows.coords()
    assert np.array_equal(c, golden["it_iv_rows_chrom"])
    assert np.array_equal(s, golden["it_iv_rows_start"])
    assert np.array_equal(e, golden["it_iv_rows_end"])
    assert np.array_equal(perm[k], golden["it_iv_rows_scope"])
    # known answer: tests/testthat/test-iterator-overlaps.R:444 -- 2 rows for gintervals(1, 0, 1000) at iterator 500
    assert gpu_ctx.rows_fixedbin(500, [0], [0], [1000]).n == 2


def test_sparse_and_coverage_block_narrowing_modes():
    """Rows of mixed chromosomes in one block (per-thread global search), hulls wider than the shared-memory stage
    (narrowed global range) and ordinary neighbouring rows (staged range) must all agree with the oracle."""
    from misha_b200 import synth
    sizes = [3_000_000, 2_000_000]
    ctx2 = gpu.Context(sizes)
    try:
        rng = np.random.default_rng(17)
        sp, iv = gpu.SparseTrack(ctx2), gpu.IntervalSet(ctx2)
        recs = []
        for cid in range(2):
            s, e, v = synth.sparse_records(cid, sizes[cid], 6000)
            sp.stage(cid, s, e, v)
            iv.stage(cid, s, e)
            recs.append((s, e, v))
        n = 6000
        chrom = rng.integers(0, 2, n).astype(np.int32)  # interleaved chromosomes, unsorted positions
        start = (rng.random(n) * (np.array(sizes)[chrom] - 10)).astype(np.int64)
        end = np.minimum(start + rng.choice([1, 30, 500, 40000, 2_500_000], n), np.array(sizes)[chrom])
        cases = [(chrom, start, end)]
        for cid in range(2):  # sorted single-chromosome rows: staged path; with huge shifts: narrowed-global path
            s = np.sort(rng.integers(0, sizes[cid] - 10, n))
            cases.append((np.full(n, cid, np.int32), s, np.minimum(s + rng.integers(1, 3000, n), sizes[cid])))
        for ch, s, e in cases:
            rows = ctx2.rows_upload(ch, s, e)
            for sshift, eshift in ((0, 0), (-2000, 2000), (-1_500_000, 1_500_000)):
                names = ("avg", "sum", "min", "max", "stddev", "nearest", "size")
                got = sp.window(rows, list(names) + [("quantile", 0.25)], sshift, eshift)
                cov = iv.coverage(rows, sshift, eshift)
                for cid in range(2):
                    m = ch == cid
                    if not m.any():
                        continue
                    rs, re, rv = recs[cid]
                    exp = port.sparse_window(rs, re, rv, sizes[cid], s[m], e[m], sshift, eshift, 0.25)
                    for k, name in enumerate(names):
                        assert_same(got[k][m], exp[name], f"{name} chrom{cid} {sshift}")
                    assert_same(got[7][m], exp["quantile"], f"quantile chrom{cid} {sshift}")
                    assert_same(cov[m], port.coverage(rs, re, sizes[cid], s[m], e[m], sshift, eshift), f"coverage chrom{cid} {sshift}")
    finally:
        ctx2.close()


def test_merge_kernels_vs_search_kernels_and_oracle():
    """The merge kernels (coarse position index + per-lane forward walk) against the per-row search kernels and the
    oracle: unsorted rows, blocks that span chromosomes, records far denser and far sparser than the rows, implicit
    fixed-bin rows."""
    sizes = [300000, 50000, 200000]
    ctx2 = gpu.Context(sizes)
    try:
        rng = np.random.default_rng(17)
        t = gpu.SparseTrack(ctx2)
        iv = gpu.IntervalSet(ctx2)
        recs = []
        for cid, size in enumerate(sizes):
            n = (20000, 7, 0)[cid]  # dense / very sparse / empty
            st = np.sort(rng.choice(size // 10 - 2, n, replace=False)) * 10
            en = st + rng.integers(1, 10, n)
            v = rng.random(n).astype(np.float32)
            v[rng.random(n) < 0.05] = np.nan
            t.stage(cid, st, en, v)
            iv.stage(cid, st, en)
            recs.append((st, en, v))
        n = 30000
        chrom = rng.integers(0, 3, n).astype(np.int32)
        chrom[:10000] = np.sort(chrom[:10000])  # first part sorted by chromosome, the rest fully shuffled
        size_of = np.array(sizes)[chrom]
        s = (rng.random(n) * (size_of - 2)).astype(np.int64)
        s[:10000] = np.concatenate([np.sort(s[:10000][chrom[:10000] == c]) for c in range(3)])
        e = np.minimum(s + rng.choice([1, 5, 20, 300, 5000, 100000], n), size_of)
        rows = ctx2.rows_upload(chrom, s, e)
        names = ["avg", "sum", "min", "max", "nearest", "size", "exists"]
        for sshift, eshift in ((0, 0), (-250, 250)):
            new = t.window(rows, names, sshift, eshift)
            newc = iv.coverage(rows, sshift=sshift, eshift=eshift)
            ctx2.set_option("force_sweep", 1)
            try:
                old = t.window(rows, names, sshift, eshift)
                oldc = iv.coverage(rows, sshift=sshift, eshift=eshift)
            finally:
                ctx2.set_option("force_sweep", 0)
            for k, name in enumerate(names):
                assert_same(new[k], old[k], f"merge vs search {name} shift {sshift}")
            assert_same(newc, oldc, f"coverage merge vs search shift {sshift}")
            for c in range(3):
                m = chrom == c
                onames = [x for x in names if x != "exists"]
                exp = port.sparse_window(*recs[c], sizes[c], s[m], e[m], sshift, eshift, 0.5, funcs=tuple(onames))
                for k, name in enumerate(names):
                    if name == "exists":
                        assert_same(new[k][m], np.where(np.isnan(exp["size"]), np.nan, (exp["size"] > 0).astype(float)), f"chrom{c} exists")
                    else:
                        assert_same(new[k][m], exp[name], f"chrom{c} {name} vs oracle")
                assert_same(newc[m], port.coverage(recs[c][0], recs[c][1], sizes[c], s[m], e[m], sshift, eshift), f"chrom{c} coverage vs oracle")
        # implicit fixed-bin rows over the whole genome
        frows = ctx2.rows_fixedbin(20, np.arange(3, dtype=np.int32), np.zeros(3, np.int64), np.array(sizes, np.int64))
        new = t.window(frows, ["nearest", "avg"])
        newc = iv.coverage(frows)
        ctx2.set_option("force_sweep", 1)
        try:
            old = t.window(frows, ["nearest", "avg"])
            oldc = iv.coverage(frows)
        finally:
            ctx2.set_option("force_sweep", 0)
        assert_same(new[0], old[0], "fixed-bin nearest")
        assert_same(new[1], old[1], "fixed-bin avg")
        assert_same(newc, oldc, "fixed-bin coverage")
    finally:
        ctx2.close()


def test_pwm_float_and_strict_transcendentals_agree(gpu_ctx, seq_testdb, golden):
    """Default: float expf / logf per anchor; pwm_strict: double exp / log rounded to float. Both within the 1e-6 bar of each
    other (and of the reference, test_pwm_golden); pure-float modes are identical."""
    logp = gpu.pssm_logp(golden["pwm_pssm20"], 0.01)
    rng = np.random.default_rng(23)
    size = CHROMS[0][1]
    s = np.sort(rng.integers(0, size - 600, 3000))
    e = s + rng.integers(1, 600, 3000)
    rows = gpu_ctx.rows_upload(np.zeros(3000, np.int32), s, e)
    res = {}
    for strict in (0, 1):
        gpu_ctx.set_option("pwm_strict", strict)
        try:
            for bid in (True, False):
                for mode in ("pwm", "pwm.max"):
                    res[(strict, bid, mode)] = seq_testdb.pwm(rows, logp, bidirect=bid, mode=mode)
        finally:
            gpu_ctx.set_option("pwm_strict", 0)
    for bid in (True, False):
        for mode in ("pwm", "pwm.max"):
            a, b = res[(0, bid, mode)], res[(1, bid, mode)]
            if not bid and mode == "pwm.max":
                assert_same(a, b, "single-strand max has no transcendental")
            else:
                assert_close(a, b, RTOL, f"float vs strict bid={bid} {mode}")


def test_sparse_order_dependent_functions(gpu_ctx, sparse_testdb):
    """lse / first / last / positions on a sparse track (record order, positions = record starts) against the reference's
    compiled reader (ref_vectors_ext.npz) and the oracle with shifts; lse accumulates pairwise in float in record order
    like the reference, so it differs only by the last bits of expf / logf."""
    import os
    from conftest import GOLDEN
    ext = ("lse", "first", "last", "first.pos", "last.pos", "min.pos", "max.pos")
    g = np.load(os.path.join(GOLDEN, "ref_vectors_ext.npz"))
    for cid, (c, size) in enumerate(CHROMS):
        s, e = g[f"{c}_start"], g[f"{c}_end"]
        rows = gpu_ctx.rows_upload(np.full(len(s), cid), s, e)
        got = sparse_testdb.window(rows, list(ext))
        for k, name in enumerate(ext):
            if name == "lse":
                np.testing.assert_allclose(got[k], g[f"sparse_{c}_lse"], rtol=1e-6, atol=1e-6, equal_nan=True)
            else:
                assert_same(got[k], g[f"sparse_{c}_{name}"], f"{c} {name}")
        rs, re, rv = read_sparse("sparse_track", c)
        for sshift, eshift in ((-2500, 2500), (300, 800)):
            got = sparse_testdb.window(rows, ["avg", ("first.pos", 1), ("max.pos", 1), "last", ("min.pos", 0)], sshift, eshift)
            exp = port.sparse_window_ext(rs, re, rv, size, s, e, sshift, eshift)
            avg = port.sparse_window(rs, re, rv, size, s, e, sshift, eshift, 0.5, ("avg",))["avg"]
            ws = np.maximum(s + sshift, 0).astype(np.float64)
            ok = (s + sshift < size) & (np.minimum(e + eshift, size) > np.maximum(s + sshift, 0))
            assert_same(got[0], avg, f"{c} avg next to ext")
            assert_same(got[1][ok], (exp["first.pos"] - ws)[ok], f"{c} first.pos relative")
            assert_same(got[2][ok], (exp["max.pos"] - ws)[ok], f"{c} max.pos relative")
            assert_same(got[3], exp["last"], f"{c} last")
            assert_same(got[4], exp["min.pos"], f"{c} min.pos")


def test_pwm_max_pos_golden_and_oracle(gpu_ctx, seq_testdb, golden):
    """pwm.max.pos (PWMScorer MAX_LIKELIHOOD_POS -> DnaPSSM::max_like_match): best strand per anchor, first strictly greater
    anchor in target order, N scored as the row average, 1-based signed position in float -- bit for bit against the
    reference's compiled scorer (ref_vectors_ext.npz) and the oracle with shifts. Rows clipped below the motif length with
    both strands are the reference's uninitialised-sign case (NaN here, excluded)."""
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "ref_vectors_ext.npz"))
    logp = gpu.pssm_logp(golden["pwm_pssm20"], 0.01)
    L = logp.shape[0]
    for cid, (c, size) in enumerate(CHROMS):
        s, e = golden[f"pwm_{c}_start"], golden[f"pwm_{c}_end"]
        rows = gpu_ctx.rows_upload(np.full(len(s), cid), s, e)
        seq = read_seq(c)
        for bid in (True, False):
            for ext in (True, False):
                for strand in (1, -1):
                    exp = g[f"pwm_{c}_pwm.max.pos_bid{int(bid)}_ext{int(ext)}_str{strand}"]
                    got = seq_testdb.pwm(rows, logp, bidirect=bid, extend=ext, mode="pwm.max.pos", strand=strand)
                    defined = ~(bid & ext & (np.minimum(e + L - 1, size) - s < L))
                    assert defined.sum() >= len(s) - 2
                    assert_same(got[defined], exp[defined], f"{c} pwm.max.pos bid={bid} ext={ext} strand={strand}")
                    assert np.all(np.isnan(got[~defined]))
                    for sshift, eshift in ((-150, 90), (7, 300)):
                        got = seq_testdb.pwm(rows, logp, bidirect=bid, extend=ext, mode="pwm.max.pos", strand=strand, sshift=sshift, eshift=eshift)
                        exp = port.pwm(seq, logp, s, e, bid, ext, "pwm.max.pos", strand, 0.0, sshift, eshift)
                        assert_same(got, exp, f"{c} pwm.max.pos shifted bid={bid} ext={ext} strand={strand}")
    # long rows (several 128-anchor chunks per warp), ties between equal anchors on a low-complexity PSSM
    flat = gpu.pssm_logp(np.tile([[0.7, 0.1, 0.1, 0.1]], (3, 1)), 0.01)
    seq = read_seq("chr2")
    s = np.array([0, 1000, 250000, 298000]); e = np.array([5000, 3300, 251000, 300000])
    rows = gpu_ctx.rows_upload(np.full(len(s), 1), s, e)
    for lp in (logp, flat):
        for bid in (True, False):
            for strand in (1, -1):
                got = seq_testdb.pwm(rows, lp, bidirect=bid, mode="pwm.max.pos", strand=strand)
                assert_same(got, port.pwm(seq, lp, s, e, bid, True, "pwm.max.pos", strand), f"long rows bid={bid} strand={strand}")


def test_host_buffer_pipeline_sparse_coverage_pwm(gpu_ctx, sparse_testdb, seq_testdb, golden):
    """mg_h_sparse_window / mg_h_coverage / mg_h_pwm (host rows in, host columns out through the three-slot copy pipeline)
    equal the device-resident calls bit for bit, over several chunks and a ragged last one."""
    rng = np.random.default_rng(64)
    n = 5000
    chrom = np.sort(rng.integers(0, 3, n)).astype(np.int32)
    sizes = np.array([z for _, z in CHROMS])
    s = np.array([rng.integers(0, sizes[c] - 700) for c in chrom])
    order = np.lexsort((s, chrom))
    chrom, s = chrom[order], s[order]
    e = s + rng.integers(1, 600, n)
    rows = gpu_ctx.rows_upload(chrom, s, e)
    gpu_ctx.set_option("pipe_chunk_rows", 1024)
    try:
        funcs = ["avg", "nearest", "max", ("quantile", 0.5), ("max.pos", 1)]
        dev = sparse_testdb.window(rows, funcs, -300, 450)
        host = sparse_testdb.window_host(chrom, s, e, funcs, -300, 450)
        for k in range(len(funcs)):
            assert_same(host[k], dev[k], f"sparse host pipeline {funcs[k]}")
        iv = gpu.IntervalSet(gpu_ctx)
        for cid, (c, size) in enumerate(CHROMS):
            st = np.sort(rng.integers(0, size - 500, 300))
            en = st + rng.integers(1, 400, 300)
            us, ue = [], []
            for a, b in zip(st, en):
                if us and a <= ue[-1]:
                    ue[-1] = max(ue[-1], b)
                else:
                    us.append(a); ue.append(b)
            iv.stage(cid, np.array(us), np.array(ue))
        assert_same(iv.coverage_host(chrom, s, e, -100, 100), iv.coverage(rows, -100, 100), "coverage host pipeline")
        iv.free()
        logp = gpu.pssm_logp(golden["pwm_pssm20"], 0.01)
        for mode in ("pwm", "pwm.max", "pwm.max.pos", "pwm.count"):
            got = seq_testdb.pwm_host(chrom, s, e, logp, mode=mode, score_thresh=-25.0, sshift=-10, eshift=10)
            assert_same(got, seq_testdb.pwm(rows, logp, mode=mode, score_thresh=-25.0, sshift=-10, eshift=10), f"pwm host pipeline {mode}")
        assert len(sparse_testdb.window_host(chrom[:0], s[:0], e[:0], ["avg"])[0]) == 0
    finally:
        gpu_ctx.set_option("pipe_chunk_rows", 1 << 19)


def test_coverage_rasterised_grid_blocks():
    """Coverage over fixed-bin rows rasterises the source intervals into the block's rows (difference array for completely
    covered rows + atomics for the partly covered ones). Against the oracle and the per-row cursor path (cov_raster = 0):
    bin sizes that do and do not divide the coordinates, windows of 1..8 bins and longer (the cursor path takes over),
    intervals longer than a whole block, 1 bp intervals, touching intervals, scopes that start off the bin grid."""
    sizes = [700000, 90000]
    ctx2 = gpu.Context(sizes)
    try:
        rng = np.random.default_rng(2027)
        iv = gpu.IntervalSet(ctx2)
        srcs = []
        for cid, size in enumerate(sizes):
            cuts = np.unique(rng.integers(0, size, 6000))
            st, en = cuts[0:-1:2].copy(), cuts[1::2].copy()
            m = min(len(st), len(en))
            st, en = st[:m], en[:m]
            en[::7] = st[::7] + 1                          # 1 bp intervals
            st[5::11] = en[4::11][:len(st[5::11])]         # touching neighbours
            keep = en > st
            st, en = st[keep], en[keep]
            if cid == 0:                                   # one interval longer than several blocks of rows, one ending at the chromosome end
                inside = (st >= 200000) & (st < 420000)
                st, en = st[~inside], en[~inside]
                st = np.sort(np.concatenate([st, [200003]])); en = np.sort(np.concatenate([en[en <= 200003] if False else en, [419999]]))
            # re-unify: sorted, disjoint
            order = np.argsort(st)
            st, en = st[order], np.maximum.accumulate(en[order])
            us, ue = [], []
            for a, b in zip(st, en):
                if us and a <= ue[-1]:
                    ue[-1] = max(ue[-1], b)
                else:
                    us.append(a); ue.append(b)
            st, en = np.array(us, np.int64), np.array(ue, np.int64)
            iv.stage(cid, st, en)
            srcs.append((st, en))
        scope_c = np.array([0, 0, 1], np.int32)
        scope_s = np.array([0, 350007, 13], np.int64)
        scope_e = np.array([333333, 700000, 90000], np.int64)
        for binsize in (20, 7, 50, 1000):
            rows = ctx2.rows_fixedbin(binsize, scope_c, scope_s, scope_e)
            c, s, e, _ = rows.coords()
            for sshift, eshift in ((0, 0), (-binsize, binsize), (3, 2 * binsize + 5), (-4 * binsize, 3 * binsize), (-20 * binsize, 20 * binsize), (5, -3)):
                got = iv.coverage(rows, sshift, eshift)                      # 16 rows per thread (few records per block)
                ctx2.set_option("merge_rows", 8)
                try:
                    assert_same(iv.coverage(rows, sshift, eshift), got, f"8 vs 16 rows per thread, bin {binsize} shift {sshift},{eshift}")
                finally:
                    ctx2.set_option("merge_rows", 0)
                ctx2.set_option("cov_raster", 0)
                try:
                    cur = iv.coverage(rows, sshift, eshift)
                finally:
                    ctx2.set_option("cov_raster", 1)
                assert_same(got, cur, f"raster vs cursor bin {binsize} shift {sshift},{eshift}")
                for cid in range(2):
                    m = c == cid
                    assert_same(got[m], port.coverage(srcs[cid][0], srcs[cid][1], sizes[cid], s[m], e[m], sshift, eshift),
                                f"raster vs oracle chrom{cid} bin {binsize} shift {sshift},{eshift}")
    finally:
        ctx2.close()


def test_sparse_merge_rows_per_thread_fixed_bin():
    """Fixed-bin rows over a sparse track: the merge kernel at 16 rows per thread (chosen when a block is expected to hold few
    records) and at 8 give the same columns, equal to the oracle; scope intervals that start off the grid, shifts, a
    chromosome without records."""
    sizes = [900000, 120000, 50000]
    ctx2 = gpu.Context(sizes)
    try:
        rng = np.random.default_rng(515)
        t = gpu.SparseTrack(ctx2)
        recs = []
        for cid, size in enumerate(sizes):
            n = (2500, 300, 0)[cid]
            st = np.sort(rng.choice(size // 40 - 2, n, replace=False)) * 40
            en = st + rng.integers(1, 40, n)
            v = rng.random(n).astype(np.float32)
            v[rng.random(n) < 0.04] = np.nan
            t.stage(cid, st, en, v)
            recs.append((st, en, v))
        scope_c = np.array([0, 0, 1, 2], np.int32)
        scope_s = np.array([11, 500003, 0, 7], np.int64)
        scope_e = np.array([480000, 900000, 120000, 50000], np.int64)
        names = ["nearest", "avg", "max", "size"]
        for binsize in (20, 33):
            rows = ctx2.rows_fixedbin(binsize, scope_c, scope_s, scope_e)
            c, s, e, _ = rows.coords()
            for sshift, eshift in ((0, 0), (-90, 140)):
                got = t.window(rows, names, sshift, eshift)
                one = t.window(rows, ["nearest"], sshift, eshift)[0]
                ctx2.set_option("merge_rows", 8)
                try:
                    ref8 = t.window(rows, names, sshift, eshift)
                finally:
                    ctx2.set_option("merge_rows", 0)
                assert_same(one, got[0], "single-function kernel")
                for k, name in enumerate(names):
                    assert_same(got[k], ref8[k], f"{name}: 16 vs 8 rows per thread, bin {binsize}")
                    for cid in range(3):
                        m = c == cid
                        exp = port.sparse_window(*recs[cid], sizes[cid], s[m], e[m], sshift, eshift, 0.5, funcs=(name,))[name]
                        assert_same(got[k][m], exp, f"{name} chrom{cid} bin {binsize} shift {sshift}")
    finally:
        ctx2.close()


def test_merge_kernels_long_scope_device_grid_path():
    """More than 32 scope intervals: the regular-grid blocks are not worked out on the host (GridTable.n = -1) and every thread
    block decides on the device (mg_sf_grid_block), the scope search reads global memory. Same results as the oracle."""
    sizes = [2_000_000]
    ctx2 = gpu.Context(sizes)
    try:
        rng = np.random.default_rng(90210)
        n = 6000
        st = np.sort(rng.choice(sizes[0] // 50 - 2, n, replace=False)) * 50
        en = st + rng.integers(1, 50, n)
        v = rng.random(n).astype(np.float32)
        t = gpu.SparseTrack(ctx2); t.stage(0, st, en, v)
        iv = gpu.IntervalSet(ctx2); iv.stage(0, st, en)
        edges = np.sort(rng.choice(np.arange(1, 1999), 79, replace=False)) * 1000 + rng.integers(0, 17, 79)
        ss, se = np.concatenate([[3], edges[1::2]]), np.concatenate([edges[0::2], [sizes[0]]])[:len(np.concatenate([[3], edges[1::2]]))]
        keep = se > ss
        ss, se = ss[keep].astype(np.int64), se[keep].astype(np.int64)
        assert len(ss) > 32
        rows = ctx2.rows_fixedbin(20, np.zeros(len(ss), np.int32), ss, se)
        c, s, e, _ = rows.coords()
        for sshift, eshift in ((0, 0), (-30, 45)):
            assert_same(iv.coverage(rows, sshift, eshift), port.coverage(st, en, sizes[0], s, e, sshift, eshift), f"coverage {sshift}")
            got = t.window(rows, ["nearest", "avg"], sshift, eshift)
            exp = port.sparse_window(st, en, v, sizes[0], s, e, sshift, eshift, 0.5, funcs=("nearest", "avg"))
            assert_same(got[0], exp["nearest"], f"nearest {sshift}")
            assert_same(got[1], exp["avg"], f"avg {sshift}")
    finally:
        ctx2.close()


def test_kmer_golden_and_oracle(gpu_ctx, seq_testdb):
    """mg_kmer (kmer.count / kmer.frac) bit for bit against the reference's compiled KmerCounter (tests/golden/kmer_vectors.npz, made by
    make_kmer_fixtures.py) -- every k-mer, strand and extend combination, chromosome-edge and shorter-than-kmer rows -- and against
    the oracle restatement with iterator shifts."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "kmer_vectors.npz"))
    for cid, (c, size) in enumerate(CHROMS):
        s, e = g[f"{c}_start"], g[f"{c}_end"]
        rows = gpu_ctx.rows_upload(np.full(len(s), cid), s, e)
        for ki, kmer in enumerate(g["kmers"]):
            for mode in ("kmer.count", "kmer.frac"):
                for ext in (True, False):
                    for strand in (0, 1, -1):
                        got = seq_testdb.kmer_dev(rows, str(kmer), mode=mode, extend=ext, strand=strand).to_host()
                        assert_same(got, g[f"{c}_k{ki}_{mode}_ext{int(ext)}_str{strand}"], f"{c} {kmer} {mode} ext={ext} strand={strand}")
        seq = read_seq(c)
        for sshift, eshift in ((-300, 200), (5, -5), (0, 40)):
            for kmer, mode, strand in (("cg", "kmer.count", 0), ("TTAGG", "kmer.frac", -1)):
                got = seq_testdb.kmer_dev(rows, kmer, mode=mode, strand=strand, sshift=sshift, eshift=eshift).to_host()
                assert_same(got, port.kmer(seq, kmer, s, e, mode, True, strand, sshift, eshift), f"{c} {kmer} shift {sshift} {eshift}")
    for bad, msg in (("", "empty"), ("ACGN", "A, C, G, T"), ("A" * 33, "at most 32")):
        with pytest.raises(gpu.MishaGpuError, match=msg):
            seq_testdb.kmer_dev(rows, bad)


def test_masked_golden_and_oracle(gpu_ctx, seq_testdb):
    """mg_masked (masked.count / masked.frac) bit for bit against the reference's compiled MaskedBpCounter (kmer_vectors.npz) and,
    with iterator shifts and long windows, against the oracle restatement."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "kmer_vectors.npz"))
    for cid, (c, size) in enumerate(CHROMS):
        s, e = g[f"{c}_start"], g[f"{c}_end"]
        rows = gpu_ctx.rows_upload(np.full(len(s), cid), s, e)
        seq = read_seq(c)
        for mode in ("masked.count", "masked.frac"):
            assert_same(seq_testdb.masked_dev(rows, mode=mode).to_host(), g[f"{c}_{mode}"], f"{c} {mode}")
            for sshift, eshift in ((-3000, 4000), (7, -3), (0, 31)):
                got = seq_testdb.masked_dev(rows, mode=mode, sshift=sshift, eshift=eshift).to_host()
                assert_same(got, port.masked(seq, s, e, mode, sshift, eshift), f"{c} {mode} shift {sshift} {eshift}")


def test_pwm_spatial_weights(gpu_ctx, seq_testdb):
    """pwm / pwm.max / pwm.count under spatial weights (mg_pwm_spatial) against the compiled reference's PWMScorer with spat_factor set
    (tests/golden/pwm_spat_vectors.npz) and the restatement under shifts."""
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "pwm_spat_vectors.npz"))
    logp = gpu.pssm_logp(g["pssm"], float(g["prior"]))
    spat, sbin, th = g["spat"], int(g["spat_bin"]), float(g["thresh"])
    for cid, (c, size) in enumerate(CHROMS):
        if c + "_start" not in g:
            continue
        s, e = g[c + "_start"], g[c + "_end"]
        rows = gpu_ctx.rows_upload(np.full(len(s), cid, np.int32), s, e)
        for bid in (1, 0):
            for strand in (1, -1):
                for ext in (1, 0):
                    for mode in ("pwm", "pwm.max", "pwm.count"):
                        got = seq_testdb.pwm(rows, logp, bidirect=bool(bid), extend=bool(ext), mode=mode, strand=strand, score_thresh=th, spat_factor=spat,
                                             spat_bin=sbin)
                        exp = g["%s_%s_b%d_s%d_e%d" % (c, mode, bid, strand, ext)]
                        if mode == "pwm.count":
                            assert_same(got, exp, "%s %s b%d s%d e%d" % (c, mode, bid, strand, ext))
                        else:
                            assert_close(got, exp, RTOL, "%s %s b%d s%d e%d" % (c, mode, bid, strand, ext))
        seq = read_seq(c)
        for ss, es in ((-50, 80), (30, 10)):
            got = seq_testdb.pwm(rows, logp, mode="pwm", sshift=ss, eshift=es, spat_factor=spat, spat_bin=sbin)
            assert_close(got, port.pwm_spat(seq, logp, s, e, spat, sbin, mode="pwm", sshift=ss, eshift=es), RTOL, "%s spatial pwm, shifts" % c)
    with pytest.raises(gpu.MishaGpuError):
        seq_testdb.pwm(rows, logp, mode="pwm.max.pos", spat_factor=spat, spat_bin=sbin)
    # no weights given: the plain scorer, bit for bit
    assert_same(seq_testdb.pwm(rows, logp, spat_factor=None), seq_testdb.pwm(rows, logp), "no weights == mg_pwm")
