"""The R-interface mirror (misha_b200.api) on the reference's bundled example DB: these read like the reference's own
testthat files (test-gdist.R, test-vtrack-coverage.R, test-iterator-overlaps.R, test-vtrack.R, test-pwm.R) and check
against the known answers they hold and against the oracle."""
import numpy as np
import pytest

from conftest import CHROMS, TESTDB, assert_close, assert_same, read_dense, read_seq, read_sparse
from misha_b200 import api, gpu
from oracle import port

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def db():
    d = api.gdb_init(TESTDB)
    yield d
    d.close()


def test_gdb_init_and_chromosomes(db):
    assert db.chrom_names == ["chr1", "chr2", "chrX"] and db.chrom_sizes.tolist() == [500000, 300000, 200000]
    assert set(db.gtrack_ls()) >= {"dense_track", "sparse_track"}
    all_iv = db.gintervals_all()
    assert all_iv["end"].tolist() == [500000, 300000, 200000]
    with pytest.raises(api.MishaError):
        db.gintervals(7, 0, 100)
    with pytest.raises(api.MishaError):
        db.gintervals(1, 100, 50)


def test_config_c1_gextract_dense_iterator_1000(db):
    """BASELINE.json configs[0]: gextract("dense_track", gintervals.all(), iterator = 1000) -> 1000 rows."""
    r = db.gextract("dense_track", intervals=db.gintervals_all(), iterator=1000)
    assert len(r["start"]) == 1000 and list(r) == ["chrom", "start", "end", "dense_track", "intervalID"]
    assert r["intervalID"].tolist() == [1] * 500 + [2] * 300 + [3] * 200
    off = 0
    for cid, (c, size) in enumerate(CHROMS):
        bs, bins = read_dense("dense_track", c)
        n = size // 1000
        assert np.array_equal(r["start"][off:off + n], np.arange(n) * 1000) and np.all(r["chrom"][off:off + n] == c)
        exp = port.dense_window(bins, bs, size, r["start"][off:off + n], r["end"][off:off + n], funcs=("avg",))["avg"]
        assert_close(r["dense_track"][off:off + n], exp, 1e-6, c)
        off += n


def test_gdist_known_answers(db):
    """tests/testthat/test-gdist.R:69-72."""
    d = db.gdist("dense_track", [0, 0.2, 0.5, 1], intervals=db.gintervals(1, 0, 500000))
    assert d.sum() == 6546
    d = db.gdist("dense_track", [0, 0.2, 0.5, 1], intervals=db.gintervals(1, 0, 50000))
    assert d.sum() == 901
    with pytest.raises(api.MishaError, match="not sorted|not unique"):
        db.gdist("dense_track", [0, 0.5, 0.2], intervals=db.gintervals(1, 0, 50000))
    # 2-D: dense x sparse over a 200 bp iterator
    d2 = db.gdist("dense_track", [0, 0.1, 0.2, 1], "sparse_track", [0, 0.3, 0.5, 1], intervals=db.gintervals(1), iterator=200)
    r = db.gextract("dense_track", "sparse_track", intervals=db.gintervals(1), iterator=200)
    exp = port.hist(np.stack([r["dense_track"], r["sparse_track"]], axis=1), [[0, 0.1, 0.2, 1], [0, 0.3, 0.5, 1]])
    assert np.array_equal(d2, exp.astype(np.float64))


def test_vtrack_coverage_known_answers(db):
    """tests/testthat/test-vtrack-coverage.R:5-29,48-69."""
    src = db.gintervals([1, 1], [100, 300], [200, 400])
    db.gvtrack_create("cov", src, "coverage")
    r = db.gextract("cov", intervals=db.gintervals([1, 1], [0, 100], [100, 200]), iterator=db.gintervals([1, 1], [0, 100], [100, 200]))
    assert r["cov"].tolist() == [0.0, 1.0]
    it = db.gintervals([1] * 5, [0, 100, 200, 300, 400], [100, 200, 300, 400, 500])
    assert db.gextract("cov", intervals=it, iterator=it)["cov"].tolist() == [0, 1, 0, 1, 0]
    it = db.gintervals([1] * 3, [0, 150, 350], [50, 250, 450])
    assert db.gextract("cov", intervals=it, iterator=it)["cov"].tolist() == [0.0, 0.5, 0.5]
    # overlapping sources are unified before coverage is computed
    db.gvtrack_create("cov2", db.gintervals([1, 1, 1], [100, 150, 300], [200, 250, 400]), "coverage")
    it = db.gintervals(1, 0, 500)
    assert db.gextract("cov2", intervals=it, iterator=it)["cov2"].tolist() == [0.5]
    db.gvtrack_rm("cov"); db.gvtrack_rm("cov2")


def test_iterator_overlaps_and_interval_ids(db):
    """tests/testthat/test-iterator-overlaps.R:444 and the intervalID contract (1-based row of the scope data frame)."""
    r = db.gextract("dense_track", intervals=db.gintervals(1, 0, 1000), iterator=500)
    assert len(r["start"]) == 2 and r["start"].tolist() == [0, 500] and r["end"].tolist() == [500, 1000]
    scope = db.gintervals(["X", 1, 2, 1], [100, 5000, 30, 120], [260, 5100, 75, 4990])  # unsorted, overlapping scope rows
    r = db.gextract("dense_track", intervals=scope, iterator=50)
    c, s, e, k = port.fixedbin_rows(50, *[a[np.lexsort((scope["start"], db._cids(scope)))] for a in (db._cids(scope), scope["start"], scope["end"])])
    order = np.lexsort((scope["start"], db._cids(scope)))
    assert np.array_equal(r["start"], s) and np.array_equal(r["end"], e) and np.array_equal(r["intervalID"], order[k] + 1)


def test_vtracks_on_dense_with_shifts(db):
    """tests/testthat/test-vtrack.R style: min / max / stddev / quantile / sum with an iterator modifier."""
    for func, prm in (("avg", None), ("max", None), ("min", None), ("sum", None), ("stddev", None), ("quantile", 0.9), ("quantile", 0.5)):
        db.gvtrack_create("v", "dense_track", func, prm)
        db.gvtrack_iterator("v", sshift=-130, eshift=220)
        iv = db.gintervals([1, 1, 2, "X"], [0, 250000, 1000, 199000], [2000, 253000, 9000, 200000])
        r = db.gextract("v", intervals=iv, iterator=100)
        off = 0
        for cid, (c, size) in enumerate(CHROMS):
            m = r["chrom"] == c
            bs, bins = read_dense("dense_track", c)
            exp = port.dense_window(bins, bs, size, r["start"][m], r["end"][m], -130, 220, prm or 0.5, (func,))[func]
            if func in ("avg", "sum", "stddev"):
                assert_close(r["v"][m], exp, 1e-6, f"{func} {c}")
            else:
                assert_same(r["v"][m], exp, f"{func} {c}")
    db.gvtrack_rm("v")
    with pytest.raises(api.MishaError):
        db.gvtrack_create("q", "dense_track", "quantile")  # percentile missing
    with pytest.raises(api.MishaError):
        db.gvtrack_create("q", "no_such_track", "avg")


def test_vtrack_order_dependent_functions(db):
    """gvtrack.create with lse / first / last / *.pos.abs / *.pos.relative (R/vtrack.R; src/TrackVarProcessor.cpp:113-173)."""
    iv = db.gintervals([1, 2, "X"], [0, 1000, 150000], [30000, 9000, 200000])
    for func, col, rel in (("lse", "lse", False), ("first", "first", False), ("last", "last", False), ("max.pos.abs", "max.pos", False),
                           ("max.pos.relative", "max.pos", True), ("min.pos.relative", "min.pos", True), ("first.pos.abs", "first.pos", False),
                           ("last.pos.relative", "last.pos", True)):
        for track in ("dense_track", "sparse_track"):
            db.gvtrack_create("v", track, func)
            db.gvtrack_iterator("v", sshift=-130, eshift=220)
            r = db.gextract("v", intervals=iv, iterator=170)
            for cid, (c, size) in enumerate(CHROMS):
                m = r["chrom"] == c
                s, e = r["start"][m], r["end"][m]
                if track == "dense_track":
                    bs, bins = read_dense("dense_track", c)
                    exp = port.dense_window_ext(bins, bs, size, s, e, -130, 220, 0)[col]
                else:
                    exp = port.sparse_window_ext(*read_sparse("sparse_track", c), size, s, e, -130, 220)[col]
                if rel:
                    exp = exp - np.maximum(s - 130, 0)
                if func == "lse":
                    np.testing.assert_allclose(r["v"][m], exp, rtol=1e-6, atol=1e-6, equal_nan=True)
                else:
                    assert_same(r["v"][m], exp, f"{track} {func} {c}")
    db.gvtrack_rm("v")
    with pytest.raises(api.MishaError):
        db.gvtrack_create("v", "dense_track", "sample")  # draws from R's RNG: not offered


def test_sparse_track_implicit_iterator_and_nearest(db):
    r = db.gextract("sparse_track", intervals=db.gintervals(1, 0, 10000))  # iterator = the track's own records
    rs, re, rv = read_sparse("sparse_track", "chr1")
    m = rs < 10000
    assert np.array_equal(r["start"], rs[m]) and np.array_equal(r["end"], np.minimum(re[m], 10000))
    assert_same(r["sparse_track"], rv[m].astype(np.float64), "values")
    db.gvtrack_create("near", "sparse_track", "nearest")
    r = db.gextract("near", intervals=db.gintervals(1, 0, 20000), iterator=20)
    exp = port.sparse_window(rs, re, rv, 500000, r["start"], r["end"], funcs=("nearest",))["nearest"]
    assert_same(r["near"], exp, "nearest")
    db.gvtrack_rm("near")


def test_gscreen_and_expressions(db):
    scr = db.gscreen("dense_track > 0.2", intervals=db.gintervals(1), iterator=None)
    bs, bins = read_dense("dense_track", "chr1")
    c, s, e, _ = port.fixedbin_rows(50, [0], [0], [500000])
    v = bins.astype(np.float64); v[np.isinf(v)] = np.nan
    ec, es, ee = port.screen_merge(c, s, e, (v > 0.2).astype(np.int8))
    assert np.array_equal(scr["start"], es) and np.array_equal(scr["end"], ee) and np.all(scr["chrom"] == "chr1")
    # config C4 shape: nearest + coverage predicate over a 20 bp iterator
    rs, re, rv = read_sparse("sparse_track", "chrX")
    db.gvtrack_create("near", "sparse_track", "nearest")
    db.gvtrack_create("cov", api.Intervals(["chrX"] * len(rs), rs, re), "coverage")
    scr = db.gscreen("near > 0.4 & cov > 0.25", intervals=db.gintervals("X"), iterator=20)
    c, s, e, _ = port.fixedbin_rows(20, [2], [0], [200000])
    near = port.sparse_window(rs, re, rv, 200000, s, e, funcs=("nearest",))["nearest"]
    _, us, ue = port.unify(np.zeros(len(rs)), rs, re, True)
    cov = port.coverage(us, ue, 200000, s, e)
    with np.errstate(invalid="ignore"):
        ec, es, ee = port.screen_merge(c, s, e, ((near > 0.4) & (cov > 0.25)).astype(np.int8))
    assert np.array_equal(scr["start"], es) and np.array_equal(scr["end"], ee)
    # arithmetic expression over two variables
    r = db.gextract("dense_track * 2 + near", intervals=db.gintervals("X", 0, 5000), iterator=50, colnames=["x"])
    r2 = db.gextract("dense_track", "near", intervals=db.gintervals("X", 0, 5000), iterator=50)
    assert_same(r["x"], r2["dense_track"] * 2 + r2["near"], "expression")
    with pytest.raises(api.MishaError, match="logical"):
        db.gscreen("dense_track", intervals=db.gintervals(1))
    db.gvtrack_rm("near"); db.gvtrack_rm("cov")


def test_gsummary(db):
    out = db.gsummary("dense_track")
    v = np.concatenate([read_dense("dense_track", c)[1] for c, _ in CHROMS]).astype(np.float64)
    v[np.isinf(v)] = np.nan
    exp = port.summary(v)
    assert list(out) == ["Total intervals", "NaN intervals", "Min", "Max", "Sum", "Mean", "Std dev"]
    got = np.array(list(out.values()))
    assert_same(got[:4], exp[:4], "counts/min/max")
    assert_close(got[4:], exp[4:], 1e-9, "moments")
    db.gvtrack_create("mx", "dense_track", "max")
    db.gvtrack_iterator("mx", -500, 500)
    out = db.gsummary("mx", intervals=db.gintervals([1, 2], [0, 0], [100000, 100000]), iterator=300)
    r = db.gextract("mx", intervals=db.gintervals([1, 2], [0, 0], [100000, 100000]), iterator=300)
    exp = port.summary(r["mx"])
    got = np.array(list(out.values()))
    assert_same(got[:4], exp[:4], "vtrack counts/min/max")
    assert_close(got[4:], exp[4:], 1e-9, "vtrack moments")
    db.gvtrack_rm("mx")


def test_pwm_vtrack(db, golden):
    """tests/testthat/test-pwm.R style: pwm / pwm.max vtracks against the independent oracle, tolerance 1e-6."""
    pssm = golden["pwm_pssm20"]
    iv = db.gintervals([1, 1, 2, "X"], [1000, 200000, 50000, 199500], [1500, 200200, 50020, 200000])
    for func in ("pwm", "pwm.max", "pwm.count"):
        db.gvtrack_create("motif", None, func, dict(pssm=pssm, bidirect=True, prior=0.01, extend=True, **({"score.thresh": -22.0} if func == "pwm.count" else {})))
        r = db.gextract("motif", intervals=iv, iterator=iv)
        logp = port.pssm_logp(pssm, 0.01)
        for cid, (c, size) in enumerate(CHROMS):
            m = r["chrom"] == c
            exp = port.pwm(read_seq(c), logp, r["start"][m], r["end"][m], True, True, func, 1, -22.0)
            if func == "pwm.count":
                assert np.abs(r["motif"][m] - exp).max() <= 1
            else:
                assert_close(r["motif"][m], exp, 1e-6, f"{func} {c}")
    db.gvtrack_rm("motif")
    with pytest.raises(api.MishaError):
        db.gvtrack_create("motif", None, "pwm", dict(pssm=np.ones((5, 3))))


def test_gquantiles(db):
    """gquantiles over a windowed vtrack (the authors' flagship workload, NEWS.md:130) and over a bare track."""
    bs, bins = read_dense("dense_track", "chr1")
    db.gvtrack_create("wmax", "dense_track", "max")
    db.gvtrack_iterator("wmax", sshift=-500, eshift=500)
    pcts = [0.1, 0.5, 0.9, 0.999]
    got = db.gquantiles("wmax", pcts, intervals=db.gintervals(1), iterator=200)
    c, s, e, _ = port.fixedbin_rows(200, [0], [0], [500000])
    col = port.dense_window(bins, bs, 500000, s, e, -500, 500, 0.5, ("max",))["max"]
    exp, _ = port.gquantiles(col, pcts)
    assert [got[p] for p in pcts] == list(exp)
    got = db.gquantiles("dense_track", 0.5, intervals=db.gintervals(1, 0, 100000))
    v = bins[:2000].astype(np.float64)
    v[np.isinf(v)] = np.nan
    assert got[0.5] == port.gquantiles(v, [0.5])[0][0]
    with pytest.raises(api.MishaError, match="not in"):
        db.gquantiles("dense_track", 2)
    db.gvtrack_rm("wmax")


def test_gintervals_quantiles(db):
    """gintervals.quantiles: per interval of the sorted (possibly overlapping) intervals, the quantiles of the expression over the
    iterator rows inside it; NaN where an interval holds no value. Expected from the oracle: rows of the interval, the value
    column, the reference's StreamPercentiler<double> restatement (port.gquantiles)."""
    pcts = [0.0, 0.25, 0.5, 0.9, 1.0]
    iv = db.gintervals([2, 1, 1, "X", 1], [1000, 300000, 10, 150000, 299000], [9000, 330000, 2010, 200000, 305000])  # unsorted, two overlap
    db.gvtrack_create("wavg", "dense_track", "avg")
    db.gvtrack_iterator("wavg", sshift=-120, eshift=80)
    for expr, iterator, sshift, eshift in (("dense_track", None, 0, 0), ("wavg", 170, -120, 80)):
        r = db.gintervals_quantiles(expr, pcts, intervals=iv, iterator=iterator)
        cid = np.array([[c for c, _ in CHROMS].index(x) for x in r["chrom"]])
        assert np.all(np.diff(cid) >= 0) and len(cid) == 5
        for k in range(5):
            c, size = CHROMS[cid[k]]
            bs, bins = read_dense("dense_track", c)
            _, s, e, _ = port.fixedbin_rows(bs if iterator is None else iterator, [0], [r["start"][k]], [r["end"][k]])
            col = port.dense_window(bins, bs, size, s, e, sshift, eshift, 0.5, ("avg",))["avg"]
            exp, nk = port.gquantiles(col, pcts)
            for j, p in enumerate(pcts):
                assert (np.isnan(r[p][k]) and nk == 0) or r[p][k] == exp[j], (expr, k, p)
    db.gvtrack_rm("wavg")
    with pytest.raises(api.MishaError, match="not in"):
        db.gintervals_quantiles("dense_track", [0.5, 1.5], intervals=iv)


def test_gintervals_summary_and_gbins(db):
    """gintervals.summary / gbins.summary / gbins.quantiles of the API mirror against the oracle's IntervalSummary and percentiler
    restatements over the gextract columns of the same rows."""
    iv = db.gintervals([2, 1, 1, "X", 1], [1000, 300000, 10, 150000, 299000], [9000, 330000, 2010, 200000, 305000])
    r = db.gintervals_summary("dense_track", intervals=iv, iterator=150)
    assert len(r["start"]) == 5 and list(r)[3:] == list(gpu.SUMMARY_NAMES)
    for k in range(5):
        one = db.gintervals([r["chrom"][k]], [r["start"][k]], [r["end"][k]])
        col = db.gextract("dense_track", intervals=one, iterator=150)["dense_track"]
        exp = port.summary(col)
        got = np.array([r[nm][k] for nm in gpu.SUMMARY_NAMES])
        assert_same(got[:4], exp[:4], "counts/min/max")
        assert_close(got[4:6], exp[4:6], 1e-11, "sum/mean")
        assert_close(got[6:], exp[6:], 1e-7, "sd")
    scope = db.gintervals([1, 2], [0, 0], [200000, 100000])
    e = db.gextract("dense_track", "sparse_track", intervals=scope, iterator=200)
    dv, sv = e["dense_track"], e["sparse_track"]
    br = [0, 0.1, 0.2, 0.4, 1]
    bidx = port.val2bin(br, sv)
    s = db.gbins_summary("sparse_track", br, expr="dense_track", intervals=scope, iterator=200)
    q = db.gbins_quantiles("sparse_track", br, expr="dense_track", percentiles=[0.25, 0.5, 0.9], intervals=scope, iterator=200)
    assert s.shape == (4, 7) and q.shape == (4, 3)
    for k in range(4):
        exp = port.summary(dv[bidx == k])
        assert_same(s[k, :4], exp[:4], "bin counts/min/max")
        assert_close(s[k, 4:6], exp[4:6], 1e-11, "bin sum/mean")
        assert_close(s[k, 6:], exp[6:], 1e-7, "bin sd")
        eq, _ = port.gquantiles(dv[bidx == k], [0.25, 0.5, 0.9])
        assert np.array_equal(q[k], eq, equal_nan=True)
    s2 = db.gbins_summary("sparse_track", br, "dense_track", [0, 0.05, 0.5], expr="dense_track", intervals=scope, iterator=200)
    assert s2.shape == (4, 2, 7)
    assert s2[..., 0].sum() == np.sum((bidx >= 0) & (port.val2bin([0, 0.05, 0.5], dv) >= 0))
    with pytest.raises(api.MishaError):
        db.gbins_summary("sparse_track", expr="dense_track")


def test_kmer_vtrack(db):
    """gvtrack.create(NULL, "kmer.count" / "kmer.frac", list(kmer=, extend=, strand=)) through gextract, against the KmerCounter restatement;
    parameter checks of .vtrack_params_kmer (R/vtrack.R:292-329)."""
    iv = db.gintervals([1, 1, 2, "X"], [1000, 200000, 50000, 199500], [1500, 200200, 50020, 200000])
    for func, params in (("kmer.count", dict(kmer="CG")), ("kmer.frac", dict(kmer="ttagg", strand=1, extend=False)), ("kmer.count", "GGCC")):
        db.gvtrack_create("km", None, func, params)
        db.gvtrack_iterator("km", sshift=-50, eshift=50)
        r = db.gextract("km", intervals=iv, iterator=iv)
        p = params if isinstance(params, dict) else dict(kmer=params)
        for c, _ in CHROMS:
            m = r["chrom"] == c
            exp = port.kmer(read_seq(c), p["kmer"], r["start"][m], r["end"][m], func, p.get("extend", True), p.get("strand", 0), -50, 50)
            assert_same(r["km"][m], exp, f"{func} {c}")
        db.gvtrack_rm("km")
    with pytest.raises(api.MishaError, match="empty"):
        db.gvtrack_create("km", None, "kmer.count", dict(kmer=""))
    with pytest.raises(api.MishaError, match="unknown"):
        db.gvtrack_create("km", None, "kmer.count", dict(kmer="ACG", bidirect=True))


def test_indexed_database_layout(db, tmp_path):
    """The bundled DB converted to the single-file layouts (track.idx + track.dat per track, seq/genome.idx + genome.seq; written by
    misha_b200/formats.py, which the CPU suite checks against the reference's own readers): every entry point returns what it returns
    on the per-chromosome layout."""
    import os
    import shutil
    from misha_b200 import formats
    root = str(tmp_path / "indexed")
    os.makedirs(root)
    shutil.copy(os.path.join(TESTDB, "chrom_sizes.txt"), root)
    names = [c for c, _ in CHROMS]
    for track, kind in (("dense_track", "dense"), ("sparse_track", "sparse")):
        src = os.path.join(TESTDB, "tracks", track + ".track")
        blobs = {cid: open(os.path.join(src, c), "rb").read() for cid, c in enumerate(names)}
        formats.write_indexed_track(os.path.join(root, "tracks", track + ".track"), kind, blobs)
    formats.write_indexed_genome(os.path.join(root, "seq"), [(cid, c, read_seq(c)) for cid, c in enumerate(names)])
    d2 = api.gdb_init(root)
    try:
        assert d2._track_type("dense_track") == "dense" and d2._track_type("sparse_track") == "sparse"
        iv = db.gintervals_all()
        for d in (db, d2):
            d.gvtrack_create("q", "dense_track", "quantile", 0.5)
            d.gvtrack_iterator("q", sshift=-300, eshift=300)
            d.gvtrack_create("near", "sparse_track", "nearest")
            d.gvtrack_create("k", None, "kmer.count", {"kmer": "ACG", "strand": 0})
        a = db.gextract("dense_track", "q", "near", "k", intervals=iv, iterator=500)
        b = d2.gextract("dense_track", "q", "near", "k", intervals=d2.gintervals_all(), iterator=500)
        for col in ("start", "end", "dense_track", "q", "near", "k", "intervalID"):
            assert_same(np.asarray(b[col], dtype=np.float64), np.asarray(a[col], dtype=np.float64), "indexed layout: " + col)
        sa, sb = db.gsummary("dense_track"), d2.gsummary("dense_track")
        assert list(sa) == list(sb) and np.array_equal(list(sa.values()), list(sb.values()), equal_nan=True)
    finally:
        d2.close()


def test_gtrack_create_from_an_expression(db, tmp_path):
    """gtrack.create (src/GenomeTrackCreate.cpp:35-215): a dense track from an expression over a fixed-bin iterator, a sparse one over
    an intervals iterator, per-chromosome and indexed; what is written reads back as the expression's float values."""
    import os
    import shutil
    root = str(tmp_path / "db")
    shutil.copytree(TESTDB, root)
    d = api.gdb_init(root)
    try:
        d.gvtrack_create("mx", "dense_track", "max")
        d.gvtrack_iterator("mx", sshift=-100, eshift=100)
        src = d.gextract("mx * 2 + 1", intervals=d.gintervals_all(), iterator=100)
        for name, indexed in (("created.pc", False), ("created.ix", True)):
            d.gtrack_create(name, "mx * 2 + 1", iterator=100, indexed=indexed)
            assert d.gtrack_exists(name) and d._track_type(name) == "dense"
            files = sorted(os.listdir(d._track_dir(name)))
            assert files == (["track.dat", "track.idx"] if indexed else ["chr1", "chr2", "chrX"])
            back = d.gextract(name, intervals=d.gintervals_all())  # implicit iterator = the new track's bin size
            assert np.array_equal(back["start"], src["start"])
            assert_same(back[name], src["mx * 2 + 1"].astype(np.float32).astype(np.float64), "dense track created from an expression (%s)" % name)
        with pytest.raises(api.MishaError, match="already exists"):
            d.gtrack_create("created.pc", "dense_track", iterator=100)
        d.gtrack_rm("created.pc")
        assert not d.gtrack_exists("created.pc")
        # sparse: one record per iterator interval
        iv = d.gintervals([1, 1, 2], [100, 5000, 777], [900, 5200, 1500])
        d.gtrack_create("created.sp", "dense_track", iterator=iv)
        assert d._track_type("created.sp") == "sparse"
        want = d.gextract("dense_track", intervals=iv, iterator=iv)
        rs, re_, rv = d._sparse_records("created.sp", 0)
        assert rs.tolist() == [100, 5000] and re_.tolist() == [900, 5200]
        assert_same(rv.astype(np.float64), want["dense_track"][:2].astype(np.float32).astype(np.float64), "sparse track values")
        assert len(d._sparse_records("created.sp", 2)[0]) == 0 and os.path.exists(os.path.join(d._track_dir("created.sp"), "chrX"))
    finally:
        d.close()


def test_value_based_vtracks(db):
    """gvtrack.create(src = data frame with a value column): an in-memory sparse track (GenomeTrackInMemory,
    src/GenomeTrackInMemory.cpp:36-270, "count-based statistics matching GenomeTrackSparse"): the same answers as the sparse reader's
    restatement over those records."""
    rs, re_, rv = read_sparse("sparse_track", "chr1")
    pick = np.arange(0, len(rs), 3)
    src = api.Intervals(["chr1"] * len(pick), rs[pick], re_[pick], score=rv[pick].astype(np.float64))
    iv = db.gintervals([1] * 5, [0, 1000, 40000, 100000, 499000], [600, 9000, 41000, 230000, 500000])
    names = ("avg", "sum", "min", "max", "nearest", "size", "stddev")
    for f in names:
        db.gvtrack_create("vb_" + f, src, f)
    db.gvtrack_create("vb_q", src, "quantile", 0.3)
    db.gvtrack_iterator("vb_max", sshift=-500, eshift=500)
    got = db.gextract(*["vb_" + f for f in names], "vb_q", intervals=iv, iterator=iv)
    exp = port.sparse_window(rs[pick], re_[pick], rv[pick], 500000, iv["start"], iv["end"], 0, 0, 0.3, funcs=names + ("quantile",))
    for f in names:
        if f == "max":
            continue
        assert_same(got["vb_" + f], exp[f], "value-based " + f)
    assert_same(got["vb_q"], exp["quantile"], "value-based quantile")
    exps = port.sparse_window(rs[pick], re_[pick], rv[pick], 500000, iv["start"], iv["end"], -500, 500, 0.5, funcs=("max",))
    assert_same(got["vb_max"], exps["max"], "value-based max under shifts")
    # interval functions over the same data frame keep their meaning; overlapping value-based intervals are refused
    db.gvtrack_create("vb_cov", src, "coverage")
    assert np.all((db.gextract("vb_cov", intervals=iv, iterator=iv)["vb_cov"] >= 0))
    bad = api.Intervals(["chr1", "chr1"], [100, 150], [200, 300], v=[1.0, 2.0])
    with pytest.raises(api.MishaError, match="overlap"):
        db.gvtrack_create("vb_bad", bad, "avg")
    for f in names + ("q", "cov"):
        db.gvtrack_rm("vb_" + f)


def test_array_track_slices(tmp_path):
    """An array track (src/GenomeTrackArrays.cpp) under gvtrack.array.slice: the per-record fold on the host, the window functions on the
    device; expected values from the sparse restatement over the folded records (tests/test_formats.py pins the fold and the file format
    to the reference's own reader / writer)."""
    import os
    import shutil
    from misha_b200 import formats
    root = str(tmp_path / "db")
    shutil.copytree(TESTDB, root)
    rng = np.random.default_rng(31)
    ncol = 8
    recs = {}
    for cid, (c, size) in enumerate(CHROMS):
        n = 300
        st = np.sort(rng.choice(size // 100 - 2, n, replace=False)).astype(np.int64) * 100
        en = st + rng.integers(1, 100, n)
        counts = rng.integers(1, ncol + 1, n)
        rf = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        cols = np.concatenate([np.sort(rng.choice(ncol, k, replace=False)) for k in counts]).astype(np.uint32)
        vals = rng.normal(size=len(cols)).astype(np.float32)
        recs[cid] = (st, en, rf, vals, cols)
    d = os.path.join(root, "tracks", "arr.track")
    formats.write_indexed_track(d, "array", {cid: formats.array_blob(*recs[cid]) for cid in recs})  # the indexed layout of an array track
    db2 = api.gdb_init(root)
    try:
        assert db2._track_type("arr") == "array"
        db2.gvtrack_create("a_all", "arr", "avg")
        db2.gvtrack_create("a_max", "arr", "max")
        db2.gvtrack_array_slice("a_max", [1, 4, 6], "sum")
        db2.gvtrack_create("a_q", "arr", "nearest")
        db2.gvtrack_array_slice("a_q", [0, 2, 3, 7], "quantile", 0.25)
        db2.gvtrack_iterator("a_q", sshift=-2000, eshift=2000)
        iv = db2.gintervals_all()
        got = db2.gextract("a_all", "a_max", "a_q", intervals=iv, iterator=5000)
        off = 0
        for cid, (c, size) in enumerate(CHROMS):
            st, en, rf, vals, cols = recs[cid]
            s = np.arange(0, size, 5000, dtype=np.int64)
            e = np.minimum(s + 5000, size)
            n = len(s)
            for name, sl, sf, sp, f, ss, es in (("a_all", None, "avg", 0.5, "avg", 0, 0), ("a_max", [1, 4, 6], "sum", 0.5, "max", 0, 0),
                                                ("a_q", [0, 2, 3, 7], "quantile", 0.25, "nearest", -2000, 2000)):
                rv = formats.array_slice_values(rf, vals, cols, sl, sf, sp)
                exp = port.sparse_window(st, en, rv, size, s, e, ss, es, 0.5, funcs=(f,))[f]
                assert_same(got[name][off:off + n], exp, "%s %s" % (c, name))
            off += n
        with pytest.raises(api.MishaError, match="array tracks only"):
            db2.gvtrack_create("d", "dense_track", "avg")
            db2.gvtrack_array_slice("d", [0])
    finally:
        db2.close()
