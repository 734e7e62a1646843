"""CPU tests: the oracle (oracle/oracle.c) is pinned against
  (a) tests/golden/ref_vectors.npz -- outputs of the reference's OWN compiled objects (make_fixtures.py),
  (b) the known answers the reference's tests hold for this path,
  (c) the compiled reference live, when oracle/_ref/libmisha_ref.so is present (build container and GPU box)."""
import numpy as np
import pytest

import os

from conftest import CHROMS, GOLDEN, TESTDB, assert_same, read_dense, read_seq, read_sparse
from oracle import port, ref

DENSE_FUNCS = ("avg", "min", "max", "stddev", "sum", "quantile")


def test_dense_matches_reference_golden(golden):
    for c, size in CHROMS:
        bs, bins = read_dense("dense_track", c)
        s, e = golden[f"dense_{c}_start"], golden[f"dense_{c}_end"]
        out = port.dense_window(bins, bs, size, s, e, 0, 0, 0.9, DENSE_FUNCS)
        for k in DENSE_FUNCS:
            assert_same(out[k], golden[f"dense_{c}_{k}"], f"{c} {k}")
        for q in (0.5, 0.0, 1.0):
            out = port.dense_window(bins, bs, size, s, e, 0, 0, q, ("quantile",))
            assert_same(out["quantile"], golden[f"dense_{c}_quantile_{q}"], f"{c} q{q}")
        # the reference's single-function and Kahan sliding fast paths give the same floats on this data
        s, e = golden[f"dense_{c}_slide_start"], golden[f"dense_{c}_slide_end"]
        out = port.dense_window(bins, bs, size, s, e, 0, 0, 0.5, ("avg", "sum"))
        assert_same(out["avg"], golden[f"dense_{c}_slide_avg"], f"{c} slide avg")
        assert_same(out["sum"], golden[f"dense_{c}_slide_sum"], f"{c} slide sum")


def test_dense_survey_known_values():
    """chr1 [0,1000) of the bundled dense_track."""
    bs, bins = read_dense("dense_track", "chr1")
    out = port.dense_window(bins, bs, 500000, [0], [1000], 0, 0, 0.9, DENSE_FUNCS)
    assert np.float32(out["avg"][0]) == np.float32(0.0898888856)
    assert np.float32(out["sum"][0]) == np.float32(1.79777777)
    assert np.float32(out["stddev"][0]) == np.float32(0.075674057)
    assert np.float32(out["quantile"][0]) == np.float32(0.178000003)


def test_sparse_matches_reference_golden(golden):
    funcs = ("avg", "min", "max", "stddev", "sum", "nearest", "quantile")
    for c, size in CHROMS:
        rs, re, rv = read_sparse("sparse_track", c)
        s, e = golden[f"sparse_{c}_start"], golden[f"sparse_{c}_end"]
        out = port.sparse_window(rs, re, rv, size, s, e, 0, 0, 0.9, funcs)
        for k in funcs:
            assert_same(out[k], golden[f"sparse_{c}_{k}"], f"{c} {k}")


def test_sparse_survey_known_values():
    """nearest between records (0,50) and (100,150) ties to the earlier record."""
    rs, re, rv = read_sparse("sparse_track", "chr1")
    out = port.sparse_window(rs, re, rv, 500000, [50, 200, 0], [100, 250, 400], funcs=("avg", "nearest", "max"))
    assert np.isnan(out["avg"][0]) and np.float32(out["nearest"][0]) == np.float32(0.355555564)
    assert np.float32(out["nearest"][1]) == np.float32(0.4)
    assert np.float32(out["avg"][2]) == np.float32(0.371851861) and np.float32(out["max"][2]) == np.float32(0.4)


def test_gdist_known_answers():
    """tests/testthat/test-gdist.R:69-72."""
    _, bins = read_dense("dense_track", "chr1")
    v = bins.astype(np.float64)
    v[np.isinf(v)] = np.nan
    assert int(port.hist(v, [[0, 0.2, 0.5, 1.0]]).sum()) == 6546
    assert int(port.hist(v[:1000], [[0, 0.2, 0.5, 1.0]]).sum()) == 901


def test_val2bin_matches_reference_golden(golden):
    vals = golden["bins_vals"]
    for name in ("uniform", "ragged", "uniform_f32"):
        br = golden[f"bins_{name}_breaks"]
        for il in (0, 1):
            assert np.array_equal(port.val2bin(br, vals, bool(il)), golden[f"bins_{name}_il{il}"]), f"{name} il={il}"


def test_coverage_known_answers():
    """tests/testthat/test-vtrack-coverage.R:5-29,48-69 (source [100,200),[300,400) on chr1)."""
    cs, ce = [100, 300], [200, 400]
    assert_same(port.coverage(cs, ce, 500000, [0, 100], [100, 200]), [0.0, 1.0])
    assert_same(port.coverage(cs, ce, 500000, [0, 150, 350], [50, 250, 450]), [0.0, 0.5, 0.5])
    assert_same(port.coverage(cs, ce, 500000, [0, 100, 200, 300, 400], [100, 200, 300, 400, 500]), [0, 1, 0, 1, 0])
    assert_same(port.coverage(cs, ce, 500000, [0], [500]), [0.4])
    assert np.isnan(port.coverage(cs, ce, 500000, [100], [110], 50, -50)[0])


def test_iterators_match_reference_golden(golden):
    sc, ss, se = golden["it_scope_chrom"], golden["it_scope_start"], golden["it_scope_end"]
    perm = port.sort_perm(sc, ss)
    for b in (500, 50, 333):
        c, s, e, k = port.fixedbin_rows(b, sc[perm], ss[perm], se[perm])
        assert np.array_equal(c, golden[f"it_fixed{b}_chrom"]) and np.array_equal(s, golden[f"it_fixed{b}_start"])
        assert np.array_equal(e, golden[f"it_fixed{b}_end"]) and np.array_equal(perm[k], golden[f"it_fixed{b}_scope"])
    for touch in (0, 1):
        u = port.unify(golden["it_iv_chrom"], golden["it_iv_start"], golden["it_iv_end"], bool(touch))
        for a, key in zip(u, ("chrom", "start", "end")):
            assert np.array_equal(a, golden[f"unify{touch}_{key}"])
    ic, is_, ie = port.unify(golden["it_iv_chrom"], golden["it_iv_start"], golden["it_iv_end"], False)
    sc, ss, se = golden["it_scope2_chrom"], golden["it_scope2_start"], golden["it_scope2_end"]
    perm = port.sort_perm(sc, ss)
    c, s, e, k = port.intervals_rows(ic, is_, ie, sc[perm], ss[perm], se[perm])
    assert np.array_equal(c, golden["it_iv_rows_chrom"]) and np.array_equal(s, golden["it_iv_rows_start"])
    assert np.array_equal(e, golden["it_iv_rows_end"]) and np.array_equal(perm[k], golden["it_iv_rows_scope"])
    # tests/testthat/test-iterator-overlaps.R:444: 2 rows for gintervals(1, 0, 1000) at iterator 500
    assert len(port.fixedbin_rows(500, [0], [0], [1000])[0]) == 2


def test_percentile_matches_reference_golden(golden):
    for n in (1, 2, 3, 10, 511, 512, 525):
        v = golden[f"pct_vals_{n}"]
        got = [port.percentile(v, p) for p in (0.0, 0.1, 0.5, 0.9, 0.999, 1.0)]
        assert_same(np.array(got), golden[f"pct_out_{n}"], f"n={n}")


def test_pwm_matches_reference_golden(golden):
    logp = port.pssm_logp(golden["pwm_pssm20"], 0.01)
    assert np.array_equal(logp, golden["pwm_logp20"])
    for c, size in CHROMS:
        seq = read_seq(c)
        s, e = golden[f"pwm_{c}_start"], golden[f"pwm_{c}_end"]
        for bid in (True, False):
            for ext in (True, False):
                for mode in ("pwm", "pwm.max", "pwm.count"):
                    for strand in (1, -1):
                        if mode == "pwm.max" and strand == -1:
                            continue
                        got = port.pwm(seq, logp, s, e, bid, ext, mode, strand, -25.0)
                        assert_same(got, golden[f"pwm_{c}_{mode}_bid{int(bid)}_ext{int(ext)}_str{strand}"], f"{c} {mode} {bid} {ext} {strand}")


def test_pwm_max_pos_matches_reference_golden(golden):
    """pwm.max.pos of oracle.c against the reference's compiled PWMScorer (ref_vectors_ext.npz); rows clipped below the motif
    length with both strands carry the reference's uninitialised sign and are excluded."""
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "ref_vectors_ext.npz"))
    logp = port.pssm_logp(golden["pwm_pssm20"], 0.01)
    for c, size in CHROMS:
        seq = read_seq(c)
        s, e = golden[f"pwm_{c}_start"], golden[f"pwm_{c}_end"]
        for bid in (True, False):
            for ext in (True, False):
                for strand in (1, -1):
                    got = port.pwm(seq, logp, s, e, bid, ext, "pwm.max.pos", strand)
                    exp = g[f"pwm_{c}_pwm.max.pos_bid{int(bid)}_ext{int(ext)}_str{strand}"]
                    defined = ~(bid & ext & (np.minimum(e + logp.shape[0] - 1, size) - s < logp.shape[0]))
                    assert_same(got[defined], exp[defined], f"{c} pwm.max.pos {bid} {ext} {strand}")


def test_summary_and_screen_merge():
    v = np.array([1.0, np.nan, 3.0, 2.0])
    out = port.summary(v)
    assert out[0] == 4 and out[1] == 1 and out[2] == 1 and out[3] == 3 and out[4] == 6 and out[5] == 2
    assert abs(out[6] - 1.0) < 1e-15
    assert np.all(np.isnan(port.summary([np.nan])[2:]))
    c, s, e = port.screen_merge([0, 0, 0, 1], [0, 10, 30, 30], [10, 20, 40, 40], [1, 1, 1, 1])
    assert c.tolist() == [0, 0, 1] and s.tolist() == [0, 30, 30] and e.tolist() == [20, 40, 40]


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libmisha_ref.so not built")
def test_port_matches_compiled_reference_live():
    """Seeded cross-check against the reference's own objects (sorted rows, as the scanner feeds them)."""
    import os
    rng = np.random.default_rng(123)
    for cid, (c, size) in enumerate(CHROMS):
        s = np.sort(rng.integers(0, size - 1, 3000))
        e = np.minimum(s + rng.choice([1, 20, 50, 777, 10000], 3000), size)
        bs, bins = read_dense("dense_track", c)
        r = ref.dense_eval(os.path.join(TESTDB, "tracks", "dense_track.track", c), cid, s, e, DENSE_FUNCS, 0.37)
        o = port.dense_window(bins, bs, size, s, e, 0, 0, 0.37, DENSE_FUNCS)
        for k in DENSE_FUNCS:
            assert_same(o[k], r[k], f"dense {c} {k}")
        rs, re, rv = read_sparse("sparse_track", c)
        funcs = ("avg", "min", "max", "stddev", "sum", "nearest", "quantile")
        r = ref.sparse_eval(os.path.join(TESTDB, "tracks", "sparse_track.track", c), cid, s, e, funcs, 0.37)
        o = port.sparse_window(rs, re, rv, size, s, e, 0, 0, 0.37, funcs)
        for k in funcs:
            assert_same(o[k], r[k], f"sparse {c} {k}")


def test_gquantiles_oracle_vs_reference_vectors():
    """oracle.c's gquantiles against vectors produced by the reference's own StreamPercentiler<double>
    (tests/golden/make_gquantiles_fixtures.py), bit for bit; and against the compiled reference live when it is present."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "gquantiles_vectors.npz"))
    pcts = g["pcts"]
    for key in g.files:
        if not key.startswith("vals_"):
            continue
        got, n = port.gquantiles(g[key], pcts)
        assert np.array_equal(got, g["q_" + key[5:]], equal_nan=True), key
        assert n == int(np.sum(~np.isnan(g[key].astype(np.float32))))
    try:
        from oracle import ref
        ref.lib()
    except Exception:  # noqa: BLE001 -- no compiled reference on this box
        return
    rng = np.random.default_rng(5)
    v = rng.normal(size=7777)
    v[::13] = np.nan
    assert np.array_equal(port.gquantiles(v, [0.1, 0.5, 0.93])[0], ref.gquantiles(v, [0.1, 0.5, 0.93]))


def test_order_dependent_functions_match_reference_golden():
    """lse, first / last and the four position functions of oracle.c against the reference's own compiled readers
    (tests/golden/ref_vectors_ext.npz, make_fixtures_ext.py): bit for bit, the single-bin assignment quirk included. The
    reference's generic-path lse may reuse its sliding state when two random rows happen to line up: 1e-6 there."""
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "ref_vectors_ext.npz"))
    for c, size in CHROMS:
        bs, bins = read_dense("dense_track", c)
        s, e = g[f"{c}_start"], g[f"{c}_end"]
        out = port.dense_window_ext(bins, bs, size, s, e, 0, 0, 1)
        for k in port.EXT_COLS:
            if k == "lse":
                np.testing.assert_allclose(out[k], g[f"dense_{c}_{k}"], rtol=1e-6, atol=1e-6, equal_nan=True)
                assert np.mean((out[k] == g[f"dense_{c}_{k}"]) | np.isnan(out[k])) > 0.99
            else:
                assert_same(out[k], g[f"dense_{c}_{k}"], f"dense {c} {k}")
        assert_same(port.dense_window_ext(bins, bs, size, s, e, 0, 0, 0)["lse"], g[f"dense_{c}_lse_reducer"], f"dense {c} lse (reducer path)")
        rs, re, rv = read_sparse("sparse_track", c)
        out = port.sparse_window_ext(rs, re, rv, size, s, e)
        for k in port.EXT_COLS:
            assert_same(out[k], g[f"sparse_{c}_{k}"], f"sparse {c} {k}")


@pytest.mark.skipif(not ref.available(), reason="compiled reference not present")
def test_order_dependent_functions_match_reference_live():
    rng = np.random.default_rng(77)
    import os
    from conftest import TESTDB
    for cid, (c, size) in enumerate(CHROMS):
        s = rng.integers(0, size - 1, 1500)
        e = np.minimum(s + rng.choice([1, 30, 50, 77, 450, 2500, 30000], 1500), size)
        bs, bins = read_dense("dense_track", c)
        f = os.path.join(TESTDB, "tracks", "dense_track.track", c)
        got = port.dense_window_ext(bins, bs, size, s, e, 0, 0, 0)
        exp = ref.dense_eval_ext(f, cid, s, e, tuple(k for k in ref.EXT_COLS if k != "lse"), extra_funcs=("max",))
        for k in exp:
            assert_same(got[k], exp[k], f"dense {c} {k}")
        assert_same(got["lse"], ref.dense_eval_ext(f, cid, s, e, ("lse",))["lse"], f"dense {c} lse")
        rs, re, rv = read_sparse("sparse_track", c)
        f = os.path.join(TESTDB, "tracks", "sparse_track.track", c)
        got = port.sparse_window_ext(rs, re, rv, size, s, e)
        exp = ref.sparse_eval_ext(f, cid, s, e)
        for k in exp:
            assert_same(got[k], exp[k], f"sparse {c} {k}")


def test_kmer_matches_reference_golden():
    """kmer.count / kmer.frac of oracle.c against the reference's compiled KmerCounter (tests/golden/kmer_vectors.npz)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "kmer_vectors.npz"))
    for c, _ in CHROMS:
        seq = read_seq(c)
        s, e = g[f"{c}_start"], g[f"{c}_end"]
        for ki, kmer in enumerate(g["kmers"]):
            for mode in ("kmer.count", "kmer.frac"):
                for ext in (True, False):
                    for strand in (0, 1, -1):
                        got = port.kmer(seq, str(kmer), s, e, mode, ext, strand)
                        assert_same(got, g[f"{c}_k{ki}_{mode}_ext{int(ext)}_str{strand}"], f"{c} {kmer} {mode} {ext} {strand}")


def test_masked_matches_reference_golden():
    """masked.count / masked.frac of oracle.c against the reference's compiled MaskedBpCounter (tests/golden/kmer_vectors.npz)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "kmer_vectors.npz"))
    for c, _ in CHROMS:
        seq = read_seq(c)
        s, e = g[f"{c}_start"], g[f"{c}_end"]
        for mode in ("masked.count", "masked.frac"):
            exp = g[f"{c}_{mode}"]
            assert exp.max() > 0
            assert_same(port.masked(seq, s, e, mode), exp, f"{c} {mode}")


def test_path_dependent_counts_and_lse_against_the_compiled_reference():
    """size / exists / lse through the reference's reducer path and through its generic path (tests/golden/path_vectors.npz, made by
    make_path_fixtures.py from the compiled reference): the restatement reproduces both."""
    g = np.load(os.path.join(GOLDEN, "path_vectors.npz"))
    bins, bs, size, s, e = g["bins"], int(g["bin_size"]), int(g["chromsize"]), g["start"], g["end"]
    for name, generic in (("reducer", False), ("generic", True)):
        c = port.dense_counts(bins, bs, size, s, e, generic=generic)
        assert_same(c["size"], g[name + "_size"], name + " size")
        assert_same(c["exists"], g[name + "_exists"], name + " exists")
        lse = port.dense_window_ext(bins, bs, size, s, e, lse_mode=int(generic))["lse"]
        np.testing.assert_allclose(lse, g[name + "_lse"], rtol=1e-6, atol=1e-6, equal_nan=True)
    assert np.sum(g["reducer_size"] != g["generic_size"]) > 100, "the fixture holds windows on which the two paths disagree"


def test_pwm_spatial_weights_against_the_compiled_reference():
    """pwm / pwm.max / pwm.count under spatial weights: the restatement against PWMScorer with spat_factor set (pwm_spat_vectors.npz)."""
    g = np.load(os.path.join(GOLDEN, "pwm_spat_vectors.npz"))
    logp = port.pssm_logp(g["pssm"], float(g["prior"]))
    for c, _ in CHROMS:
        if c + "_start" not in g:
            continue
        seq, s, e = read_seq(c), g[c + "_start"], g[c + "_end"]
        for bid in (1, 0):
            for strand in (1, -1):
                for ext in (1, 0):
                    for mode in ("pwm", "pwm.max", "pwm.count"):
                        exp = g["%s_%s_b%d_s%d_e%d" % (c, mode, bid, strand, ext)]
                        got = port.pwm_spat(seq, logp, s, e, g["spat"], int(g["spat_bin"]), bidirect=bool(bid), extend=bool(ext), mode=mode, strand=strand,
                                            score_thresh=float(g["thresh"]))
                        if mode == "pwm.count":
                            assert_same(got, exp, "%s %s" % (c, mode))
                        else:
                            np.testing.assert_allclose(got, exp, rtol=1e-6, atol=1e-6, equal_nan=True)
