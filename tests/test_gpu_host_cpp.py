"""The C++ host layer (misha_b200/host: GpuTrackExprScanner, the mirror of the reference's TrackExprScanner for this path)
driven through mg_host_demo, against the oracle: row coordinates / intervalID bit-exact, values at the usual bars. The demo
itself checks that the reference-shaped per-row loop (next() / last_real()) and the whole-column access agree bitwise."""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import assert_close, assert_same
from misha_b200 import build
from oracle import port

RTOL = 1e-6
HERE = os.path.dirname(os.path.abspath(__file__))


def _write_intervals(path, chrom, start, end):
    raw = np.empty(1 + 3 * len(chrom), np.int64)
    raw[0] = len(chrom)
    raw[1::3], raw[2::3], raw[3::3] = chrom, start, end
    raw.tofile(path)


def _read_out(path, nexprs):
    b = open(path, "rb").read()
    n, ne = struct.unpack_from("<qq", b, 0)
    assert ne == nexprs
    off = 16
    chrom = np.frombuffer(b, np.int32, n, off); off += 4 * n
    start = np.frombuffer(b, np.int64, n, off); off += 8 * n
    end = np.frombuffer(b, np.int64, n, off); off += 8 * n
    iid = np.frombuffer(b, np.int32, n, off); off += 4 * n
    cols = []
    for _ in range(ne):
        cols.append(np.frombuffer(b, np.float64, n, off)); off += 8 * n
    return chrom, start, end, iid, cols


def test_host_layer_builds_and_exports():
    """CPU: the host layer compiles against include/mishagpu.h and links libmishagpu.so; the scanner's interface is there."""
    build.build()
    so = build.build_host()
    syms = subprocess.run(["nm", "-DC", so], capture_output=True, text=True).stdout
    for name in ("mishagpu::GpuTrackExprScanner::begin", "mishagpu::GpuTrackExprScanner::next()", "mishagpu::gextract", "mishagpu::gsummary_dense"):
        assert name in syms, name
    assert os.access(build.HOST_DEMO, os.X_OK)


@pytest.mark.gpu
def test_cpp_scanner_gextract_vs_oracle(tmp_path):
    build.build_host()
    rng = np.random.default_rng(77)
    sizes = [400000, 250000]
    bs = 20
    bins = []
    for c, size in enumerate(sizes):
        b = rng.lognormal(size=(size + bs - 1) // bs).astype(np.float32)
        b[rng.random(len(b)) < 0.05] = np.nan
        b[rng.integers(0, len(b), 5)] = np.inf
        b.tofile(tmp_path / f"dense{c}.f32")
        bins.append(b)
    recs = []
    for c, size in enumerate(sizes):
        st = np.sort(rng.choice(size // 50 - 2, 800, replace=False)) * 50
        en = st + rng.integers(1, 50, 800)
        val = rng.random(800).astype(np.float32)
        with open(tmp_path / f"sparse{c}.bin", "wb") as f:
            f.write(struct.pack("<q", 800)); st.astype(np.int64).tofile(f); en.astype(np.int64).tofile(f); val.tofile(f)
        with open(tmp_path / f"iv{c}.bin", "wb") as f:
            f.write(struct.pack("<q", 800)); st.astype(np.int64).tofile(f); en.astype(np.int64).tofile(f)
        recs.append((st, en, val))
    # scope in arbitrary order: intervalID must refer to the caller's order
    n = 3000
    chrom = rng.integers(0, 2, n).astype(np.int64)
    start = (rng.random(n) * (np.array(sizes)[chrom] - 400)).astype(np.int64)
    end = start + rng.integers(1, 400, n)
    _write_intervals(tmp_path / "scope.bin", chrom, start, end)
    spec = tmp_path / "spec.txt"
    spec.write_text("\n".join([
        "chroms 2 %d %d" % tuple(sizes),
        "dense d %d %s %s" % (bs, tmp_path / "dense0.f32", tmp_path / "dense1.f32"),
        "sparse s %s %s" % (tmp_path / "sparse0.bin", tmp_path / "sparse1.bin"),
        "ivset iv %s %s" % (tmp_path / "iv0.bin", tmp_path / "iv1.bin"),
        "vtrack v_avg dense d 0 0 -5000 5000",
        "vtrack v_max dense d 3 0 -5000 5000",
        "vtrack v_q dense d 5 0.9 -5000 5000",
        "vtrack v_sum dense d 1 0 0 0",
        "vtrack v_near sparse s 6 0 0 0",
        "vtrack v_cov coverage iv 0 0 -100 100",
        "scope %s" % (tmp_path / "scope.bin"),
        "iterator intervals %s" % (tmp_path / "scope.bin"),
        "buffer 1024",
        "exprs v_avg v_max v_q v_sum v_near v_cov",
        "out %s" % (tmp_path / "out.bin"),
    ]) + "\n")
    r = subprocess.run([build.HOST_DEMO, str(spec)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    oc, os_, oe, iid, cols = _read_out(tmp_path / "out.bin", 6)
    # the intervals iterator over the scope itself: sorted, overlaps merged (touching kept apart); every row maps to a scope row
    assert np.all(np.diff(oc) >= 0)
    assert np.all((start[iid - 1] < oe) & (end[iid - 1] > os_) & (chrom[iid - 1] == oc)), "intervalID points at an overlapping scope row"
    for c in range(2):
        m = oc == c
        exp = port.dense_window(bins[c], bs, sizes[c], os_[m], oe[m], -5000, 5000, 0.9, ("avg", "max", "quantile"))
        assert_close(cols[0][m], exp["avg"], RTOL, "avg")
        assert_same(cols[1][m], exp["max"], "max")
        assert_same(cols[2][m], exp["quantile"], "quantile")
        exp0 = port.dense_window(bins[c], bs, sizes[c], os_[m], oe[m], 0, 0, 0.5, ("sum",))
        assert_close(cols[3][m], exp0["sum"], RTOL, "sum")
        near = port.sparse_window(*recs[c], sizes[c], os_[m], oe[m], funcs=("nearest",))["nearest"]
        assert_same(cols[4][m], near, "nearest")
        _, us, ue = port.unify(np.zeros(800), recs[c][0], recs[c][1], True)
        assert_same(cols[5][m], port.coverage(us, ue, sizes[c], os_[m], oe[m], -100, 100), "coverage")


@pytest.mark.gpu
def test_cpp_scanner_fixedbin_and_summary(tmp_path):
    build.build_host()
    rng = np.random.default_rng(78)
    sizes = [300000]
    bs = 50
    b = rng.normal(size=sizes[0] // bs).astype(np.float32)
    b[100:140] = np.nan
    b.tofile(tmp_path / "dense0.f32")
    _write_intervals(tmp_path / "scope.bin", [0, 0], [120000, 1010], [200000, 90000])  # unsorted scope
    spec = tmp_path / "spec.txt"
    spec.write_text("\n".join([
        "chroms 1 %d" % sizes[0],
        "dense d %d %s" % (bs, tmp_path / "dense0.f32"),
        "vtrack x dense d 0 0 0 0",
        "scope %s" % (tmp_path / "scope.bin"),
        "iterator 1000",
        "buffer 1024",
        "exprs x",
        "out %s" % (tmp_path / "out.bin"),
        "summary d %d %s" % (bs, tmp_path / "sum.bin"),
    ]) + "\n")
    r = subprocess.run([build.HOST_DEMO, str(spec)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    oc, os_, oe, iid, cols = _read_out(tmp_path / "out.bin", 1)
    ec, es, ee, ek = port.fixedbin_rows(1000, [0, 0], [1010, 120000], [90000, 200000])
    assert np.array_equal(os_, es) and np.array_equal(oe, ee)
    assert np.array_equal(iid, np.where(ek == 0, 2, 1)), "intervalID = row of the caller's (unsorted) scope"
    exp = port.dense_window(b, bs, sizes[0], es, ee, 0, 0, 0.5, ("avg",))
    assert_close(cols[0], exp["avg"], RTOL, "avg @ iterator 1000")
    got = np.fromfile(tmp_path / "sum.bin", np.float64)
    c, s, e, _ = port.fixedbin_rows(bs, [0, 0], [1010, 120000], [90000, 200000])
    vals = port.dense_window(b, bs, sizes[0], s, e, 0, 0, 0.5, ("avg",))["avg"]
    exp_s = port.summary(vals)
    assert np.array_equal(got[:4], exp_s[:4]) and np.allclose(got[4:], exp_s[4:], rtol=1e-9)


@pytest.mark.gpu
def test_cpp_scanner_sequence_and_order_dependent_vtracks(tmp_path):
    """The host scanner with sequence vtracks (vtrack_pwm: pwm, pwm.max, pwm.max.pos through mg_pwm) next to dense vtracks of the
    order-dependent functions (lse, first, max.pos.relative), over a fixed-bin iterator: against the oracle."""
    build.build_host()
    rng = np.random.default_rng(79)
    sizes = [120000, 60000]
    bs = 20
    bins, seqs = [], []
    for c, size in enumerate(sizes):
        b = rng.normal(size=(size + bs - 1) // bs).astype(np.float32)
        b[rng.random(len(b)) < 0.07] = np.nan
        b.tofile(tmp_path / f"dense{c}.f32")
        bins.append(b)
        s = rng.choice(np.frombuffer(b"ACGTacgtN", np.uint8), size, p=[0.2, 0.2, 0.2, 0.2, 0.04, 0.04, 0.04, 0.04, 0.04])
        s.tofile(tmp_path / f"seq{c}.txt")
        seqs.append(s.tobytes())
    pssm = rng.dirichlet([0.5] * 4, size=12)
    pssm.astype(np.float64).tofile(tmp_path / "pssm.f64")
    _write_intervals(tmp_path / "scope.bin", [1, 0, 0], [100, 50000, 37], [59000, 119990, 30011])
    spec = tmp_path / "spec.txt"
    spec.write_text("\n".join([
        "chroms 2 %d %d" % tuple(sizes),
        "dense d %d %s %s" % (bs, tmp_path / "dense0.f32", tmp_path / "dense1.f32"),
        "seq g %s %s" % (tmp_path / "seq0.txt", tmp_path / "seq1.txt"),
        "vtrack v_lse dense d 9 0 -300 300",
        "vtrack v_first dense d 10 0 -300 300",
        "vtrack v_maxpos dense d 15 1 -300 300",
        "pwm v_pwm g 0 %s 0.01 1 1 1 0 0 0" % (tmp_path / "pssm.f64"),
        "pwm v_pmax g 1 %s 0.01 1 1 1 0 -40 40" % (tmp_path / "pssm.f64"),
        "pwm v_ppos g 2 %s 0.01 1 0 -1 0 0 0" % (tmp_path / "pssm.f64"),
        "kmer v_kcount g 0 ac 1 0 -25 25",
        "kmer v_kfrac g 1 TGA 0 -1 0 0",
        "scope %s" % (tmp_path / "scope.bin"),
        "iterator 170",
        "buffer 512",
        "exprs v_lse v_first v_maxpos v_pwm v_pmax v_ppos v_kcount v_kfrac",
        "out %s" % (tmp_path / "out.bin"),
    ]) + "\n")
    r = subprocess.run([build.HOST_DEMO, str(spec)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    oc, os_, oe, iid, cols = _read_out(tmp_path / "out.bin", 8)
    logp = port.pssm_logp(pssm, 0.01)
    assert len(oc) > 900
    for c in range(2):
        m = oc == c
        ext = port.dense_window_ext(bins[c], bs, sizes[c], os_[m], oe[m], -300, 300, 0)
        np.testing.assert_allclose(cols[0][m], ext["lse"], rtol=2e-7, atol=0, equal_nan=True)
        assert_same(cols[1][m], ext["first"], "first")
        assert_same(cols[2][m], ext["max.pos"] - np.maximum(os_[m] - 300, 0), "max.pos.relative")
        assert_close(cols[3][m], port.pwm(seqs[c], logp, os_[m], oe[m], True, True, "pwm", 1), RTOL, "pwm")
        assert_close(cols[4][m], port.pwm(seqs[c], logp, os_[m], oe[m], True, True, "pwm.max", 1, 0.0, -40, 40), RTOL, "pwm.max")
        assert_same(cols[5][m], port.pwm(seqs[c], logp, os_[m], oe[m], True, False, "pwm.max.pos", -1), "pwm.max.pos")
        assert_same(cols[6][m], port.kmer(seqs[c], "ac", os_[m], oe[m], "kmer.count", True, 0, -25, 25), "kmer.count")  # vtrack_kmer -> mg_kmer
        assert_same(cols[7][m], port.kmer(seqs[c], "TGA", os_[m], oe[m], "kmer.frac", False, -1), "kmer.frac")


@pytest.mark.gpu
def test_cpp_group_session(tmp_path):
    """GpuGroupSession (the host layer over mg_group_*): gextract / gsummary / gdist / gquantiles over a group of every device of
    the box (one on a single-GPU box), against the oracle."""
    import torch
    build.build_host()
    ndev = min(torch.cuda.device_count(), 2)
    rng = np.random.default_rng(79)
    sizes = [400000, 150000, 260000]
    bs = 20
    bins, files = [], []
    for c, size in enumerate(sizes):
        b = np.clip(np.exp(rng.normal(size=-(-size // bs))), 0, 100).astype(np.float32)
        b[rng.integers(0, len(b), 50)] = np.nan
        b.tofile(tmp_path / ("dense%d.f32" % c))
        bins.append(b); files.append(str(tmp_path / ("dense%d.f32" % c)))
    n = 3000
    ch = rng.integers(0, 3, n).astype(np.int32)
    st = (rng.random(n) * (np.array(sizes)[ch] - 500)).astype(np.int64)
    en = st + rng.integers(1, 400, n)
    _write_intervals(tmp_path / "scope.bin", ch, st, en)  # unsorted, overlapping
    spec = tmp_path / "spec.txt"
    spec.write_text("\n".join([
        "chroms 3 %d %d %d" % tuple(sizes),
        "dense d %d %s" % (bs, " ".join(files)),
        "scope %s" % (tmp_path / "scope.bin"),
        "group %d d %d %d 0.9 -2000 2000 %s" % (ndev, bs, 5, tmp_path / "grp.bin"),  # quantile(0.9) under +-2 kb
    ]) + "\n")
    r = subprocess.run([build.HOST_DEMO, str(spec)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = np.fromfile(tmp_path / "grp.bin", np.uint8)
    m = int(raw[:8].view(np.int64)[0])
    assert m == n
    col = raw[8:8 + 8 * m].view(np.float64)
    rest = raw[8 + 8 * m:]
    sum7, hist, q = rest[:56].view(np.float64), rest[56:56 + 800].view(np.uint64), rest[856:880].view(np.float64)
    o = np.lexsort((st, ch))  # rows come back in the scope's sorted order (stable)
    for c in range(3):
        mk = ch[o] == c
        exp = port.dense_window(bins[c], bs, sizes[c], st[o][mk], en[o][mk], -2000, 2000, 0.9, ("quantile",))["quantile"]
        assert_same(col[mk], exp, "group gextract quantile chrom %d" % c)
    uc, us, ue = port.unify(ch, st, en)
    vals = []
    for c in range(3):
        mk = uc == c
        _, s, e, _ = port.fixedbin_rows(bs, uc[mk], us[mk], ue[mk])
        vals.append(port.dense_window(bins[c], bs, sizes[c], s, e, 0, 0, 0.5, ("avg",))["avg"])
    vals = np.concatenate(vals)
    exp_s = port.summary(vals)
    assert np.array_equal(sum7[:4], exp_s[:4]) and np.allclose(sum7[4:], exp_s[4:], rtol=1e-9)
    assert np.array_equal(hist, port.hist(vals, [np.arange(101.0)], True))
    assert_same(q, port.gquantiles(vals, [0.1, 0.5, 0.9])[0], "group gquantiles")
