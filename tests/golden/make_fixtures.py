#!/usr/bin/env python
"""Regenerates tests/golden/ from the reference checkout (run in the build container only).

  1. tests/golden/testdb/   the part of the reference's bundled example DB (inst/testdb.tar.gz) this path reads:
                            chrom_sizes.txt, dense_track, sparse_track, seq/*.seq. Data fixture, not source.
  2. tests/golden/ref_vectors.npz   inputs + outputs of the reference's OWN compiled objects
                            (oracle/_ref/libmisha_ref.so, built by `make -C oracle ref` from the unmodified
                            sources) on seeded inputs over that DB. These pin oracle/oracle.c and, through it
                            and directly, the CUDA path.
"""
import os
import shutil
import sys
import tarfile
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

REF_TESTDB = "/root/reference/inst/testdb.tar.gz"
CHROMS = [("chr1", 500000), ("chr2", 300000), ("chrX", 200000)]  # chromid order (alphabetical, R/db-root.R:157-162)
SEED = 20261019


def extract_testdb(dst):
    with tempfile.TemporaryDirectory() as tmp:
        with tarfile.open(REF_TESTDB) as tf:
            tf.extractall(tmp)
        src = os.path.join(tmp, "trackdb", "test")
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        os.makedirs(os.path.join(dst, "seq"))
        shutil.copy(os.path.join(src, "chrom_sizes.txt"), dst)
        for c, _ in CHROMS:
            shutil.copy(os.path.join(src, "seq", c + ".seq"), os.path.join(dst, "seq"))
        for t in ("dense_track", "sparse_track"):
            os.makedirs(os.path.join(dst, "tracks", t + ".track"))
            for c, _ in CHROMS:
                shutil.copy(os.path.join(src, "tracks", t + ".track", c), os.path.join(dst, "tracks", t + ".track"))


def windows(rng, size, n, sort=True):
    s = rng.integers(0, size - 1, n)
    if sort:
        s = np.sort(s)
    ln = rng.choice([1, 7, 10, 49, 50, 51, 100, 999, 1000, 5000, 20000], n)
    e = np.minimum(s + ln, size)
    return s.astype(np.int64), e.astype(np.int64)


def main():
    db = os.path.join(HERE, "testdb")
    extract_testdb(db)
    rng = np.random.default_rng(SEED)
    out = {}
    dense_funcs = ("avg", "min", "max", "stddev", "sum", "quantile")
    sparse_funcs = ("avg", "min", "max", "stddev", "sum", "nearest", "quantile")
    for cid, (c, size) in enumerate(CHROMS):
        # dense: generic path (all funcs), rows in sorted order plus fixed edge cases
        s, e = windows(rng, size, 4000)
        edge_s = np.array([0, 0, 37, size - 10, size - 50, size - 51, 0, 12345], dtype=np.int64)
        edge_e = np.array([1000, 50, 38, size, size, size, size, 12346], dtype=np.int64)
        s, e = np.concatenate([s, edge_s]), np.concatenate([e, edge_e])
        f = os.path.join(db, "tracks", "dense_track.track", c)
        for q in (0.9, 0.5, 0.0, 1.0):
            r = ref.dense_eval(f, cid, s, e, dense_funcs, q)
            if q == 0.9:
                for k in dense_funcs:
                    out[f"dense_{c}_{k}"] = r[k]
            else:
                out[f"dense_{c}_quantile_{q}"] = r["quantile"]
        out[f"dense_{c}_start"], out[f"dense_{c}_end"] = s, e
        # the reference's single-function / sliding fast paths over a regular sliding scan (avg, sum alone)
        ss = np.arange(0, size - 3000, 50, dtype=np.int64)
        se = ss + 2000
        out[f"dense_{c}_slide_start"], out[f"dense_{c}_slide_end"] = ss, se
        out[f"dense_{c}_slide_avg"] = ref.dense_eval(f, cid, ss, se, ("avg",))["avg"]
        out[f"dense_{c}_slide_sum"] = ref.dense_eval(f, cid, ss, se, ("sum",))["sum"]
        # sparse, sorted rows (the scanner always feeds sorted rows)
        s, e = windows(rng, size, 4000)
        f = os.path.join(db, "tracks", "sparse_track.track", c)
        r = ref.sparse_eval(f, cid, s, e, sparse_funcs, 0.9)
        for k in sparse_funcs:
            out[f"sparse_{c}_{k}"] = r[k]
        out[f"sparse_{c}_start"], out[f"sparse_{c}_end"] = s, e

    # PWM: 20-mer Dirichlet(0.5) PSSM and the 6-row PSSM of the bundled pssms/ (motif 0 core ATTAAT-like)
    names = [c for c, _ in CHROMS]
    sizes = [z for _, z in CHROMS]
    pssm20 = rng.dirichlet([0.5] * 4, size=20)
    out["pwm_pssm20"] = pssm20
    out["pwm_logp20"] = ref.pssm_logp(pssm20, 0.01)
    for cid, (c, size) in enumerate(CHROMS):
        s = np.sort(rng.integers(0, size - 1, 1500)).astype(np.int64)
        ln = rng.choice([1, 5, 19, 20, 21, 100, 500], 1500)
        e = np.minimum(s + ln, size).astype(np.int64)
        s = np.concatenate([s, [size - 10, size - 25, size - 500, 0]]).astype(np.int64)
        e = np.concatenate([e, [size, size, size, 1]]).astype(np.int64)
        out[f"pwm_{c}_start"], out[f"pwm_{c}_end"] = s, e
        ch = np.full(len(s), cid, dtype=np.int32)
        for bid in (True, False):
            for ext in (True, False):
                for mode in ("pwm", "pwm.max", "pwm.count"):
                    for strand in (1, -1):
                        if mode == "pwm.max" and strand == -1:
                            continue  # the reference's sliding path is history-dependent there
                        key = f"pwm_{c}_{mode}_bid{int(bid)}_ext{int(ext)}_str{strand}"
                        out[key] = ref.pwm_eval(db, names, sizes, pssm20, ch, s, e, bid, 0.01, ext, mode, strand, -25.0)

    # gdist bins: uniform and non-uniform breaks, include.lowest on/off
    vals = np.concatenate([rng.uniform(-0.2, 1.2, 5000), [0.0, 0.2, 0.5, 1.0, np.nan, 0.1999999999, 0.2000000001]])
    out["bins_vals"] = vals
    for name, br in (("uniform", np.linspace(0, 1, 101)), ("ragged", np.array([0, 0.2, 0.5, 1.0])),
                     ("uniform_f32", np.arange(11) * 0.1)):
        out[f"bins_{name}_breaks"] = br
        for il in (0, 1):
            out[f"bins_{name}_il{il}"] = ref.val2bin(br, vals, bool(il))

    # iterators
    schrom = np.array([0, 0, 1, 2, 0], dtype=np.int32)
    sstart = np.array([0, 1030, 500, 0, 480000], dtype=np.int64)
    send = np.array([1000, 2777, 2500, 140, 500000], dtype=np.int64)
    out["it_scope_chrom"], out["it_scope_start"], out["it_scope_end"] = schrom, sstart, send
    for b in (500, 50, 333):
        r = ref.fixedbin_rows(b, schrom, sstart, send)
        for k, a in zip(("chrom", "start", "end", "scope"), r):
            out[f"it_fixed{b}_{k}"] = a
    ichrom = rng.integers(0, 3, 400).astype(np.int32)
    istart = rng.integers(0, 190000, 400).astype(np.int64)
    iend = istart + rng.integers(1, 3000, 400)
    out["it_iv_chrom"], out["it_iv_start"], out["it_iv_end"] = ichrom, istart, iend
    sc2 = np.array([0, 0, 1, 2, 2], dtype=np.int32)
    ss2 = np.array([1000, 50000, 0, 100, 150000], dtype=np.int64)
    se2 = np.array([60000, 120000, 300000, 90000, 200000], dtype=np.int64)
    out["it_scope2_chrom"], out["it_scope2_start"], out["it_scope2_end"] = sc2, ss2, se2
    r = ref.intervals_rows(ichrom, istart, iend, sc2, ss2, se2)
    for k, a in zip(("chrom", "start", "end", "scope"), r):
        out[f"it_iv_rows_{k}"] = a
    for touch in (0, 1):
        r = ref.unify(ichrom, istart, iend, bool(touch))
        for k, a in zip(("chrom", "start", "end"), r):
            out[f"unify{touch}_{k}"] = a

    # percentile
    for n in (1, 2, 3, 10, 511, 512, 525):
        v = rng.lognormal(size=n).astype(np.float32)
        out[f"pct_vals_{n}"] = v
        out[f"pct_out_{n}"] = np.array([ref.percentile(v, p) for p in (0.0, 0.1, 0.5, 0.9, 0.999, 1.0)], dtype=np.float64)

    np.savez_compressed(os.path.join(HERE, "ref_vectors.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
