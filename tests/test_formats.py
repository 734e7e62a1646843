"""misha_b200/formats.py against the reference's own readers (compiled into oracle/_ref): an indexed track / genome written here is
read back by GenomeTrackFixedBin / GenomeTrackSparse / GenomeSeqFetch (through TrackIndex / GenomeIndex, checksums included) with the
values that went in, and the per-chromosome and indexed layouts of one track stage the same arrays."""
import os
import struct

import numpy as np
import pytest

from misha_b200 import formats, synth
from oracle import port, ref

SIZES = [50_000, 20_000, 31_000]
NAMES = ["chr1", "chr2", "chrX"]


@pytest.fixture(scope="module")
def db(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("indexed_db"))
    rng = np.random.default_rng(11)
    dense = {i: np.where(rng.random(-(-s // 20)) < 0.1, np.nan, rng.normal(size=-(-s // 20))).astype(np.float32) for i, s in enumerate(SIZES)}
    dense[1] = None  # an empty contig in the indexed layout
    sparse = {i: synth.sparse_records(i, s, 60) for i, s in enumerate(SIZES)}
    seq = {i: synth.sequence(i, s) for i, s in enumerate(SIZES)}
    tdir = os.path.join(root, "tracks")
    formats.write_indexed_track(os.path.join(tdir, "d.track"), "dense", {i: (formats.dense_blob(b, 20) if b is not None else b"") for i, b in dense.items()})
    formats.write_indexed_track(os.path.join(tdir, "s.track"), "sparse", {i: formats.sparse_blob(*r) for i, r in sparse.items()})
    formats.write_indexed_genome(os.path.join(root, "seq"), [(i, NAMES[i], seq[i]) for i in range(3)])
    for i, b in dense.items():  # the same dense track per chromosome
        if b is not None:
            synth.write_dense_chrom(root, "dpc", NAMES[i], b, 20)
    return dict(root=root, tdir=tdir, dense=dense, sparse=sparse, seq=seq)


def test_crc64_known_answer():
    # CRC-64/XZ ("123456789") = 0x995DC9BBDF1939FA: the reflected ECMA polynomial with ~0 init and final complement (src/CRC64.h)
    assert formats.crc64([b"123456789"]) == 0x995DC9BBDF1939FA
    assert formats.crc64([b"1234", b"56789"]) == 0x995DC9BBDF1939FA


def test_indexed_track_round_trip_and_errors(db):
    d = os.path.join(db["tdir"], "d.track")
    idx = formats.read_track_index(d)
    assert idx[1] == "dense" and set(idx[0]) == {0, 1, 2} and idx[0][1][1] == 0
    assert formats.track_type(d, NAMES) == "dense" and formats.track_type(os.path.join(db["tdir"], "s.track"), NAMES) == "sparse"
    for i in (0, 2):
        bs, bins = formats.read_dense_chrom(d, NAMES[i], i, idx)
        assert bs == 20 and np.array_equal(bins, db["dense"][i], equal_nan=True)
        pc = formats.read_dense_chrom(os.path.join(db["tdir"], "dpc.track"), NAMES[i], i, None)
        assert pc[0] == 20 and np.array_equal(pc[1], bins, equal_nan=True)
    assert formats.read_dense_chrom(d, NAMES[1], 1, idx) is None
    sidx = formats.read_track_index(os.path.join(db["tdir"], "s.track"))
    for i in range(3):
        s, e, v = formats.read_sparse_chrom(os.path.join(db["tdir"], "s.track"), NAMES[i], i, sidx)
        assert np.array_equal(s, db["sparse"][i][0]) and np.array_equal(e, db["sparse"][i][1]) and np.array_equal(v, db["sparse"][i][2], equal_nan=True)
    # a flipped byte in the entry table is caught by the checksum, a bad magic by the header check
    raw = bytearray(open(os.path.join(d, "track.idx"), "rb").read())
    raw[36 + 8] ^= 1
    bad = os.path.join(db["root"], "bad.track")
    os.makedirs(bad, exist_ok=True)
    open(os.path.join(bad, "track.idx"), "wb").write(bytes(raw))
    with pytest.raises(formats.FormatError, match="checksum"):
        formats.read_track_index(bad)
    open(os.path.join(bad, "track.idx"), "wb").write(b"NOTMAGIC" + bytes(raw[8:]))
    with pytest.raises(formats.FormatError, match="MISHATDX"):
        formats.read_track_index(bad)


def test_the_reference_reads_what_we_write(db):
    """GenomeTrackFixedBin / GenomeTrackSparse (init_read -> TrackIndex::load -> track.dat) and GenomeSeqFetch (GenomeIndex -> genome.seq)
    of the compiled reference over the files formats.py wrote."""
    if not ref.available():
        pytest.skip("compiled reference not present")
    rng = np.random.default_rng(3)
    for i in (0, 2):
        s = np.sort(rng.integers(0, SIZES[i] - 500, 400)).astype(np.int64)
        e = s + rng.integers(1, 400, 400)
        got = ref.dense_eval(os.path.join(db["tdir"], "d.track", NAMES[i]), i, s, e, ("avg", "max"))
        exp = port.dense_window(db["dense"][i], 20, SIZES[i], s, e, 0, 0, 0.5, ("avg", "max"))
        np.testing.assert_allclose(got["avg"], exp["avg"], rtol=1e-6, equal_nan=True)
        assert np.array_equal(got["max"], exp["max"], equal_nan=True)
        gs = ref.sparse_eval(os.path.join(db["tdir"], "s.track", NAMES[i]), i, s, e, ("avg", "nearest"))
        es = port.sparse_window(*db["sparse"][i], SIZES[i], s, e, 0, 0, 0.5, ("avg", "nearest"))
        assert np.array_equal(gs["avg"], es["avg"], equal_nan=True) and np.array_equal(gs["nearest"], es["nearest"], equal_nan=True)
    # the sequence through the reference's PWM scorer (GenomeSeqFetch in indexed mode) equals the restatement over our bytes
    pssm = synth.pssm(8)
    s = np.sort(rng.integers(0, SIZES[2] - 300, 200)).astype(np.int64)
    e = s + 200
    got = ref.pwm_eval(db["root"], NAMES, SIZES, pssm, np.full(len(s), 2, np.int32), s, e, bidirect=True, prior=0.01, extend=True, mode="pwm")
    exp = port.pwm(db["seq"][2], port.pssm_logp(pssm, 0.01), s, e, bidirect=True, extend=True, mode="pwm")
    np.testing.assert_allclose(got, exp, rtol=1e-6, equal_nan=True)
    gidx = formats.read_genome_index(os.path.join(db["root"], "seq"))
    assert gidx[2][0] == "chrX" and gidx[2][2] == SIZES[2]
    assert np.array_equal(formats.read_seq_chrom(os.path.join(db["root"], "seq"), "chrX", 2, gidx), db["seq"][2])


def test_array_tracks_against_the_reference(tmp_path):
    """Array tracks (src/GenomeTrackArrays.cpp): files written by the reference's own writer are parsed by formats.read_array_chrom, and
    the per-record value under a slice / slice function (array_slice_values) followed by the sparse-track arithmetic reproduces what the
    reference's reader returns for avg / min / max / sum / nearest / stddev of windows; our writer's bytes equal the reference's."""
    if not ref.available():
        pytest.skip("compiled reference not present")
    rng = np.random.default_rng(21)
    size, nrec, ncol = 200_000, 400, 12
    start = np.sort(rng.choice(size // 50 - 2, nrec, replace=False)).astype(np.int64) * 50
    end = start + rng.integers(1, 50, nrec)
    counts = rng.integers(1, ncol + 1, nrec)
    rec_first = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    cols = np.concatenate([np.sort(rng.choice(ncol, k, replace=False)) for k in counts]).astype(np.uint32)
    vals = rng.normal(size=len(cols)).astype(np.float32)
    d = tmp_path / "arr.track"
    os.makedirs(d)
    path = str(d / "chr1")
    ref.array_write(path, 0, start, end, rec_first, vals, cols)
    assert open(path, "rb").read() == formats.array_blob(start, end, rec_first, vals, cols), "our writer == the reference's bytes"
    assert formats.track_type(str(d), ["chr1"]) == "array"
    s2, e2, rf2, v2, c2 = formats.read_array_chrom(str(d), "chr1", 0)
    assert np.array_equal(s2, start) and np.array_equal(e2, end) and np.array_equal(rf2, rec_first) and np.array_equal(v2, vals) and np.array_equal(c2, cols)
    ws = np.sort(rng.integers(0, size - 3000, 500)).astype(np.int64)
    we = ws + rng.integers(1, 3000, 500)
    for slice_cols in ([], [0, 3, 4, 9], [11], [2, 5]):
        for sfunc, sp in (("avg", 0.5), ("min", 0.5), ("max", 0.5), ("sum", 0.5), ("stddev", 0.5), ("quantile", 0.3)):
            rec_val = formats.array_slice_values(rec_first, vals, cols, slice_cols, sfunc, sp)
            names = ("avg", "min", "max", "sum", "nearest", "stddev")
            exp = ref.array_eval(path, 0, ws, we, slice_cols, sfunc, sp, funcs=names)
            got = port.sparse_window(start, end, rec_val, size, ws, we, 0, 0, 0.5, funcs=names)
            for f in names:
                np.testing.assert_allclose(got[f], exp[f], rtol=2e-6, atol=1e-6, equal_nan=True, err_msg="slice %s %s -> %s" % (slice_cols, sfunc, f))
