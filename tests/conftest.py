import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
TESTDB = os.path.join(GOLDEN, "testdb")
CHROMS = [("chr1", 500000), ("chr2", 300000), ("chrX", 200000)]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def same(a, b):
    """Bit-level equality of two float arrays with NaN == NaN."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def assert_same(a, b, what=""):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    bad = ~((a == b) | (np.isnan(a) & np.isnan(b)))
    assert not bad.any(), f"{what}: {bad.sum()} of {a.size} differ, first at {np.flatnonzero(bad)[:5]}: {a[bad][:5]} vs {b[bad][:5]}"


def assert_close(a, b, rtol, what=""):
    """NaN pattern and infinities must match exactly; finite values within rtol relative."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    na, nb = np.isnan(a), np.isnan(b)
    assert np.array_equal(na, nb), f"{what}: NaN pattern differs at {np.flatnonzero(na != nb)[:5]}"
    fin = np.isfinite(a) & np.isfinite(b)
    assert np.array_equal(a[~fin & ~na], b[~fin & ~na]), f"{what}: infinities differ"
    err = np.abs(a[fin] - b[fin])
    tol = rtol * np.maximum(np.abs(a[fin]), np.abs(b[fin]))
    bad = err > tol
    assert not bad.any(), f"{what}: {bad.sum()} exceed rtol={rtol}; worst rel {np.max(err[bad] / np.maximum(np.abs(a[fin][bad]), 1e-300))}"


def read_dense(track, chrom):
    raw = np.fromfile(os.path.join(TESTDB, "tracks", track + ".track", chrom), dtype=np.uint8)
    bin_size = int(np.frombuffer(raw[:4].tobytes(), dtype=np.uint32)[0])
    return bin_size, np.frombuffer(raw[4:].tobytes(), dtype=np.float32).copy()


def read_sparse(track, chrom):
    raw = np.fromfile(os.path.join(TESTDB, "tracks", track + ".track", chrom), dtype=np.uint8)
    assert np.frombuffer(raw[:4].tobytes(), dtype=np.int32)[0] == -1
    rec = np.frombuffer(raw[4:].tobytes(), dtype=np.dtype([("s", "<i8"), ("e", "<i8"), ("v", "<f4")]))
    return rec["s"].copy(), rec["e"].copy(), rec["v"].copy()


def read_seq(chrom):
    return open(os.path.join(TESTDB, "seq", chrom + ".seq"), "rb").read()


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN, "ref_vectors.npz"))


@pytest.fixture(scope="session")
def gpu_ctx():
    from misha_b200 import gpu
    ctx = gpu.Context([s for _, s in CHROMS])
    yield ctx
    ctx.close()


@pytest.fixture(scope="session")
def dense_testdb(gpu_ctx):
    from misha_b200 import gpu
    bs, _ = read_dense("dense_track", "chr1")
    t = gpu.DenseTrack(gpu_ctx, bs)
    for cid, (c, _) in enumerate(CHROMS):
        t.stage(cid, read_dense("dense_track", c)[1])
    return t
