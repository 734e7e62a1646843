"""Golden vectors for the PATH-DEPENDENT answers of the reference's dense reader (run in the build container, where /root/reference
exists and oracle/_ref is compiled from it):

    python tests/golden/make_path_fixtures.py   ->  tests/golden/path_vectors.npz

size / exists / lse of GenomeTrackFixedBin::read_interval with (a) only those three functions registered -- its reducer path,
classify_fast_path_mode mode 1 -- and (b) avg and max registered next to them -- its generic path (src/GenomeTrackFixedBin.cpp:90-147).
The windows include many that lie inside one bin, NaN or not, aligned to the bin or not, and windows clamped at the chromosome end.
Every window is fed to a FRESH reader (one call per window): the stateless answer, without the sliding-window history.
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402


def main():
    rng = np.random.default_rng(20261020)
    size, bs = 400_010, 20
    nb = -(-size // bs)
    bins = rng.normal(size=nb).astype(np.float32)
    bins[rng.random(nb) < 0.15] = np.nan
    bins[1000:1100] = np.nan
    bins[-1] = np.nan
    d = tempfile.mkdtemp()
    path = os.path.join(d, "chr1")
    with open(path, "wb") as f:
        f.write(np.uint32(bs).tobytes())
        f.write(bins.tobytes())
    s1 = rng.integers(0, size - 1, 1500)            # inside one bin (mostly)
    e1 = np.minimum(s1 + rng.integers(1, bs + 1, 1500), size)
    s2 = rng.integers(0, nb, 500) * bs               # exactly one bin
    e2 = np.minimum(s2 + bs, size)
    s3 = rng.integers(0, size - 1, 1500)            # a few bins
    e3 = np.minimum(s3 + rng.integers(1, 400, 1500), size)
    s4 = np.array([size - 10, size - 30, size - 1, 0, 19, 20000, 20019])
    e4 = np.array([size, size, size, 1, 21, 22000, 20021])
    start = np.concatenate([s1, s2, s3, s4]).astype(np.int64)
    end = np.concatenate([e1, e2, e3, e4]).astype(np.int64)
    out = {"bins": bins, "bin_size": np.int64(bs), "chromsize": np.int64(size), "start": start, "end": end}
    for name, generic in (("reducer", False), ("generic", True)):
        cols = {k: np.empty(len(start)) for k in ("size", "exists", "lse")}
        for i in range(len(start)):  # a fresh reader per window
            r = ref.dense_eval_counts(path, 0, start[i:i + 1], end[i:i + 1], generic)
            for k in cols:
                cols[k][i] = r[k][0]
        for k, v in cols.items():
            out["%s_%s" % (name, k)] = v
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "path_vectors.npz"), **out)
    diff = int(np.sum(out["reducer_size"] != out["generic_size"]))
    print("wrote path_vectors.npz: %d windows, %d where the two paths disagree on size" % (len(start), diff))


if __name__ == "__main__":
    main()
