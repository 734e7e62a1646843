#!/usr/bin/env python
"""Regenerates tests/golden/kmer_vectors.npz (run in the build container only): outputs of the reference's OWN compiled KmerCounter
(oracle/_ref/libmisha_ref.so, src/KmerCounter.cpp) for kmer.count / kmer.frac over the bundled example DB's sequence
(tests/golden/testdb/seq): random intervals plus chromosome-edge and shorter-than-kmer ones, every strand / extend combination;
and of its MaskedBpCounter (src/MaskedBpCounter.cpp) for masked.count / masked.frac over the same intervals."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

CHROMS = [("chr1", 500000), ("chr2", 300000), ("chrX", 200000)]
KMERS = ["A", "CG", "TTAGGG", "ACGTACGT", "GGCCAGGCTGGT", "acgtnacg"[:4], "CCCCCCCCCCCCCCCCCCCCCCCCCCCCCCCC"]
SEED = 20261021


def main():
    db = os.path.join(HERE, "testdb")
    rng = np.random.default_rng(SEED)
    names, sizes = [c for c, _ in CHROMS], [z for _, z in CHROMS]
    out = {"kmers": np.array(KMERS)}
    for cid, (c, size) in enumerate(CHROMS):
        n = 400
        s = rng.integers(0, size - 1, n)
        ln = rng.choice([1, 2, 5, 8, 31, 32, 33, 100, 500, 5000], n)
        e = np.minimum(s + ln, size)
        s = np.concatenate([s, [0, 0, size - 3, size - 1, size - 40, size - 8]]).astype(np.int64)
        e = np.concatenate([e, [1, 1000, size, size, size, size - 2]]).astype(np.int64)
        out[f"{c}_start"], out[f"{c}_end"] = s, e
        for mode in ("masked.count", "masked.frac"):
            out[f"{c}_{mode}"] = ref.masked_eval(db, names, sizes, np.full(len(s), cid), s, e, mode)
        for ki, kmer in enumerate(KMERS):
            for mode in ("kmer.count", "kmer.frac"):
                for ext in (True, False):
                    for strand in (0, 1, -1):
                        out[f"{c}_k{ki}_{mode}_ext{int(ext)}_str{strand}"] = ref.kmer_eval(db, names, sizes, kmer, np.full(len(s), cid), s, e, mode, ext, strand)
    np.savez_compressed(os.path.join(HERE, "kmer_vectors.npz"), **out)
    print("wrote kmer_vectors.npz,", len(out), "arrays")


if __name__ == "__main__":
    main()
