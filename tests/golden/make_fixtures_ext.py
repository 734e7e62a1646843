#!/usr/bin/env python
"""Regenerates tests/golden/ref_vectors_ext.npz (run in the build container only): outputs of the reference's OWN compiled
readers (oracle/_ref/libmisha_ref.so) for the order-dependent vtrack functions -- lse, first / last, first / last / min / max
positions -- over the bundled example DB (tests/golden/testdb, extracted by make_fixtures.py). Rows are unsorted random
windows, so the reference's sliding-window lse state is never reused (fresh-window semantics)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

CHROMS = [("chr1", 500000), ("chr2", 300000), ("chrX", 200000)]
SEED = 20261020


def main():
    db = os.path.join(HERE, "testdb")
    rng = np.random.default_rng(SEED)
    out = {}
    for cid, (c, size) in enumerate(CHROMS):
        n = 3000
        s = rng.integers(0, size - 1, n)
        ln = rng.choice([1, 7, 49, 50, 51, 100, 999, 1000, 5000, 20000], n)
        e = np.minimum(s + ln, size)
        s = np.concatenate([s, [0, 0, 37, size - 10, size - 51, 12345]]).astype(np.int64)
        e = np.concatenate([e, [1000, 50, 38, size, size, 12346]]).astype(np.int64)
        out[f"{c}_start"], out[f"{c}_end"] = s, e
        f = os.path.join(db, "tracks", "dense_track.track", c)
        r = ref.dense_eval_ext(f, cid, s, e, ref.EXT_COLS, extra_funcs=("avg", "max"))  # generic path: float pairwise lse
        for k in ref.EXT_COLS:
            out[f"dense_{c}_{k}"] = r[k]
        out[f"dense_{c}_lse_reducer"] = ref.dense_eval_ext(f, cid, s, e, ("lse",))["lse"]  # reducer path: double max-trick
        f = os.path.join(db, "tracks", "sparse_track.track", c)
        r = ref.sparse_eval_ext(f, cid, s, e, ref.EXT_COLS, extra_funcs=("avg",))
        for k in ref.EXT_COLS:
            out[f"sparse_{c}_{k}"] = r[k]
    # pwm.max.pos (PWMScorer MAX_LIKELIHOOD_POS -> DnaPSSM::max_like_match): same PSSM and rows as ref_vectors.npz's pwm_* entries
    base = np.load(os.path.join(HERE, "ref_vectors.npz"))
    names = [c for c, _ in CHROMS]
    sizes = [z for _, z in CHROMS]
    for cid, (c, size) in enumerate(CHROMS):
        s, e = base[f"pwm_{c}_start"], base[f"pwm_{c}_end"]
        ch = np.full(len(s), cid, dtype=np.int32)
        for bid in (True, False):
            for ext in (True, False):
                for strand in (1, -1):
                    out[f"pwm_{c}_pwm.max.pos_bid{int(bid)}_ext{int(ext)}_str{strand}"] = ref.pwm_eval(
                        db, names, sizes, base["pwm_pssm20"], ch, s, e, bid, 0.01, ext, "pwm.max.pos", strand, 0.0)
    np.savez_compressed(os.path.join(HERE, "ref_vectors_ext.npz"), **out)
    print("wrote ref_vectors_ext.npz:", len(out), "arrays")


if __name__ == "__main__":
    main()
