#!/usr/bin/env python
"""Golden vectors for gquantiles from the reference's OWN StreamPercentiler<double> (oracle/_ref, compiled in place from
/root/reference/src): python tests/golden/make_gquantiles_fixtures.py  ->  tests/golden/gquantiles_vectors.npz"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import ref  # noqa: E402

rng = np.random.default_rng(20261020)
out = {}
pcts = np.array([0.0, 0.001, 0.05, 0.25, 0.5, 0.75, 0.9, 0.999, 1.0, 1 / 3, 1.5, -0.2])
out["pcts"] = pcts
for name, v in (
    ("normal", rng.normal(size=20001)),
    ("lognormal_nan", np.where(rng.random(30000) < 0.07, np.nan, rng.lognormal(size=30000))),
    ("ties", np.round(rng.normal(size=5000) * 3)),
    ("tiny", np.array([3.5, np.nan, -1.25])),
    ("single", np.array([2.0])),
    ("allnan", np.full(7, np.nan)),
    ("float_narrowing", rng.normal(size=4000) * 1e-3 + 1.0),  # doubles that are not floats: the reference narrows them first
):
    out["vals_" + name] = v.astype(np.float64)
    out["q_" + name] = ref.gquantiles(v, pcts)
np.savez_compressed(os.path.join(HERE, "gquantiles_vectors.npz"), **out)
print("wrote", len(out), "arrays")
