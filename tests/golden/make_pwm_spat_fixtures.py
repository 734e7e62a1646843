"""Golden vectors for PWM vtracks under spatial weights (run in the build container: needs /root/reference compiled into oracle/_ref):

    python tests/golden/make_pwm_spat_fixtures.py   ->  tests/golden/pwm_spat_vectors.npz

PWMScorer::score_interval of the compiled reference with spat_factor / spat_bin set (src/PWMScorer.cpp:112-165, 219-249, 340-362;
DnaPSSM::integrate_energy_logspat / _max_logspat, src/DnaPSSM.cpp:933-1075), a FRESH scorer per interval (the stateless answer: the
reference's spatial sliding window drifts by ~3e-6 and differs on some rows from its own fresh evaluation), over the bundled DB's chr1 /
chrX: pwm, pwm.max, pwm.count x bidirect x strand x extend.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import CHROMS, TESTDB  # noqa: E402
from oracle import ref  # noqa: E402


def main():
    rng = np.random.default_rng(20261021)
    names = [c for c, _ in CHROMS]
    sizes = [s for _, s in CHROMS]
    pssm = rng.dirichlet([0.5] * 4, size=12)
    spat = np.array([0.5, 1.0, 2.0, 4.0, 0.25, 1e-40, 3.0], dtype=np.float32)
    out = {"pssm": pssm, "spat": spat, "spat_bin": np.int64(40), "prior": np.float64(0.01), "thresh": np.float32(-14.0)}
    for cid in (0, 2):
        n = 400
        s = np.sort(rng.integers(0, sizes[cid] - 700, n)).astype(np.int64)
        e = s + rng.integers(3, 600, n)
        s[:3] = [0, sizes[cid] - 30, sizes[cid] - 8]
        e[:3] = [9, sizes[cid], sizes[cid]]
        out["%s_start" % names[cid]], out["%s_end" % names[cid]] = s, e
        for bid in (1, 0):
            for strand in (1, -1):
                for ext in (1, 0):
                    for mode in ("pwm", "pwm.max", "pwm.count"):
                        key = "%s_%s_b%d_s%d_e%d" % (names[cid], mode, bid, strand, ext)
                        out[key] = ref.pwm_eval_spat(TESTDB, names, sizes, pssm, np.full(n, cid, np.int32), s, e, spat, 40, bidirect=bool(bid), prior=0.01,
                                                     extend=bool(ext), mode=mode, strand=strand, score_thresh=-14.0, fresh=True)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "pwm_spat_vectors.npz"), **out)
    print("wrote pwm_spat_vectors.npz with %d arrays" % len(out))


if __name__ == "__main__":
    main()
