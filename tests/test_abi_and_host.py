"""CPU tests of the boundary and the host logic: the C-ABI library loads and exports every declared symbol, the product
never touches the oracle, the synthetic writers speak the reference's on-disk formats, multi-rank reductions work (gloo)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np

from conftest import ROOT


def test_library_exports_every_declared_symbol():
    from misha_b200 import _lib, build
    so = build.build()
    lib = ctypes.CDLL(so)
    declared = _lib.declared_symbols()
    assert len(declared) >= 40
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    lib.mg_abi_version.restype = ctypes.c_int
    assert lib.mg_abi_version() == 1


def test_built_for_sm_100a():
    from misha_b200 import build
    so = build.build()
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", so], capture_output=True, text=True)
    if out.returncode == 0:
        assert "sm_100a" in out.stdout


def test_no_cpu_fallback_init_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        return
    from misha_b200 import gpu
    try:
        gpu.Context([1000])
    except gpu.MishaGpuError as e:
        assert "no CUDA device" in str(e) or "CUDA" in str(e)
    else:
        raise AssertionError("Context() must fail without a GPU")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "misha_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "liboracle" not in txt and "libmisha_ref" not in txt, f


def test_synth_formats_roundtrip(tmp_path):
    from misha_b200 import synth
    from oracle import ref
    chroms = [("chr1", 100000), ("chr2", 60000)]
    synth.write_chrom_sizes(str(tmp_path), chroms)
    assert open(tmp_path / "chrom_sizes.txt").read() == "1\t100000\n2\t60000\n"
    bins = synth.dense_bins(0, 100000, 20)
    assert len(bins) == 5000 and np.isnan(bins).any()
    synth.write_dense_chrom(str(tmp_path), "d", "chr1", bins, 20)
    f = str(tmp_path / "tracks" / "d.track" / "chr1")
    assert os.path.getsize(f) == 4 + 4 * 5000
    s, e, v = synth.sparse_records(0, 100000, 150)
    assert np.all(s[1:] >= e[:-1]) and np.all(s < e) and e[-1] <= 100000
    synth.write_sparse_chrom(str(tmp_path), "s", "chr1", s, e, v)
    assert os.path.getsize(tmp_path / "tracks" / "s.track" / "chr1") == 4 + 20 * 150
    if ref.available():  # the reference's own readers accept the files
        r = ref.dense_eval(f, 0, [0], [100000], ("avg",))
        assert np.isfinite(r["avg"][0])
        r = ref.sparse_eval(str(tmp_path / "tracks" / "s.track" / "chr1"), 0, [0], [100000], ("avg",))
        assert np.isfinite(r["avg"][0])
    a, b = synth.intervals(1, 60000, 100)
    assert np.all(a[1:] > b[:-1]) and b[-1] <= 60000 and np.all((b - a >= 100) & (b - a <= 300))
    assert np.array_equal(synth.dense_bins(0, 100000, 20), bins, equal_nan=True)  # seeded


def test_lpt_sharding_and_counts():
    from misha_b200 import synth
    chroms = synth.genome()
    assert len(chroms) == 24 and chroms[0][0] == "chr1" and chroms[1][0] == "chr10"
    cnt = synth.interval_counts(chroms, 10_000_000)
    assert cnt.sum() == 10_000_000
    for w in (1, 2, 4, 8):
        owner = synth.lpt_shards(chroms, w)
        load = np.zeros(w)
        for i, (_, s) in enumerate(chroms):
            load[owner[i]] += s
        assert load.max() / load.mean() < 1.12


_WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
from misha_b200 import dist as mdist
from oracle import port
rank = int(os.environ["RANK"])
dist.init_process_group("gloo")
rng = np.random.default_rng(5)
v = rng.lognormal(size=20001); v[::13] = np.nan
halves = [v[:9000], v[9000:]]
mine = halves[rank]
fin = mine[~np.isnan(mine)]
part = np.array([len(mine), len(fin), fin.sum(), fin.min(), fin.max(), (fin * fin).sum()])
tot = mdist.reduce_summary(part)
exp = port.summary(v)
assert tot[0] == exp[0] and tot[0] - tot[1] == exp[1] and tot[3] == exp[2] and tot[4] == exp[3]
assert abs(tot[2] - exp[4]) <= 1e-12 * abs(exp[4])
# gbins.summary across ranks: every rank holds the six partials of the same bins
key = rng.random(len(v)); keys = [key[:9000], key[9000:]]
bb = np.array([0, 0.2, 0.5, 0.7, 1.0])
def bin_partials(vals, ks):
    b = port.val2bin(bb, ks)
    out = []
    for j in range(4):
        x = vals[b == j]; f = x[~np.isnan(x)]
        out.append([len(x), len(f), f.sum(), f.min() if len(f) else 1.7976931348623157e308, f.max() if len(f) else -1.7976931348623157e308, (f * f).sum()])
    return np.array(out)
tb = mdist.reduce_summary(bin_partials(mine, keys[rank]))
eb = bin_partials(v, key)
assert tb.shape == (4, 6) and np.array_equal(tb[:, [0, 1, 3, 4]], eb[:, [0, 1, 3, 4]])
assert np.all(np.abs(tb[:, [2, 5]] - eb[:, [2, 5]]) <= 1e-12 * np.abs(eb[:, [2, 5]]))
br = np.linspace(0, 10, 41)
h = mdist.reduce_hist(port.hist(mine, [br]))
assert np.array_equal(h, port.hist(v, [br]))
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_two_rank_reductions_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29731", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"rank {r} ok" in o


def test_r_expression_precedence_rewrite():
    """R: comparison binds tighter than ! & |; the host rewrites before evaluating, and evaluates with R's three-valued logic: a NaN
    operand makes a comparison NA, NA is never TRUE (`!x > c` and `x != c` must not select the row), FALSE & NA is FALSE,
    TRUE | NA is TRUE (misha_b200/api.py)."""
    from misha_b200.api import RLogical, _r_logic_to_python, r_eval
    a, b, c = np.array([0.5, 0.1, 0.9, np.nan]), np.array([0.3, 0.3, 0.1, 0.3]), np.array([1.0, 5.0, 2.0, 3.0])
    env = {"a": a, "b": b, "c": c}

    def true_rows(expr):
        r = r_eval(_r_logic_to_python(expr).replace("is.na(", "is_na("), env)
        assert isinstance(r, RLogical)
        return r.true().tolist()

    assert true_rows("a > 0.4 & b > 0.25") == [True, False, False, False]
    assert true_rows("a > 0.4 | (b < 0.2 & !c >= 3)") == [True, False, True, False]
    assert true_rows("!a > 0.4") == [False, True, False, False]          # !NA is NA
    assert true_rows("a != b && c > 1.5") == [False, True, True, False]  # NA & TRUE is NA
    assert true_rows("a != b | c > 1.5") == [True, True, True, True]     # NA | TRUE is TRUE
    assert true_rows("!(a > 0.4 & c > 100)") == [True, True, True, True]  # NA & FALSE is FALSE, its negation TRUE
    assert true_rows("is.na(a) | a > 0.6") == [False, False, True, True]
    na = r_eval(_r_logic_to_python("a > 0.4"), env)
    assert np.array_equal(na.as_numeric(), [1.0, 0.0, 1.0, np.nan], equal_nan=True)  # what gextract stores for a logical expression
    num = r_eval(_r_logic_to_python("a*2+c").replace("^", "**"), env)
    assert np.allclose(num.v[:3], [2.0, 5.2, 3.8]) and np.isnan(num.v[3])
    assert np.allclose(r_eval("pmin(a, b) + log2(c)", env).v[:3], [0.3, 0.1 + np.log2(5), 0.1 + 1.0])


def test_r_side_glue_compiles_against_the_reference_headers():
    """misha_b200/host/rglue/GpuSession.cpp -- the .Call-shaped entry points a maintainer adds to misha's src/  -- is
    type-checked against the reference's OWN headers and the declaration-only R shim: the R side of the boundary is known to build even
    though no R exists in this image. Needs /root/reference (the build container); skipped on the GPU box."""
    import shutil
    ref_src = "/root/reference/src"
    if not os.path.isdir(ref_src) or shutil.which("g++") is None:
        pytest.skip("reference sources or g++ not present")
    src = os.path.join(ROOT, "misha_b200", "host", "rglue", "GpuSession.cpp")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-w", "-include", "cfloat", "-include", "cstring", "-I", os.path.join(ROOT, "oracle", "shim"),
                        "-I", ref_src, "-I", os.path.join(ROOT, "include"), src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    txt = open(src).read()
    for sym in ("gtracksummary_gpu", "gtrackdist_gpu", "gquantiles_gpu", "gextract_gpu", "mg_group_dense_summary", "mg_group_h_dense_window"):
        assert sym in txt
