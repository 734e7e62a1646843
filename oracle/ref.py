"""ctypes wrapper over oracle/_ref/libmisha_ref.so -- the reference's own compiled hot-path objects.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py). The library is built by `make -C oracle ref` where
/root/reference exists (the build container) and travels prebuilt to the GPU box; `available()` says whether
it can be loaded here.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libmisha_ref.so")
REFERENCE_ROOT = "/root/reference"
_lib = None

c_i64 = ctypes.c_int64
c_vp = ctypes.c_void_p

# Column order of ref_dense_eval / ref_sparse_eval, and GenomeTrack1D::Functions bits (src/GenomeTrack1D.h:46).
COLS = ("avg", "min", "max", "nearest", "stddev", "sum", "quantile")
FUNC_BIT = {"avg": 0, "min": 1, "max": 2, "nearest": 3, "stddev": 4, "sum": 5}


def build():
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "src")):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


def available():
    return os.path.exists(_SO)


def lib():
    global _lib
    if _lib is None:
        if not available():
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.ref_last_error.restype = ctypes.c_char_p
        for name in ("ref_fixedbin_rows", "ref_intervals_rows", "ref_unify"):
            getattr(_lib, name).restype = c_i64
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(c_vp)


def _check(rc):
    if rc < 0:
        raise RuntimeError("reference error: " + lib().ref_last_error().decode())
    return rc


def _mask(funcs):
    m = 0
    for f in funcs:
        if f in FUNC_BIT:
            m |= 1 << FUNC_BIT[f]
    return m


def _eval(fn, chrom_file, chromid, start, end, funcs, percentile):
    start = np.ascontiguousarray(start, dtype=np.int64)
    end = np.ascontiguousarray(end, dtype=np.int64)
    out = np.empty((len(start), len(COLS)), dtype=np.float32)
    _check(fn(chrom_file.encode(), ctypes.c_int(chromid), ctypes.c_uint32(_mask(funcs)), ctypes.c_int(int("quantile" in funcs)),
              ctypes.c_double(percentile), c_i64(len(start)), _p(start), _p(end), _p(out)))
    return {f: out[:, COLS.index(f)].astype(np.float64) for f in funcs}


def dense_eval(chrom_file, chromid, start, end, funcs=("avg",), percentile=0.5):
    """Rows are fed to ONE GenomeTrackFixedBin object in order (history-dependent fast paths included)."""
    return _eval(lib().ref_dense_eval, chrom_file, chromid, start, end, funcs, percentile)


def sparse_eval(chrom_file, chromid, start, end, funcs=("avg",), percentile=0.5):
    return _eval(lib().ref_sparse_eval, chrom_file, chromid, start, end, funcs, percentile)


EXT_COLS = ("lse", "first", "last", "first.pos", "last.pos", "min.pos", "max.pos")
# GenomeTrack1D::Functions (src/GenomeTrack1D.h:47)
EXT_BIT = {"lse": 6, "max.pos": 7, "min.pos": 8, "first": 13, "first.pos": 14, "last": 15, "last.pos": 16}


def _eval_ext(fn, chrom_file, chromid, start, end, funcs, extra_funcs=()):
    start = np.ascontiguousarray(start, dtype=np.int64)
    end = np.ascontiguousarray(end, dtype=np.int64)
    mask = _mask(extra_funcs)
    for f in funcs:
        mask |= 1 << EXT_BIT[f]
    out = np.empty((len(start), len(EXT_COLS)), dtype=np.float64)
    _check(fn(chrom_file.encode(), ctypes.c_int(chromid), ctypes.c_uint32(mask), c_i64(len(start)), _p(start), _p(end), _p(out)))
    return {f: out[:, EXT_COLS.index(f)].copy() for f in funcs}


def dense_eval_ext(chrom_file, chromid, start, end, funcs=EXT_COLS, extra_funcs=()):
    """lse / first / last / positions from ONE GenomeTrackFixedBin fed in row order. funcs = ("lse",) alone runs the reference's
    reducer path, anything else (or extra_funcs such as ("avg", "max")) its generic path."""
    return _eval_ext(lib().ref_dense_eval_ext, chrom_file, chromid, start, end, funcs, extra_funcs)


def dense_eval_counts(chrom_file, chromid, start, end, generic):
    """{size, exists, lse} of a dense track from the reference's reader with only those three registered (its reducer path) or with
    avg and max next to them (generic=True: its generic path)."""
    start = np.ascontiguousarray(start, dtype=np.int64)
    end = np.ascontiguousarray(end, dtype=np.int64)
    mask = (1 << 10) | (1 << 9) | (1 << 6)  # SIZE, EXISTS, LSE (GenomeTrack1D::Functions, src/GenomeTrack1D.h:46)
    if generic:
        mask |= (1 << 0) | (1 << 2)
    out = np.empty((len(start), 3), dtype=np.float64)
    _check(lib().ref_dense_eval_counts(chrom_file.encode(), ctypes.c_int(chromid), ctypes.c_uint32(mask), c_i64(len(start)), _p(start), _p(end), _p(out)))
    return {"size": out[:, 0].copy(), "exists": out[:, 1].copy(), "lse": out[:, 2].copy()}


def sparse_eval_ext(chrom_file, chromid, start, end, funcs=EXT_COLS, extra_funcs=()):
    return _eval_ext(lib().ref_sparse_eval_ext, chrom_file, chromid, start, end, funcs, extra_funcs)


SLICE_FUNCS = {"avg": 0, "min": 1, "max": 2, "stddev": 3, "sum": 4, "quantile": 5}


def array_write(chrom_file, chromid, start, end, rec_first, vals, idx):
    """An array-track chromosome file through the reference's own writer (GenomeTrackArrays::write_next_interval)."""
    start, end = np.ascontiguousarray(start, dtype=np.int64), np.ascontiguousarray(end, dtype=np.int64)
    rf = np.ascontiguousarray(rec_first, dtype=np.int64)
    v, ix = np.ascontiguousarray(vals, dtype=np.float32), np.ascontiguousarray(idx, dtype=np.uint32)
    _check(lib().ref_array_write(chrom_file.encode(), ctypes.c_int(chromid), c_i64(len(start)), _p(start), _p(end), _p(rf), _p(v), _p(ix)))


def array_eval(chrom_file, chromid, start, end, slice_cols, slice_func="avg", slice_percentile=0.5, funcs=("avg",), percentile=0.5):
    """GenomeTrackArrays::read_interval with a slice / slice function set (gvtrack.array.slice)."""
    start, end = np.ascontiguousarray(start, dtype=np.int64), np.ascontiguousarray(end, dtype=np.int64)
    sl = np.ascontiguousarray(slice_cols, dtype=np.uint32)
    out = np.empty((len(start), len(COLS)), dtype=np.float32)
    _check(lib().ref_array_eval(chrom_file.encode(), ctypes.c_int(chromid), ctypes.c_int(len(sl)), _p(sl), ctypes.c_int(SLICE_FUNCS[slice_func]),
                                ctypes.c_double(slice_percentile), ctypes.c_uint32(_mask(funcs)), ctypes.c_int(int("quantile" in funcs)),
                                ctypes.c_double(percentile), c_i64(len(start)), _p(start), _p(end), _p(out)))
    return {f: out[:, COLS.index(f)].astype(np.float64) for f in funcs}


def gquantiles(vals, percentiles):
    v = np.ascontiguousarray(vals, dtype=np.float64)
    p = np.ascontiguousarray(percentiles, dtype=np.float64)
    out = np.empty(len(p), dtype=np.float64)
    _check(lib().ref_gquantiles(_p(v), c_i64(len(v)), ctypes.c_int(len(p)), _p(p), _p(out)))
    return out


def percentile(vals, p):
    v = np.ascontiguousarray(vals, dtype=np.float32)
    out = ctypes.c_float()
    _check(lib().ref_percentile(_p(v), c_i64(len(v)), ctypes.c_double(p), ctypes.byref(out)))
    return float(out.value)


def val2bin(breaks, vals, include_lowest=False, right=True):
    b = np.ascontiguousarray(breaks, dtype=np.float64)
    v = np.ascontiguousarray(vals, dtype=np.float64)
    out = np.empty(len(v), dtype=np.int32)
    _check(lib().ref_val2bin(_p(b), ctypes.c_int(len(b)), ctypes.c_int(int(include_lowest)), ctypes.c_int(int(right)), _p(v),
                             c_i64(len(v)), _p(out)))
    return out


def _names(names):
    arr = (ctypes.c_char_p * len(names))(*[n.encode() for n in names])
    return arr


PWM_MODES = {"pwm": 0, "pwm.max": 1, "pwm.max.pos": 2, "pwm.count": 3}


def pwm_eval(groot, chrom_names, chrom_sizes, pssm, chromid, start, end, bidirect=True, prior=0.01, extend=True, mode="pwm",
             strand=1, score_thresh=0.0):
    p = np.ascontiguousarray(pssm, dtype=np.float64)
    sizes = np.ascontiguousarray(chrom_sizes, dtype=np.int64)
    chromid = np.ascontiguousarray(chromid, dtype=np.int32)
    start = np.ascontiguousarray(start, dtype=np.int64)
    end = np.ascontiguousarray(end, dtype=np.int64)
    out = np.empty(len(start), dtype=np.float32)
    _check(lib().ref_pwm_eval(groot.encode(), ctypes.c_int(len(chrom_names)), _names(chrom_names), _p(sizes), _p(p),
                              ctypes.c_int(p.shape[0]), ctypes.c_int(int(bidirect)), ctypes.c_double(prior), ctypes.c_int(int(extend)),
                              ctypes.c_int(PWM_MODES[mode]), ctypes.c_int(strand), ctypes.c_float(score_thresh), c_i64(len(start)),
                              _p(chromid), _p(start), _p(end), _p(out)))
    return out.astype(np.float64)


def pwm_eval_spat(groot, chrom_names, chrom_sizes, pssm, chromid, start, end, spat, spat_bin, bidirect=True, prior=0.01, extend=True, mode="pwm",
                  strand=1, score_thresh=0.0, fresh=True):
    """PWMScorer::score_interval with spatial weights; fresh: a new scorer per interval (no sliding-window history)."""
    p = np.ascontiguousarray(pssm, dtype=np.float64)
    sizes = np.ascontiguousarray(chrom_sizes, dtype=np.int64)
    chromid = np.ascontiguousarray(chromid, dtype=np.int32)
    start = np.ascontiguousarray(start, dtype=np.int64)
    end = np.ascontiguousarray(end, dtype=np.int64)
    sp = np.ascontiguousarray(spat, dtype=np.float32)
    out = np.empty(len(start), dtype=np.float32)
    _check(lib().ref_pwm_eval_spat(groot.encode(), ctypes.c_int(len(chrom_names)), _names(chrom_names), _p(sizes), _p(p), ctypes.c_int(p.shape[0]),
                                   ctypes.c_int(int(bidirect)), ctypes.c_double(prior), ctypes.c_int(int(extend)), ctypes.c_int(PWM_MODES[mode]),
                                   ctypes.c_int(strand), ctypes.c_float(score_thresh), _p(sp), ctypes.c_int(len(sp)), ctypes.c_int(int(spat_bin)),
                                   ctypes.c_int(int(fresh)), c_i64(len(start)), _p(chromid), _p(start), _p(end), _p(out)))
    return out.astype(np.float64)


def kmer_eval(groot, chrom_names, chrom_sizes, kmer, chromid, start, end, mode="kmer.count", extend=True, strand=0):
    """KmerCounter::score_interval (src/KmerCounter.cpp:38-101) per interval."""
    sizes = np.ascontiguousarray(chrom_sizes, dtype=np.int64)
    chromid = np.ascontiguousarray(chromid, dtype=np.int32)
    start = np.ascontiguousarray(start, dtype=np.int64)
    end = np.ascontiguousarray(end, dtype=np.int64)
    out = np.empty(len(start), dtype=np.float32)
    _check(lib().ref_kmer_eval(groot.encode(), ctypes.c_int(len(chrom_names)), _names(chrom_names), _p(sizes), kmer.encode(),
                               ctypes.c_int({"kmer.count": 0, "kmer.frac": 1}[mode]), ctypes.c_int(int(extend)), ctypes.c_int(strand),
                               c_i64(len(start)), _p(chromid), _p(start), _p(end), _p(out)))
    return out.astype(np.float64)


def masked_eval(groot, chrom_names, chrom_sizes, chromid, start, end, mode="masked.count"):
    """MaskedBpCounter::score_interval (src/MaskedBpCounter.cpp:19-64) per interval."""
    sizes = np.ascontiguousarray(chrom_sizes, dtype=np.int64)
    chromid = np.ascontiguousarray(chromid, dtype=np.int32)
    start = np.ascontiguousarray(start, dtype=np.int64)
    end = np.ascontiguousarray(end, dtype=np.int64)
    out = np.empty(len(start), dtype=np.float32)
    _check(lib().ref_masked_eval(groot.encode(), ctypes.c_int(len(chrom_names)), _names(chrom_names), _p(sizes),
                                 ctypes.c_int({"masked.count": 0, "masked.frac": 1}[mode]), c_i64(len(start)), _p(chromid), _p(start), _p(end),
                                 _p(out)))
    return out.astype(np.float64)


def pssm_logp(pssm, prior):
    p = np.ascontiguousarray(pssm, dtype=np.float64)
    out = np.empty((p.shape[0], 4), dtype=np.float32)
    _check(lib().ref_pssm_logp(_p(p), ctypes.c_int(p.shape[0]), ctypes.c_double(prior), _p(out)))
    return out


def _rows(fn, args, est):
    cap = max(int(est), 1)
    while True:
        oc = np.empty(cap, dtype=np.int32)
        os_ = np.empty(cap, dtype=np.int64)
        oe = np.empty(cap, dtype=np.int64)
        ok = np.empty(cap, dtype=np.int64)
        r = _check(fn(*args, c_i64(cap), _p(oc), _p(os_), _p(oe), _p(ok)))
        if r <= cap:
            return oc[:r], os_[:r], oe[:r], ok[:r]
        cap = r


def fixedbin_rows(binsize, schrom, sstart, send):
    """The scope is sorted inside (as the entry points do); scope_idx refers to the ORIGINAL row (udata)."""
    schrom = np.ascontiguousarray(schrom, dtype=np.int32)
    sstart = np.ascontiguousarray(sstart, dtype=np.int64)
    send = np.ascontiguousarray(send, dtype=np.int64)
    est = int(np.sum((send - sstart) // binsize + 2)) if len(schrom) else 1
    return _rows(lib().ref_fixedbin_rows, (c_i64(binsize), c_i64(len(schrom)), _p(schrom), _p(sstart), _p(send)), est)


def intervals_rows(ichrom, istart, iend, schrom, sstart, send):
    a = [np.ascontiguousarray(x, dtype=t) for x, t in ((ichrom, np.int32), (istart, np.int64), (iend, np.int64),
                                                        (schrom, np.int32), (sstart, np.int64), (send, np.int64))]
    return _rows(lib().ref_intervals_rows,
                 (c_i64(len(a[0])), _p(a[0]), _p(a[1]), _p(a[2]), c_i64(len(a[3])), _p(a[3]), _p(a[4]), _p(a[5])),
                 len(a[0]) + len(a[3]))


def unify(chrom, start, end, unify_touching=True):
    chrom = np.array(chrom, dtype=np.int32)
    start = np.array(start, dtype=np.int64)
    end = np.array(end, dtype=np.int64)
    m = _check(lib().ref_unify(c_i64(len(chrom)), _p(chrom), _p(start), _p(end), ctypes.c_int(int(unify_touching))))
    return chrom[:m], start[:m], end[:m]
