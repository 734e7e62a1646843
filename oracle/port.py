"""ctypes wrapper over oracle.c (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "liboracle.so")
_lib = None

c_i64 = ctypes.c_int64
c_vp = ctypes.c_void_p


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "_ref/liboracle.so"])


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "oracle.c")
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_percentile.restype = ctypes.c_float
        _lib.orc_gquantiles.restype = c_i64
        _lib.orc_unify.restype = c_i64
        _lib.orc_fixedbin_rows.restype = c_i64
        _lib.orc_intervals_rows.restype = c_i64
        _lib.orc_screen_merge.restype = c_i64
        _lib.orc_val2bin.restype = ctypes.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(c_vp)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


DENSE_FUNCS = ("avg", "sum", "min", "max", "stddev", "quantile", "size")


def dense_window(bins, bin_size, chromsize, start, end, sshift=0, eshift=0, percentile=0.5, funcs=DENSE_FUNCS):
    """Dict func -> float64[n] for one chromosome of a dense track (raw bins, may hold +-inf)."""
    bins = np.ascontiguousarray(bins, dtype=np.float32)
    start, end = _i64(start), _i64(end)
    n = len(start)
    outs = {f: (np.empty(n, dtype=np.float64) if f in funcs else None) for f in DENSE_FUNCS}
    lib().orc_dense_window(
        _p(bins), c_i64(len(bins)), ctypes.c_uint32(bin_size), c_i64(chromsize), c_i64(n), _p(start), _p(end),
        c_i64(sshift), c_i64(eshift), ctypes.c_double(percentile),
        _p(outs["avg"]), _p(outs["sum"]), _p(outs["min"]), _p(outs["max"]), _p(outs["stddev"]), _p(outs["quantile"]),
        _p(outs["size"]))
    return {f: outs[f] for f in funcs}


SPARSE_FUNCS = ("avg", "sum", "min", "max", "stddev", "nearest", "quantile", "size")


def sparse_window(rs, re, rv, chromsize, start, end, sshift=0, eshift=0, percentile=0.5, funcs=SPARSE_FUNCS):
    rs, re = _i64(rs), _i64(re)
    rv = np.ascontiguousarray(rv, dtype=np.float32)
    start, end = _i64(start), _i64(end)
    n = len(start)
    outs = {f: (np.empty(n, dtype=np.float64) if f in funcs else None) for f in SPARSE_FUNCS}
    lib().orc_sparse_window(
        _p(rs), _p(re), _p(rv), c_i64(len(rs)), c_i64(chromsize), c_i64(n), _p(start), _p(end), c_i64(sshift), c_i64(eshift),
        ctypes.c_double(percentile), _p(outs["avg"]), _p(outs["sum"]), _p(outs["min"]), _p(outs["max"]), _p(outs["stddev"]),
        _p(outs["nearest"]), _p(outs["quantile"]), _p(outs["size"]))
    return {f: outs[f] for f in funcs}


EXT_COLS = ("lse", "first", "last", "first.pos", "last.pos", "min.pos", "max.pos")


def dense_window_ext(bins, bin_size, chromsize, start, end, sshift=0, eshift=0, lse_mode=0):
    """Dict name -> float64[n] of the order-dependent functions (absolute positions). lse_mode 0: the reducer path's value
    (func = "lse" alone), 1: the generic path's pairwise float accumulation."""
    bins = np.ascontiguousarray(bins, dtype=np.float32)
    start, end = _i64(start), _i64(end)
    out = np.empty((len(start), len(EXT_COLS)), dtype=np.float64)
    lib().orc_dense_window_ext(_p(bins), c_i64(len(bins)), ctypes.c_uint32(bin_size), c_i64(chromsize), c_i64(len(start)), _p(start), _p(end),
                               c_i64(sshift), c_i64(eshift), ctypes.c_int(lse_mode), _p(out))
    return {f: out[:, k].copy() for k, f in enumerate(EXT_COLS)}


def dense_counts(bins, bin_size, chromsize, start, end, sshift=0, eshift=0, generic=False):
    """size / exists of dense windows as the reference's reader reports them: the reducer path counts the non-NaN bins
    (src/GenomeTrackFixedBin.cpp:196-240); the generic path (generic=True) does the same except for a window inside ONE bin that the
    file holds, where assign_single_bin_value reports 1 / 1 whatever the bin's value (:150-180, :664-712)."""
    start, end = _i64(start), _i64(end)
    size = dense_window(bins, bin_size, chromsize, start, end, sshift, eshift, 0.5, ("size",))["size"]
    exists = np.where(np.isnan(size), np.nan, (size > 0).astype(np.float64))
    if generic:
        ws = np.maximum(start + sshift, 0)
        we = np.minimum(end + eshift, chromsize)
        sb = ws // bin_size
        eb = -(-we // bin_size)
        one = (ws < we) & (eb == sb + 1) & (sb < len(bins))
        size = np.where(one, 1.0, size)
        exists = np.where(one, 1.0, exists)
    return {"size": size, "exists": exists}


def sparse_window_ext(rs, re, rv, chromsize, start, end, sshift=0, eshift=0):
    rs, re = _i64(rs), _i64(re)
    rv = np.ascontiguousarray(rv, dtype=np.float32)
    start, end = _i64(start), _i64(end)
    out = np.empty((len(start), len(EXT_COLS)), dtype=np.float64)
    lib().orc_sparse_window_ext(_p(rs), _p(re), _p(rv), c_i64(len(rs)), c_i64(chromsize), c_i64(len(start)), _p(start), _p(end),
                                c_i64(sshift), c_i64(eshift), _p(out))
    return {f: out[:, k].copy() for k, f in enumerate(EXT_COLS)}


def percentile(vals, p):
    v = np.array(vals, dtype=np.float32)
    return float(lib().orc_percentile(_p(v), c_i64(len(v)), ctypes.c_double(p)))


def gquantiles(vals, percentiles):
    """Exact-regime gquantiles of a value column (src/GenomeTrackQuantiles.cpp:167-232). Returns (quantiles, N kept)."""
    v = np.ascontiguousarray(vals, dtype=np.float64)
    p = np.ascontiguousarray(percentiles, dtype=np.float64)
    out = np.empty(len(p), dtype=np.float64)
    n = lib().orc_gquantiles(_p(v), c_i64(len(v)), ctypes.c_int(len(p)), _p(p), _p(out))
    return out, int(n)


def unify(chrom, start, end, unify_touching=True):
    chrom, start, end = _i32(chrom).copy(), _i64(start).copy(), _i64(end).copy()
    m = lib().orc_unify(c_i64(len(chrom)), _p(chrom), _p(start), _p(end), ctypes.c_int(int(unify_touching)))
    return chrom[:m], start[:m], end[:m]


def sort_perm(chrom, start):
    chrom, start = _i32(chrom), _i64(start)
    perm = np.empty(len(chrom), dtype=np.int64)
    lib().orc_sort_perm(c_i64(len(chrom)), _p(chrom), _p(start), _p(perm))
    return perm


def coverage(cs, ce, chromsize, start, end, sshift=0, eshift=0):
    cs, ce, start, end = _i64(cs), _i64(ce), _i64(start), _i64(end)
    out = np.empty(len(start), dtype=np.float64)
    lib().orc_coverage(_p(cs), _p(ce), c_i64(len(cs)), c_i64(chromsize), c_i64(len(start)), _p(start), _p(end), c_i64(sshift),
                       c_i64(eshift), _p(out))
    return out


SUMMARY_NAMES = ("Total intervals", "NaN intervals", "Min", "Max", "Sum", "Mean", "Std dev")


def summary(vals):
    v = np.ascontiguousarray(vals, dtype=np.float64)
    out = np.empty(7, dtype=np.float64)
    lib().orc_summary(_p(v), c_i64(len(v)), _p(out))
    return out


def val2bin(breaks, vals, include_lowest=False):
    b = np.ascontiguousarray(breaks, dtype=np.float64)
    L = lib()
    return np.array([L.orc_val2bin(_p(b), ctypes.c_int(len(b)), ctypes.c_int(int(include_lowest)), ctypes.c_double(float(v)))
                     for v in vals], dtype=np.int32)


def hist(vals, breaks_list, include_lowest=False):
    """vals: n x ndim (or n,) float64; breaks_list: list of break vectors. Returns uint64 array, shape nbins per dim."""
    v = np.ascontiguousarray(vals, dtype=np.float64)
    if v.ndim == 1:
        v = v.reshape(-1, 1)
    ndim = v.shape[1]
    assert ndim == len(breaks_list)
    nb = np.array([len(b) for b in breaks_list], dtype=np.int32)
    allb = np.ascontiguousarray(np.concatenate([np.asarray(b, dtype=np.float64) for b in breaks_list]))
    shape = tuple(int(x) - 1 for x in nb)
    counts = np.zeros(int(np.prod(shape)), dtype=np.uint64)
    lib().orc_hist(_p(v), c_i64(v.shape[0]), ctypes.c_int(ndim), _p(allb), _p(nb), ctypes.c_int(int(include_lowest)), _p(counts))
    return counts.reshape(shape, order="F")


def _rows(fn, args, est):
    cap = max(int(est), 1)
    while True:
        oc = np.empty(cap, dtype=np.int32)
        os_ = np.empty(cap, dtype=np.int64)
        oe = np.empty(cap, dtype=np.int64)
        ok = np.empty(cap, dtype=np.int64)
        r = fn(*args, c_i64(cap), _p(oc), _p(os_), _p(oe), _p(ok))
        if r <= cap:
            return oc[:r], os_[:r], oe[:r], ok[:r]
        cap = r


def fixedbin_rows(binsize, schrom, sstart, send):
    """Scope must be sorted by (chrom,start). Returns (chrom, start, end, scope_idx)."""
    schrom, sstart, send = _i32(schrom), _i64(sstart), _i64(send)
    est = int(np.sum((send - sstart) // binsize + 2)) if len(schrom) else 1
    return _rows(lib().orc_fixedbin_rows, (c_i64(binsize), c_i64(len(schrom)), _p(schrom), _p(sstart), _p(send)), est)


def intervals_rows(ichrom, istart, iend, schrom, sstart, send):
    ichrom, istart, iend = _i32(ichrom), _i64(istart), _i64(iend)
    schrom, sstart, send = _i32(schrom), _i64(sstart), _i64(send)
    return _rows(lib().orc_intervals_rows,
                 (c_i64(len(ichrom)), _p(ichrom), _p(istart), _p(iend), c_i64(len(schrom)), _p(schrom), _p(sstart), _p(send)),
                 len(ichrom) + len(schrom))


def screen_merge(chrom, start, end, flag):
    chrom, start, end = _i32(chrom), _i64(start), _i64(end)
    flag = np.ascontiguousarray(flag, dtype=np.int8)
    n = len(chrom)
    oc, os_, oe = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int64), np.empty(n, dtype=np.int64)
    m = lib().orc_screen_merge(c_i64(n), _p(chrom), _p(start), _p(end), _p(flag), _p(oc), _p(os_), _p(oe))
    return oc[:m], os_[:m], oe[:m]


def pssm_logp(pssm, prior):
    p = np.ascontiguousarray(pssm, dtype=np.float64)
    out = np.empty((p.shape[0], 4), dtype=np.float32)
    lib().orc_pssm_logp(_p(p), ctypes.c_int(p.shape[0]), ctypes.c_double(prior), _p(out))
    return out


PWM_MODES = {"pwm": 0, "pwm.max": 1, "pwm.max.pos": 2, "pwm.count": 3}


def pwm(seq, logp, start, end, bidirect=True, extend=True, mode="pwm", strand=1, score_thresh=0.0, sshift=0, eshift=0):
    """seq: bytes / uint8 array of one chromosome; logp from pssm_logp. Returns float64[n]."""
    s = np.frombuffer(seq, dtype=np.uint8) if isinstance(seq, (bytes, bytearray)) else np.ascontiguousarray(seq, dtype=np.uint8)
    lp = np.ascontiguousarray(logp, dtype=np.float32)
    start, end = _i64(start), _i64(end)
    out = np.empty(len(start), dtype=np.float64)
    lib().orc_pwm(_p(s), c_i64(len(s)), _p(lp), ctypes.c_int(lp.shape[0]), ctypes.c_int(int(bidirect)), ctypes.c_int(int(extend)),
                  ctypes.c_int(PWM_MODES[mode]), ctypes.c_int(strand), ctypes.c_float(score_thresh), c_i64(len(start)), _p(start),
                  _p(end), c_i64(sshift), c_i64(eshift), _p(out))
    return out


def pwm_spat(seq, logp, start, end, spat, spat_bin, bidirect=True, extend=True, mode="pwm", strand=1, score_thresh=0.0, sshift=0, eshift=0):
    """pwm / pwm.max / pwm.count under spatial weights (spat = factors per bin of spat_bin target positions). float64[n]."""
    s = np.frombuffer(seq, dtype=np.uint8) if isinstance(seq, (bytes, bytearray)) else np.ascontiguousarray(seq, dtype=np.uint8)
    lp = np.ascontiguousarray(logp, dtype=np.float32)
    sp = np.ascontiguousarray(spat, dtype=np.float32)
    start, end = _i64(start), _i64(end)
    out = np.empty(len(start), dtype=np.float64)
    lib().orc_pwm_spat(_p(s), c_i64(len(s)), _p(lp), ctypes.c_int(lp.shape[0]), ctypes.c_int(int(bidirect)), ctypes.c_int(int(extend)),
                       ctypes.c_int(PWM_MODES[mode]), ctypes.c_int(strand), ctypes.c_float(score_thresh), _p(sp), ctypes.c_int(len(sp)),
                       ctypes.c_int(int(spat_bin)), c_i64(len(start)), _p(start), _p(end), c_i64(sshift), c_i64(eshift), _p(out))
    return out


def kmer(seq, kmer, start, end, mode="kmer.count", extend=True, strand=0, sshift=0, eshift=0):
    """kmer.count / kmer.frac of one chromosome's bytes per interval (KmerCounter restatement). Returns float64[n]."""
    s = np.frombuffer(seq, dtype=np.uint8) if isinstance(seq, (bytes, bytearray)) else np.ascontiguousarray(seq, dtype=np.uint8)
    start, end = _i64(start), _i64(end)
    out = np.empty(len(start), dtype=np.float64)
    lib().orc_kmer(_p(s), c_i64(len(s)), kmer.encode(), ctypes.c_int({"kmer.count": 0, "kmer.frac": 1}[mode]), ctypes.c_int(int(extend)),
                   ctypes.c_int(strand), c_i64(len(start)), _p(start), _p(end), c_i64(sshift), c_i64(eshift), _p(out))
    return out


def masked(seq, start, end, mode="masked.count", sshift=0, eshift=0):
    """masked.count / masked.frac of one chromosome's bytes per interval (MaskedBpCounter restatement). Returns float64[n]."""
    s = np.frombuffer(seq, dtype=np.uint8) if isinstance(seq, (bytes, bytearray)) else np.ascontiguousarray(seq, dtype=np.uint8)
    start, end = _i64(start), _i64(end)
    out = np.empty(len(start), dtype=np.float64)
    lib().orc_masked(_p(s), c_i64(len(s)), ctypes.c_int({"masked.count": 0, "masked.frac": 1}[mode]), c_i64(len(start)), _p(start), _p(end),
                     c_i64(sshift), c_i64(eshift), _p(out))
    return out
