"""Thin object layer over the C ABI (include/mishagpu.h): context, staged tracks, row sources, device columns.

Everything here is plumbing; the arithmetic happens in the CUDA kernels behind libmishagpu.so. numpy is used
for host buffers only.
"""
import ctypes

import numpy as np

from ._lib import MishaGpuError, c_dbl, c_i64, c_int, c_vp, load  # noqa: F401 (MishaGpuError re-exported)

FUNC_IDS = {"avg": 0, "sum": 1, "min": 2, "max": 3, "stddev": 4, "quantile": 5, "nearest": 6, "size": 7, "exists": 8,
            # order-dependent functions; the *.pos ones take param = 1 for positions relative to the window start
            "lse": 9, "first": 10, "last": 11, "first.pos": 12, "last.pos": 13, "min.pos": 14, "max.pos": 15}
KMER_MODES = {"kmer.count": 0, "kmer.frac": 1}
MASKED_MODES = {"masked.count": 0, "masked.frac": 1}
PWM_MODES = {"pwm": 0, "pwm.max": 1, "pwm.max.pos": 2, "pwm.count": 3}


def _p(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(c_vp)
    return c_vp(a)


def _arr(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class DevBuf:
    """A device allocation holding `n` elements of `dtype`."""

    def __init__(self, ctx, n, dtype):
        self.ctx, self.n, self.dtype = ctx, int(n), np.dtype(dtype)
        ptr = c_vp()
        ctx._check(ctx.lib.mg_dev_alloc(ctx.h, c_i64(self.n * self.dtype.itemsize), ctypes.byref(ptr)))
        self.ptr = ptr.value

    def to_host(self, stream=None, count=None):
        """Host copy of the first `count` elements (default: all)."""
        out = np.empty(self.n if count is None else int(count), dtype=self.dtype)
        if out.size:
            self.ctx._check(self.ctx.lib.mg_copy_d2h(self.ctx.h, _p(out), c_vp(self.ptr), c_i64(out.nbytes), c_vp(stream)))
        self.ctx.sync(stream)
        return out

    def from_host(self, a, stream=None):
        a = _arr(a, self.dtype)
        assert a.size == self.n
        if self.n:
            self.ctx._check(self.ctx.lib.mg_copy_h2d(self.ctx.h, c_vp(self.ptr), _p(a), c_i64(a.nbytes), c_vp(stream)))
        self.ctx.sync(stream)
        return self

    def zero(self, stream=None):
        self.ctx._check(self.ctx.lib.mg_dev_memset(self.ctx.h, c_vp(self.ptr), c_int(0), c_i64(self.n * self.dtype.itemsize), c_vp(stream)))
        return self

    def free(self):
        if self.ptr:
            self.ctx.lib.mg_dev_free(self.ctx.h, c_vp(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One per process and GPU (mg_ctx)."""

    def __init__(self, chrom_sizes, device=0):
        self.lib = load()
        sizes = _arr(chrom_sizes, np.int64)
        self.chrom_sizes = sizes
        h = c_vp()
        rc = self.lib.mg_init(c_int(device), c_int(len(sizes)), _p(sizes), ctypes.byref(h))
        if rc:
            raise MishaGpuError(self.lib.mg_last_error(None).decode())
        self.h = h
        self.device = device

    def _check(self, rc):
        if rc:
            raise MishaGpuError(self.lib.mg_last_error(self.h).decode())

    def sync(self, stream=None):
        self._check(self.lib.mg_sync(self.h, c_vp(stream)))

    def set_option(self, name, value):
        self._check(self.lib.mg_set_option(self.h, name.encode(), c_i64(int(value))))

    def launch_count(self):
        return int(self.lib.mg_launch_count(self.h))

    def device_info(self):
        sm, hbm, cc = c_int(), c_i64(), c_int()
        self._check(self.lib.mg_device_info(self.h, ctypes.byref(sm), ctypes.byref(hbm), ctypes.byref(cc)))
        return {"sm_count": sm.value, "hbm_bytes": hbm.value, "cc": cc.value}

    def alloc(self, n, dtype):
        return DevBuf(self, n, dtype)

    def pinned(self, n, dtype):
        """numpy view of pinned host memory (freed with the context)."""
        dtype = np.dtype(dtype)
        ptr = c_vp()
        self._check(self.lib.mg_host_alloc(self.h, c_i64(max(int(n) * dtype.itemsize, 1)), ctypes.byref(ptr)))
        buf = (ctypes.c_char * (int(n) * dtype.itemsize)).from_address(ptr.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(n))
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(ptr.value)
        return arr

    def close(self):
        if getattr(self, "h", None):
            for p in getattr(self, "_pinned", []):
                self.lib.mg_host_free(self.h, c_vp(p))
            self.lib.mg_shutdown(self.h)
            self.h = None

    # ---- rows ----
    def rows_upload(self, chrom, start, end, stream=None):
        chrom, start, end = _arr(chrom, np.int32), _arr(start, np.int64), _arr(end, np.int64)
        h = c_vp()
        self._check(self.lib.mg_rows_upload(self.h, c_i64(len(chrom)), _p(chrom), _p(start), _p(end), c_vp(stream), ctypes.byref(h)))
        self.sync(stream)
        return Rows(self, h)

    def rows_wrap(self, n, d_chrom, d_start, d_end):
        h = c_vp()
        self._check(self.lib.mg_rows_wrap(self.h, c_i64(n), c_vp(d_chrom), c_vp(d_start), c_vp(d_end), ctypes.byref(h)))
        return Rows(self, h)

    def rows_fixedbin(self, binsize, schrom, sstart, send):
        schrom, sstart, send = _arr(schrom, np.int32), _arr(sstart, np.int64), _arr(send, np.int64)
        h = c_vp()
        self._check(self.lib.mg_rows_fixedbin(self.h, c_i64(binsize), c_i64(len(schrom)), _p(schrom), _p(sstart), _p(send), ctypes.byref(h)))
        r = Rows(self, h)
        r.binsize = int(binsize)  # the fixed-bin iterator's bin (None for explicit rows)
        return r

    def rows_intervals(self, ichrom, istart, iend, schrom, sstart, send, stream=None):
        a = [_arr(ichrom, np.int32), _arr(istart, np.int64), _arr(iend, np.int64), _arr(schrom, np.int32), _arr(sstart, np.int64),
             _arr(send, np.int64)]
        h = c_vp()
        self._check(self.lib.mg_rows_intervals(self.h, c_i64(len(a[0])), _p(a[0]), _p(a[1]), _p(a[2]), c_i64(len(a[3])), _p(a[3]),
                                               _p(a[4]), _p(a[5]), c_vp(stream), ctypes.byref(h)))
        return Rows(self, h)


class Rows:
    def __init__(self, ctx, h):
        self.ctx, self.h = ctx, h
        self.n = int(ctx.lib.mg_rows_count(h))
        self.binsize = None

    def coords(self, first=0, count=None, stream=None):
        """(chrom int32, start int64, end int64, scope_idx int64) of rows [first, first+count) on the host."""
        count = self.n - first if count is None else count
        c = self.ctx
        dc, ds, de, dk = c.alloc(count, np.int32), c.alloc(count, np.int64), c.alloc(count, np.int64), c.alloc(count, np.int64)
        c._check(c.lib.mg_rows_coords(c.h, self.h, c_i64(first), c_i64(count), c_vp(dc.ptr), c_vp(ds.ptr), c_vp(de.ptr), c_vp(dk.ptr),
                                      c_vp(stream)))
        return dc.to_host(stream), ds.to_host(stream), de.to_host(stream), dk.to_host(stream)

    def free(self):
        if self.h:
            self.ctx.lib.mg_rows_free(self.ctx.h, self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


FUNC_GENERIC = 0x100  # MG_FUNC_GENERIC (include/mishagpu.h): "size:generic", "exists:generic", "lse:generic"


def _func_args(funcs):
    """funcs: list of names or (name, param). Returns (ids int32[], params float64[]). A ":generic" suffix on size / exists / lse
    asks for the answer of the reference's generic read_interval path (the one it takes when other functions share the track
    object) instead of its reducer path."""
    ids, params = [], []
    for f in funcs:
        name, prm = (f, 0.0) if isinstance(f, str) else f
        flag = 0
        if name.endswith(":generic"):
            name, flag = name[:-len(":generic")], FUNC_GENERIC
        ids.append(FUNC_IDS[name] | flag)
        params.append(float(prm) if prm is not None else 0.0)
    return np.array(ids, dtype=np.int32), np.array(params, dtype=np.float64)


def _out_ptrs(bufs):
    return (c_vp * len(bufs))(*[c_vp(b.ptr) for b in bufs])


class DenseTrack:
    """Dense fixed-bin track staged in HBM (mg_dense)."""

    def __init__(self, ctx, bin_size):
        self.ctx, self.bin_size = ctx, int(bin_size)
        h = c_vp()
        ctx._check(ctx.lib.mg_dense_create(ctx.h, ctypes.c_uint32(self.bin_size), ctypes.byref(h)))
        self.h = h

    def stage(self, chromid, bins):
        bins = _arr(bins, np.float32)
        self.ctx._check(self.ctx.lib.mg_dense_stage(self.ctx.h, self.h, c_int(chromid), _p(bins), c_i64(len(bins))))

    def bytes(self):
        return int(self.ctx.lib.mg_dense_bytes(self.h))

    def window_dev(self, rows, funcs, sshift=0, eshift=0, first=0, count=None, out=None, stream=None):
        """Enqueue; returns device columns (one DevBuf per function)."""
        count = rows.n - first if count is None else count
        ids, params = _func_args(funcs)
        out = out or [self.ctx.alloc(count, np.float64) for _ in ids]
        c = self.ctx
        c._check(c.lib.mg_dense_window(c.h, self.h, rows.h, c_i64(first), c_i64(count), c_i64(sshift), c_i64(eshift), c_int(len(ids)),
                                       _p(ids), _p(params), _out_ptrs(out), c_vp(stream)))
        return out

    def window(self, rows, funcs, sshift=0, eshift=0, first=0, count=None):
        cols = self.window_dev(rows, funcs, sshift, eshift, first, count)
        return [b.to_host() for b in cols]

    def window_host(self, chrom, start, end, funcs, sshift=0, eshift=0, out=None):
        """End-to-end: host arrays in, host columns out (mg_h_dense_window)."""
        chrom, start, end = _arr(chrom, np.int32), _arr(start, np.int64), _arr(end, np.int64)
        ids, params = _func_args(funcs)
        n = len(chrom)
        out = out or [np.empty(n, dtype=np.float64) for _ in ids]
        ptrs = (c_vp * len(out))(*[o.ctypes.data_as(c_vp) for o in out])
        c = self.ctx
        c._check(c.lib.mg_h_dense_window(c.h, self.h, c_i64(n), _p(chrom), _p(start), _p(end), c_i64(sshift), c_i64(eshift),
                                         c_int(len(ids)), _p(ids), _p(params), ptrs))
        return out

    def summary_dev(self, rows, func, partial, sshift=0, eshift=0, first=0, count=None, stream=None):
        count = rows.n - first if count is None else count
        name, prm = (func, 0.0) if isinstance(func, str) else func
        c = self.ctx
        c._check(c.lib.mg_dense_summary(c.h, self.h, rows.h, c_i64(first), c_i64(count), c_i64(sshift), c_i64(eshift),
                                        c_int(FUNC_IDS[name]), c_dbl(prm), c_vp(partial.ptr), c_vp(stream)))

    def hist_dev(self, rows, func, breaks, counts, include_lowest=False, sshift=0, eshift=0, first=0, count=None, stream=None):
        count = rows.n - first if count is None else count
        name, prm = (func, 0.0) if isinstance(func, str) else func
        b = _arr(breaks, np.float64)
        c = self.ctx
        c._check(c.lib.mg_dense_hist(c.h, self.h, rows.h, c_i64(first), c_i64(count), c_i64(sshift), c_i64(eshift), c_int(FUNC_IDS[name]),
                                     c_dbl(prm), _p(b), c_int(len(b)), c_int(int(include_lowest)), c_vp(counts.ptr), c_vp(stream)))

    def free(self):
        if self.h:
            self.ctx.lib.mg_dense_free(self.ctx.h, self.h)
            self.h = None


class SparseTrack:
    def __init__(self, ctx):
        self.ctx = ctx
        h = c_vp()
        ctx._check(ctx.lib.mg_sparse_create(ctx.h, ctypes.byref(h)))
        self.h = h

    def stage(self, chromid, start, end, val):
        start, end, val = _arr(start, np.int64), _arr(end, np.int64), _arr(val, np.float32)
        self.ctx._check(self.ctx.lib.mg_sparse_stage(self.ctx.h, self.h, c_int(chromid), _p(start), _p(end), _p(val), c_i64(len(start))))

    def window_dev(self, rows, funcs, sshift=0, eshift=0, first=0, count=None, out=None, stream=None):
        count = rows.n - first if count is None else count
        ids, params = _func_args(funcs)
        out = out or [self.ctx.alloc(count, np.float64) for _ in ids]
        c = self.ctx
        c._check(c.lib.mg_sparse_window(c.h, self.h, rows.h, c_i64(first), c_i64(count), c_i64(sshift), c_i64(eshift), c_int(len(ids)),
                                        _p(ids), _p(params), _out_ptrs(out), c_vp(stream)))
        return out

    def window(self, rows, funcs, sshift=0, eshift=0, first=0, count=None):
        return [b.to_host() for b in self.window_dev(rows, funcs, sshift, eshift, first, count)]

    def window_host(self, chrom, start, end, funcs, sshift=0, eshift=0, out=None):
        """End-to-end: host arrays in, host columns out (mg_h_sparse_window)."""
        chrom, start, end = _arr(chrom, np.int32), _arr(start, np.int64), _arr(end, np.int64)
        ids, params = _func_args(funcs)
        out = out or [np.empty(len(chrom), dtype=np.float64) for _ in ids]
        ptrs = (c_vp * len(out))(*[o.ctypes.data_as(c_vp) for o in out])
        c = self.ctx
        c._check(c.lib.mg_h_sparse_window(c.h, self.h, c_i64(len(chrom)), _p(chrom), _p(start), _p(end), c_i64(sshift), c_i64(eshift),
                                          c_int(len(ids)), _p(ids), _p(params), ptrs))
        return out

    def free(self):
        if self.h:
            self.ctx.lib.mg_sparse_free(self.ctx.h, self.h)
            self.h = None


class IntervalSet:
    """Sorted, overlap-unified intervals staged for `coverage` (mg_ivset)."""

    def __init__(self, ctx):
        self.ctx = ctx
        h = c_vp()
        ctx._check(ctx.lib.mg_ivset_create(ctx.h, ctypes.byref(h)))
        self.h = h

    def stage(self, chromid, start, end):
        start, end = _arr(start, np.int64), _arr(end, np.int64)
        self.ctx._check(self.ctx.lib.mg_ivset_stage(self.ctx.h, self.h, c_int(chromid), _p(start), _p(end), c_i64(len(start))))

    def coverage_dev(self, rows, sshift=0, eshift=0, first=0, count=None, out=None, stream=None):
        count = rows.n - first if count is None else count
        out = out or self.ctx.alloc(count, np.float64)
        c = self.ctx
        c._check(c.lib.mg_coverage(c.h, self.h, rows.h, c_i64(first), c_i64(count), c_i64(sshift), c_i64(eshift), c_vp(out.ptr), c_vp(stream)))
        return out

    def coverage(self, rows, sshift=0, eshift=0, first=0, count=None):
        return self.coverage_dev(rows, sshift, eshift, first, count).to_host()

    def coverage_host(self, chrom, start, end, sshift=0, eshift=0, out=None):
        """End-to-end: host arrays in, host column out (mg_h_coverage)."""
        chrom, start, end = _arr(chrom, np.int32), _arr(start, np.int64), _arr(end, np.int64)
        out = np.empty(len(chrom), dtype=np.float64) if out is None else out
        c = self.ctx
        c._check(c.lib.mg_h_coverage(c.h, self.h, c_i64(len(chrom)), _p(chrom), _p(start), _p(end), c_i64(sshift), c_i64(eshift), _p(out)))
        return out

    def free(self):
        if self.h:
            self.ctx.lib.mg_ivset_free(self.ctx.h, self.h)
            self.h = None


class Sequence:
    """Genome sequence staged 2-bit packed (mg_seq)."""

    def __init__(self, ctx):
        self.ctx = ctx
        h = c_vp()
        ctx._check(ctx.lib.mg_seq_create(ctx.h, ctypes.byref(h)))
        self.h = h

    def stage(self, chromid, ascii_bytes):
        a = np.frombuffer(ascii_bytes, dtype=np.uint8) if isinstance(ascii_bytes, (bytes, bytearray)) else _arr(ascii_bytes, np.uint8)
        self.ctx._check(self.ctx.lib.mg_seq_stage(self.ctx.h, self.h, c_int(chromid), _p(a), c_i64(len(a))))

    def pwm_dev(self, rows, logp, bidirect=True, extend=True, mode="pwm", strand=1, score_thresh=0.0, sshift=0, eshift=0, first=0,
                count=None, out=None, stream=None, spat_factor=None, spat_bin=1):
        """spat_factor / spat_bin: spatial weights (mg_pwm_spatial); None: mg_pwm."""
        count = rows.n - first if count is None else count
        lp = _arr(logp, np.float32)
        out = out or self.ctx.alloc(count, np.float64)
        c = self.ctx
        sp = _arr(spat_factor, np.float32) if spat_factor is not None else np.zeros(0, np.float32)
        c._check(c.lib.mg_pwm_spatial(c.h, self.h, _p(lp), c_int(lp.shape[0]), c_int(int(bidirect)), c_int(int(extend)), c_int(PWM_MODES[mode]),
                                      c_int(strand), ctypes.c_float(score_thresh), _p(sp) if len(sp) else None, c_int(len(sp)), c_int(int(spat_bin)),
                                      rows.h, c_i64(first), c_i64(count), c_i64(sshift), c_i64(eshift), c_vp(out.ptr), c_vp(stream)))
        return out

    def pwm(self, rows, logp, **kw):
        return self.pwm_dev(rows, logp, **kw).to_host()

    def masked_dev(self, rows, mode="masked.count", sshift=0, eshift=0, first=0, count=None, out=None, stream=None):
        """mg_masked: masked.count / masked.frac of the rows (device column)."""
        count = rows.n - first if count is None else count
        out = out or self.ctx.alloc(count, np.float64)
        c = self.ctx
        c._check(c.lib.mg_masked(c.h, self.h, c_int(MASKED_MODES[mode]), rows.h, c_i64(first), c_i64(count), c_i64(sshift), c_i64(eshift),
                                 c_vp(out.ptr), c_vp(stream)))
        return out

    def kmer_dev(self, rows, kmer, mode="kmer.count", extend=True, strand=0, sshift=0, eshift=0, first=0, count=None, out=None, stream=None):
        """mg_kmer: kmer.count / kmer.frac of the rows (device column)."""
        count = rows.n - first if count is None else count
        out = out or self.ctx.alloc(count, np.float64)
        c = self.ctx
        c._check(c.lib.mg_kmer(c.h, self.h, kmer.encode(), c_int(KMER_MODES[mode]), c_int(int(extend)), c_int(strand), rows.h, c_i64(first),
                               c_i64(count), c_i64(sshift), c_i64(eshift), c_vp(out.ptr), c_vp(stream)))
        return out

    def pwm_host(self, chrom, start, end, logp, bidirect=True, extend=True, mode="pwm", strand=1, score_thresh=0.0, sshift=0, eshift=0, out=None):
        """End-to-end: host arrays in, host column out (mg_h_pwm)."""
        chrom, start, end = _arr(chrom, np.int32), _arr(start, np.int64), _arr(end, np.int64)
        lp = _arr(logp, np.float32)
        out = np.empty(len(chrom), dtype=np.float64) if out is None else out
        c = self.ctx
        c._check(c.lib.mg_h_pwm(c.h, self.h, _p(lp), c_int(lp.shape[0]), c_int(int(bidirect)), c_int(int(extend)), c_int(PWM_MODES[mode]),
                                c_int(strand), ctypes.c_float(score_thresh), c_i64(len(chrom)), _p(chrom), _p(start), _p(end), c_i64(sshift),
                                c_i64(eshift), _p(out)))
        return out

    def free(self):
        if self.h:
            self.ctx.lib.mg_seq_free(self.ctx.h, self.h)
            self.h = None


def pssm_logp(pssm, prior):
    p = _arr(pssm, np.float64)
    out = np.empty((p.shape[0], 4), dtype=np.float32)
    rc = load().mg_pssm_logp(_p(p), c_int(p.shape[0]), c_dbl(prior), _p(out))
    if rc:
        raise MishaGpuError("mg_pssm_logp failed")
    return out


SUMMARY_NAMES = ("Total intervals", "NaN intervals", "Min", "Max", "Sum", "Mean", "Std dev")


def summary_init(ctx, stream=None):
    part = ctx.alloc(6, np.float64)
    ctx._check(ctx.lib.mg_summary_init(ctx.h, c_vp(part.ptr), c_vp(stream)))
    return part


def summary_add(ctx, dcol, n, partial, stream=None):
    ctx._check(ctx.lib.mg_summary(ctx.h, c_vp(dcol.ptr if isinstance(dcol, DevBuf) else dcol), c_i64(n), c_vp(partial.ptr), c_vp(stream)))


def summary_finalize(partial6):
    p = _arr(partial6, np.float64)
    out = np.empty(7, dtype=np.float64)
    load().mg_summary_finalize(_p(p), _p(out))
    return out


def summary_finalize_rows(partials):
    p = _arr(partials, np.float64).reshape(-1, 6)
    out = np.empty((len(p), 7), dtype=np.float64)
    load().mg_summary_finalize_rows(_p(p), c_i64(len(p)), _p(out))
    return out


def hist_dev(ctx, dcols, n, breaks_list, counts, include_lowest=False, stream=None):
    nb = np.array([len(b) for b in breaks_list], dtype=np.int32)
    allb = _arr(np.concatenate([np.asarray(b, dtype=np.float64) for b in breaks_list]), np.float64)
    ptrs = (c_vp * len(dcols))(*[c_vp(d.ptr if isinstance(d, DevBuf) else d) for d in dcols])
    ctx._check(ctx.lib.mg_hist(ctx.h, c_int(len(dcols)), ptrs, c_i64(n), _p(allb), _p(nb), c_int(int(include_lowest)), c_vp(counts.ptr),
                               c_vp(stream)))


def quantiles(ctx, dcol, n, percentiles, stream=None, first=0):
    """Exact gquantiles of n values of a device column from element `first` on: (quantiles, number of non-NaN values)."""
    p = _arr(percentiles, np.float64)
    out = np.empty(len(p), dtype=np.float64)
    nk = c_i64()
    ctx._check(ctx.lib.mg_quantiles(ctx.h, c_vp(dcol.ptr + 8 * int(first)), c_i64(n), ctypes.c_int(len(p)), _p(p), _p(out), ctypes.byref(nk),
                                    c_vp(stream)))
    return out, int(nk.value)


def quantiles_segmented(ctx, dcol, seg_first, percentiles, stream=None):
    """mg_quantiles_segmented: the quantiles of every segment [seg_first[s], seg_first[s+1]) of a device column in at most two
    launches. Returns (out[len(percentiles), nseg], number of non-NaN values per segment)."""
    p = _arr(percentiles, np.float64)
    sf = _arr(seg_first, np.int64)
    nseg = len(sf) - 1
    out = np.empty((len(p), max(nseg, 0)), dtype=np.float64)
    nk = np.empty(max(nseg, 0), dtype=np.int64)
    ctx._check(ctx.lib.mg_quantiles_segmented(ctx.h, c_vp(dcol.ptr if dcol is not None else None), c_i64(nseg), _p(sf), ctypes.c_int(len(p)),
                                              _p(p), _p(out), _p(nk), c_vp(stream)))
    return out, nk


def summary_segmented(ctx, dcol, seg_first, stream=None):
    """mg_summary_segmented: IntervalSummary of every segment of a device column; returns the nseg x 7 result rows."""
    sf = _arr(seg_first, np.int64)
    nseg = len(sf) - 1
    part = np.empty((max(nseg, 0), 6), dtype=np.float64)
    ctx._check(ctx.lib.mg_summary_segmented(ctx.h, c_vp(dcol.ptr if dcol is not None else None), c_i64(nseg), _p(sf), _p(part), c_vp(stream)))
    return summary_finalize_rows(part)


def _bins_args(dcols, breaks_list):
    nb = np.array([len(b) for b in breaks_list], dtype=np.int32)
    allb = _arr(np.concatenate([np.asarray(b, dtype=np.float64) for b in breaks_list]), np.float64)
    ptrs = (c_vp * len(dcols))(*[c_vp(d.ptr if isinstance(d, DevBuf) else d) for d in dcols])
    return nb, allb, ptrs, int(np.prod(nb - 1))


def bins_summary(ctx, dvals, dcols, n, breaks_list, include_lowest=False, stream=None, partials=False):
    """mg_bins_summary: total_bins x 7 summary rows (bins flattened first dimension fastest); partials=True returns the total_bins x 6
    IntervalSummary partials instead -- what dist.reduce_summary combines across ranks before summary_finalize_rows."""
    nb, allb, ptrs, total = _bins_args(dcols, breaks_list)
    part = np.empty((total, 6), dtype=np.float64)
    ctx._check(ctx.lib.mg_bins_summary(ctx.h, c_vp(dvals.ptr), c_int(len(dcols)), ptrs, c_i64(n), _p(allb), _p(nb), c_int(int(include_lowest)),
                                       _p(part), c_vp(stream)))
    return part if partials else summary_finalize_rows(part)


def bins_quantiles(ctx, dvals, dcols, n, breaks_list, percentiles, include_lowest=False, stream=None):
    """mg_bins_quantiles: (out[len(percentiles), total_bins], non-NaN values per bin)."""
    nb, allb, ptrs, total = _bins_args(dcols, breaks_list)
    p = _arr(percentiles, np.float64)
    out = np.empty((len(p), total), dtype=np.float64)
    nk = np.empty(total, dtype=np.int64)
    ctx._check(ctx.lib.mg_bins_quantiles(ctx.h, c_vp(dvals.ptr), c_int(len(dcols)), ptrs, c_i64(n), _p(allb), _p(nb), c_int(int(include_lowest)),
                                         c_int(len(p)), _p(p), _p(out), _p(nk), c_vp(stream)))
    return out, nk


def percentile_lookup(ctx, dcol, n, breaks, bins, stream=None):
    """mg_percentile_lookup: global.percentile* -- the device column mapped in place through the track's percentile table."""
    b, v = _arr(breaks, np.float64), _arr(bins, np.float64)
    if len(b) != len(v):
        raise MishaGpuError("percentile table: breaks and bins differ in length")
    ctx._check(ctx.lib.mg_percentile_lookup(ctx.h, c_vp(dcol.ptr), c_i64(n), _p(b), _p(v), c_int(len(b)), c_vp(stream)))


CMP_OPS = {">": 0, ">=": 1, "<": 2, "<=": 3, "==": 4, "!=": 5}


def screen_compare(ctx, cols, ops, thresholds, combine_or, n, dflag=None, stream=None):
    """d_flag[i] = 1 where every (any, with combine_or) comparison cols[k][i] ops[k] thresholds[k] holds; NaN -> 0."""
    dflag = ctx.alloc(n, np.int8) if dflag is None else dflag
    k = len(cols)
    cp = (ctypes.c_void_p * k)(*[c.ptr for c in cols])
    op = (ctypes.c_int * k)(*[CMP_OPS[o] if isinstance(o, str) else int(o) for o in ops])
    th = (ctypes.c_double * k)(*[float(t) for t in thresholds])
    ctx._check(ctx.lib.mg_screen_compare(ctx.h, ctypes.c_int(k), cp, op, th, ctypes.c_int(int(bool(combine_or))), c_i64(n), c_vp(dflag.ptr),
                                         c_vp(stream)))
    return dflag


def screen_merge(ctx, rows, dflag, first=0, count=None, stream=None, scratch=None):
    """Returns (chrom, start, end) host arrays of merged TRUE runs. `scratch` = (int32, int64, int64) device buffers of
    >= count elements to reuse across calls (allocated here otherwise)."""
    count = rows.n - first if count is None else count
    dc, ds, de = scratch if scratch is not None else (ctx.alloc(count, np.int32), ctx.alloc(count, np.int64), ctx.alloc(count, np.int64))
    nout = c_i64()
    ctx._check(ctx.lib.mg_screen_merge(ctx.h, rows.h, c_i64(first), c_i64(count), c_vp(dflag.ptr), c_vp(dc.ptr), c_vp(ds.ptr),
                                       c_vp(de.ptr), ctypes.byref(nout), c_vp(stream)))
    m = nout.value
    return dc.to_host(stream, m), ds.to_host(stream, m), de.to_host(stream, m)


def screen_merge_host(ctx, rows, dflag, host_out, first=0, count=None, stream=None):
    """The merged TRUE runs in the caller's host arrays host_out = (int32 chrom, int64 start, int64 end), e.g. Context.pinned buffers kept
    across calls: returns views of their first m entries, or None when the arrays are too short for the m runs found (m is then
    in screen_merge_host.last_count)."""
    count = rows.n - first if count is None else count
    hc, hs, he = host_out
    cap = min(len(hc), len(hs), len(he))
    nout = c_i64()
    ctx._check(ctx.lib.mg_h_screen_merge(ctx.h, rows.h, c_i64(first), c_i64(count), c_vp(dflag.ptr), _p(hc), _p(hs), _p(he), c_i64(cap),
                                         ctypes.byref(nout), c_vp(stream)))
    screen_merge_host.last_count = m = nout.value
    return None if m > cap else (hc[:m], hs[:m], he[:m])


# ---- several GPUs behind one host thread (mg_group_*) -------------------------------------------------------------------------
class Group:
    """One mg_ctx per device, driven from this thread: chromosomes staged on their owners, rows routed to them, the small
    gsummary / gdist / gquantiles partials combined inside the library (include/mishagpu.h, "several GPUs behind one host thread")."""

    def __init__(self, chrom_sizes, devices):
        self.lib = load()
        sizes = _arr(chrom_sizes, np.int64)
        dev = _arr(devices, np.int32)
        h = c_vp()
        if self.lib.mg_group_init(c_int(len(dev)), _p(dev), c_int(len(sizes)), _p(sizes), ctypes.byref(h)):
            raise MishaGpuError(self.lib.mg_group_last_error(None).decode())
        self.h, self.n, self.chrom_sizes = h, len(dev), sizes

    def _check(self, rc):
        if rc:
            raise MishaGpuError(self.lib.mg_group_last_error(self.h).decode())

    def owner(self, chromid):
        return int(self.lib.mg_group_owner(self.h, c_int(chromid)))

    def set_option(self, name, value):
        self._check(self.lib.mg_group_set_option(self.h, name.encode(), c_i64(int(value))))

    def launch_count(self):
        return sum(int(self.lib.mg_launch_count(c_vp(self.lib.mg_group_ctx(self.h, c_int(k))))) for k in range(self.n))

    def close(self):
        if getattr(self, "h", None):
            self.lib.mg_group_shutdown(self.h)
            self.h = None


class GroupDenseTrack:
    def __init__(self, group, bin_size):
        self.g = group
        h = c_vp()
        group._check(group.lib.mg_group_dense_create(group.h, ctypes.c_uint32(bin_size), ctypes.byref(h)))
        self.h = h

    def stage(self, chromid, bins):
        b = _arr(bins, np.float32)
        self.g._check(self.g.lib.mg_group_dense_stage(self.g.h, self.h, c_int(chromid), _p(b), c_i64(len(b))))

    def bytes(self, member):
        return int(self.g.lib.mg_group_dense_bytes(self.h, c_int(member)))

    def window_host(self, chrom, start, end, funcs, sshift=0, eshift=0, out=None):
        chrom, start, end = _arr(chrom, np.int32), _arr(start, np.int64), _arr(end, np.int64)
        ids, params = _func_args(funcs)
        n = len(chrom)
        out = out or [np.empty(n, dtype=np.float64) for _ in ids]
        ptrs = (c_vp * len(out))(*[o.ctypes.data_as(c_vp) for o in out])
        g = self.g
        g._check(g.lib.mg_group_h_dense_window(g.h, self.h, c_i64(n), _p(chrom), _p(start), _p(end), c_i64(sshift), c_i64(eshift),
                                               c_int(len(ids)), _p(ids), _p(params), ptrs))
        return out

    def _scope(self, schrom, sstart, send):
        return _arr(schrom, np.int32), _arr(sstart, np.int64), _arr(send, np.int64)

    def summary(self, binsize, schrom, sstart, send, func="avg", sshift=0, eshift=0):
        """gsummary over the fixed-bin iterator of the scope: the six partials (finalise with summary_finalize)."""
        c, s, e = self._scope(schrom, sstart, send)
        name, prm = (func, 0.0) if isinstance(func, str) else func
        out = np.empty(6, dtype=np.float64)
        g = self.g
        g._check(g.lib.mg_group_dense_summary(g.h, self.h, c_i64(binsize), c_i64(len(c)), _p(c), _p(s), _p(e), c_i64(sshift), c_i64(eshift),
                                              c_int(FUNC_IDS[name]), c_dbl(prm), _p(out)))
        return out

    def hist(self, binsize, schrom, sstart, send, breaks, func="avg", include_lowest=False, sshift=0, eshift=0):
        c, s, e = self._scope(schrom, sstart, send)
        name, prm = (func, 0.0) if isinstance(func, str) else func
        b = _arr(breaks, np.float64)
        out = np.zeros(len(b) - 1, dtype=np.uint64)
        g = self.g
        g._check(g.lib.mg_group_dense_hist(g.h, self.h, c_i64(binsize), c_i64(len(c)), _p(c), _p(s), _p(e), c_i64(sshift), c_i64(eshift),
                                           c_int(FUNC_IDS[name]), c_dbl(prm), _p(b), c_int(len(b)), c_int(int(include_lowest)), _p(out)))
        return out

    def quantiles(self, binsize, schrom, sstart, send, percentiles, func="avg", sshift=0, eshift=0):
        c, s, e = self._scope(schrom, sstart, send)
        name, prm = (func, 0.0) if isinstance(func, str) else func
        p = _arr(percentiles, np.float64)
        out = np.empty(len(p), dtype=np.float64)
        nk = c_i64()
        g = self.g
        g._check(g.lib.mg_group_dense_quantiles(g.h, self.h, c_i64(binsize), c_i64(len(c)), _p(c), _p(s), _p(e), c_i64(sshift), c_i64(eshift),
                                                c_int(FUNC_IDS[name]), c_dbl(prm), c_int(len(p)), _p(p), _p(out), ctypes.byref(nk)))
        return out, int(nk.value)

    def free(self):
        if self.h:
            self.g.lib.mg_group_dense_free(self.g.h, self.h)
            self.h = None
