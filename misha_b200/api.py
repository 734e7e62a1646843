"""Host-side mirror of misha's R interface for the hot path -- same names, argument meaning and error behaviour
(`gextract`, `gscreen`, `gsummary`, `gdist`, `gvtrack.create`, `gvtrack.iterator`, `gintervals`), so the parity tests read
like the reference's own tests. It holds NO arithmetic: it reads the reference's on-disk formats,
stages tracks once, turns scope + iterator into device rows and calls the C ABI (include/mishagpu.h). Track expressions
other than a bare (virtual) track name are evaluated on the returned vectors with numpy, as R evaluates them on the
1000-row batches (src/TrackExpressionScanner.cpp:352-368).

R reference points: gextract R/compute-core.R:77, gsummary :286, gscreen R/compute-utils.R:216, gdist
R/compute-distribution.R:152, gvtrack.create R/vtrack.R:1204, gvtrack.iterator :1370; C++ entry points
src/GenomeTrackExtract.cpp:518, src/GenomeTrackScreener.cpp:167, src/GenomeTrackSummary.cpp:119,
src/GenomeTrackDistribution.cpp:19.
"""
import os
import re

import numpy as np

from . import formats, gpu

FUSED_FUNCS = ("avg", "sum", "min", "max", "stddev", "quantile", "nearest", "size", "exists")
# "<f>.pos.abs" / "<f>.pos.relative" (R/vtrack.R; src/TrackVarProcessor.cpp:113-173): one device function, param = relative
POS_FUNCS = {"%s.pos.%s" % (f, k): ("%s.pos" % f, float(k == "relative")) for f in ("first", "last", "min", "max") for k in ("abs", "relative")}
TRACK_FUNCS = FUSED_FUNCS + ("lse", "first", "last") + tuple(POS_FUNCS)


def _gpu_func(vt):
    """The (name, param) the C ABI takes for a track vtrack's function."""
    if vt.func == "quantile":
        return ("quantile", float(vt.params))
    return POS_FUNCS.get(vt.func, vt.func)
PWM_FUNCS = ("pwm", "pwm.max", "pwm.max.pos", "pwm.count")
KMER_FUNCS = ("kmer.count", "kmer.frac")
MASKED_FUNCS = ("masked.count", "masked.frac")


# What rerror()/verror() raise in R (src/rdbutils.cpp:708-733): host-side validation and C-ABI failures are one error type.
MishaError = gpu.MishaGpuError


def _split_top(expr, ops):
    """Split at the top-level (parenthesis depth 0) occurrences of the one-character operators in `ops`."""
    parts, depth, cur = [], 0, []
    i = 0
    while i < len(expr):
        ch = expr[i]
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if depth == 0 and ch in ops:
            parts.append("".join(cur))
            cur = []
            if i + 1 < len(expr) and expr[i + 1] == ch:  # && and || behave like & and | on the vectors we hold
                i += 1
        else:
            cur.append(ch)
        i += 1
    parts.append("".join(cur))
    return parts


def _r_logic_to_python(expr):
    """R gives comparison operators higher precedence than !, & and |; Python's ~, & and | bind tighter than comparisons.
    Re-parenthesise so that numpy evaluates what R would: a > 1 & !b < 2  ->  ((a > 1)) & (~(b < 2))."""
    out, i, n = [], 0, len(expr)
    while i < n:  # first rewrite the inside of every parenthesised group
        if expr[i] == "(":
            depth, j = 1, i + 1
            while j < n and depth:
                depth += expr[j] == "("
                depth -= expr[j] == ")"
                j += 1
            out.append("(" + _r_logic_to_python(expr[i + 1:j - 1]) + ")")
            i = j
        else:
            out.append(expr[i])
            i += 1
    expr = "".join(out)
    ors = _split_top(expr, "|")
    if len(ors) > 1:
        return " | ".join("(" + _r_logic_to_python(o) + ")" for o in ors)
    ands = _split_top(expr, "&")
    if len(ands) > 1:
        return " & ".join("(" + _r_logic_to_python(a) + ")" for a in ands)
    t = expr.strip()
    if t.startswith("!") and not t.startswith("!="):
        return "~(" + _r_logic_to_python(t[1:]) + ")"
    return expr


# ---- R's vectors on the host side of an expression --------------------------------------------------------------------------
# numpy compares NaN as FALSE and negates that to TRUE; R compares NA as NA, and NA is never TRUE: !x > c and x != c over a NaN
# operand must not select the row (gscreen keeps TRUE rows only, src/GenomeTrackScreener.cpp:185-200). The expression is therefore
# evaluated over these two wrappers: numeric vectors whose comparisons yield three-valued logicals, combined by R's & | !.
class RLogical:
    """TRUE where t, NA where na, FALSE elsewhere."""

    def __init__(self, t, na=None):
        t = np.asarray(t, dtype=bool)
        self.na = np.zeros(t.shape, dtype=bool) if na is None else np.asarray(na, dtype=bool)
        self.t = t & ~self.na

    @property
    def f(self):
        return ~self.t & ~self.na

    def __and__(self, o):
        o = _as_logical(o)
        t, f = self.t & o.t, self.f | o.f
        return RLogical(t, ~t & ~f)

    __rand__ = __and__

    def __or__(self, o):
        o = _as_logical(o)
        t, f = self.t | o.t, self.f & o.f
        return RLogical(t, ~t & ~f)

    __ror__ = __or__

    def __invert__(self):
        return RLogical(self.f, self.na)

    def true(self):
        """The rows a filter keeps."""
        return self.t

    def as_numeric(self):
        """What R stores when a logical lands in a numeric column: 1 / 0 / NA."""
        return np.where(self.na, np.nan, self.t.astype(np.float64))


def _as_logical(x):
    if isinstance(x, RLogical):
        return x
    if isinstance(x, RNumeric):
        return RLogical(x.v != 0, np.isnan(x.v))
    return RLogical(np.asarray(x, dtype=bool))


class RNumeric:
    def __init__(self, v):
        self.v = np.asarray(v.v if isinstance(v, RNumeric) else v, dtype=np.float64)

    @staticmethod
    def _val(o):
        if isinstance(o, RNumeric):
            return o.v
        if isinstance(o, RLogical):
            return o.as_numeric()
        return np.asarray(o, dtype=np.float64)

    def _cmp(self, o, op):
        b = self._val(o)
        with np.errstate(invalid="ignore"):
            return RLogical(op(self.v, b), np.isnan(self.v) | np.isnan(b))

    def __lt__(self, o): return self._cmp(o, np.less)
    def __le__(self, o): return self._cmp(o, np.less_equal)
    def __gt__(self, o): return self._cmp(o, np.greater)
    def __ge__(self, o): return self._cmp(o, np.greater_equal)
    def __eq__(self, o): return self._cmp(o, np.equal)  # noqa: E704
    def __ne__(self, o): return self._cmp(o, np.not_equal)
    __hash__ = None

    def _bin(self, o, op, swap=False):
        a, b = (self._val(o), self.v) if swap else (self.v, self._val(o))
        with np.errstate(all="ignore"):
            return RNumeric(op(a, b))

    def __add__(self, o): return self._bin(o, np.add)
    def __radd__(self, o): return self._bin(o, np.add, True)
    def __sub__(self, o): return self._bin(o, np.subtract)
    def __rsub__(self, o): return self._bin(o, np.subtract, True)
    def __mul__(self, o): return self._bin(o, np.multiply)
    def __rmul__(self, o): return self._bin(o, np.multiply, True)
    def __truediv__(self, o): return self._bin(o, np.divide)
    def __rtruediv__(self, o): return self._bin(o, np.divide, True)
    def __pow__(self, o): return self._bin(o, np.power)
    def __rpow__(self, o): return self._bin(o, np.power, True)
    def __mod__(self, o): return self._bin(o, np.mod)
    def __neg__(self): return RNumeric(-self.v)
    def __pos__(self): return self
    def __and__(self, o): return _as_logical(self) & o
    def __or__(self, o): return _as_logical(self) | o
    def __invert__(self): return ~_as_logical(self)


def _rfun(f):
    def g(*a):
        with np.errstate(all="ignore"):
            return RNumeric(f(*[RNumeric._val(x) for x in a]))
    return g


R_FUNCS = {"log": _rfun(np.log), "log2": _rfun(np.log2), "log10": _rfun(np.log10), "exp": _rfun(np.exp), "sqrt": _rfun(np.sqrt), "abs": _rfun(np.abs),
           "pmin": _rfun(np.minimum), "pmax": _rfun(np.maximum), "floor": _rfun(np.floor), "ceiling": _rfun(np.ceil),
           "is_na": lambda x: RLogical(np.isnan(RNumeric._val(x))), "NaN": np.nan}


def r_eval(py, variables):
    """Evaluates the rewritten expression over R-like vectors. Returns an RLogical, an RNumeric or a Python scalar."""
    env = {k: RNumeric(v) for k, v in variables.items()}
    return eval(py, {"__builtins__": {}}, {**R_FUNCS, **env})  # noqa: S307 - the expression language of the host


def _value_column(iv):
    """Name of the first numeric column of an intervals set that is not a coordinate column, or None."""
    for k, v in iv.items():
        if k in ("chrom", "start", "end", "strand", "intervalID"):
            continue
        if np.issubdtype(np.asarray(v).dtype, np.number):
            return k
    return None


def rows_is_fixedbin(rows):
    return getattr(rows, "binsize", None) is not None


class Intervals(dict):
    """A 1-D intervals set: columns chrom (names), start, end (+ anything else). Mirrors the R data frame."""

    def __init__(self, chrom, start, end, **extra):
        super().__init__(chrom=np.asarray(chrom, dtype=object), start=np.asarray(start, dtype=np.int64), end=np.asarray(end, dtype=np.int64))
        for k, v in extra.items():
            self[k] = np.asarray(v)

    def __len__(self):
        return len(self["start"])


class _VTrack:
    def __init__(self, name, src, func, params):
        self.name, self.src, self.func, self.params = name, src, func, params
        self.sshift = self.eshift = 0
        self.slice, self.slice_func, self.slice_percentile = None, "avg", 0.5  # gvtrack.array.slice (array tracks only)


class MishaDB:
    """One genome database root + one GPU context (the `.misha` environment of an R session)."""

    def __init__(self, groot, device=0):
        self.groot = groot
        path = os.path.join(groot, "chrom_sizes.txt")
        if not os.path.exists(path):
            raise MishaError("%s is not a misha database root (no chrom_sizes.txt)" % groot)
        chroms = []
        for line in open(path):
            f = line.split()
            if len(f) >= 2:
                name = f[0] if f[0].startswith("chr") else "chr" + f[0]  # R/db-root.R:86-105
                chroms.append((name, int(f[1])))
        chroms.sort(key=lambda c: c[0])  # per-chromosome DB: alphabetical order (R/db-root.R:157-162)
        self.chroms = chroms
        self.chrom_names = [c for c, _ in chroms]
        self.chrom_sizes = np.array([s for _, s in chroms], dtype=np.int64)
        self.chrom_id = {c: i for i, c in enumerate(self.chrom_names)}
        for c, i in list(self.chrom_id.items()):
            self.chrom_id[c[3:]] = i  # un-prefixed aliases
        self.ctx = gpu.Context(self.chrom_sizes, device)
        self.vtracks = {}
        self._dense, self._sparse, self._ivsets, self._seq = {}, {}, {}, None
        self._track_idx = {}

    def close(self):
        self.ctx.close()

    # ---- intervals ----------------------------------------------------------------------------------------
    def gintervals(self, chroms, starts=0, ends=-1):
        chroms = np.atleast_1d(np.asarray(chroms, dtype=object))
        starts = np.broadcast_to(np.atleast_1d(starts), chroms.shape).astype(np.int64)
        ends = np.broadcast_to(np.atleast_1d(ends), chroms.shape).astype(np.int64).copy()
        names = []
        for k, c in enumerate(chroms):
            key = str(c)
            if key not in self.chrom_id:
                raise MishaError("Chromosome %s does not exist" % key)
            cid = self.chrom_id[key]
            names.append(self.chrom_names[cid])
            if ends[k] == -1:
                ends[k] = self.chrom_sizes[cid]
        iv = Intervals(names, starts, ends)
        self._validate(iv)
        return iv

    def gintervals_all(self):
        return Intervals(self.chrom_names, np.zeros(len(self.chroms), np.int64), self.chrom_sizes.copy())

    def _validate(self, iv):
        cid = self._cids(iv)
        s, e = iv["start"], iv["end"]
        bad = (s < 0) | (s >= e) | (e > self.chrom_sizes[cid])
        if bad.any():
            k = int(np.flatnonzero(bad)[0])
            raise MishaError("Invalid interval (%s, %d, %d)" % (iv["chrom"][k], s[k], e[k]))  # src/IntervalConverter.cpp:200-360

    def _cids(self, iv):
        try:
            return np.array([self.chrom_id[str(c)] for c in iv["chrom"]], dtype=np.int32)
        except KeyError as e:
            raise MishaError("Chromosome %s does not exist" % e.args[0])

    # ---- tracks -------------------------------------------------------------------------------------------
    def _track_dir(self, track):
        return os.path.join(self.groot, "tracks", track.replace(".", "/") + ".track")

    def gtrack_exists(self, track):
        return os.path.isdir(self._track_dir(track))

    def gtrack_ls(self):
        out = []
        root = os.path.join(self.groot, "tracks")
        for dp, dn, _ in os.walk(root):
            for d in dn:
                if d.endswith(".track"):
                    out.append(os.path.relpath(os.path.join(dp, d), root)[:-6].replace(os.sep, "."))
        return sorted(out)

    def _track_index(self, track):
        """The track's single-file index (track.idx, src/TrackIndex.h) or None for the per-chromosome layout; cached per track."""
        if track not in self._track_idx:
            try:
                self._track_idx[track] = formats.read_track_index(self._track_dir(track))
            except formats.FormatError as e:
                raise MishaError(str(e))
        return self._track_idx[track]

    def _track_type(self, track):
        d = self._track_dir(track)
        if not os.path.isdir(d):
            raise MishaError("Track %s does not exist" % track)
        try:
            t = formats.track_type(d, self.chrom_names)
        except formats.FormatError as e:
            raise MishaError("Track %s: %s" % (track, e))
        if t is None:
            raise MishaError("Track %s does not exist" % track)
        return t

    def _dense_track(self, track):
        if track not in self._dense:
            d = self._track_dir(track)
            idx = self._track_index(track)
            t = None
            for cid, c in enumerate(self.chrom_names):
                got = formats.read_dense_chrom(d, c, cid, idx)
                if got is None:
                    if idx is None:
                        raise MishaError("Track %s: missing chromosome file %s" % (track, c))
                    continue  # an indexed track may hold empty contigs (src/GenomeTrackFixedBin.cpp:1178-1199): their rows are NaN
                bin_size, bins = got
                if t is None:
                    t = gpu.DenseTrack(self.ctx, bin_size)
                elif t.bin_size != bin_size:
                    raise MishaError("Track %s: inconsistent bin size" % track)
                t.stage(cid, bins)
            if t is None:
                raise MishaError("Track %s holds no data" % track)
            self._dense[track] = t
        return self._dense[track]

    def _sparse_records(self, track, cid):
        return formats.read_sparse_chrom(self._track_dir(track), self.chrom_names[cid], cid, self._track_index(track))

    def _sparse_track(self, track):
        if track not in self._sparse:
            t = gpu.SparseTrack(self.ctx)
            for cid in range(len(self.chroms)):
                t.stage(cid, *self._sparse_records(track, cid))
            self._sparse[track] = t
        return self._sparse[track]

    def _sequence(self):
        if self._seq is None:
            s = gpu.Sequence(self.ctx)
            seq_dir = os.path.join(self.groot, "seq")
            try:
                gidx = formats.read_genome_index(seq_dir)  # genome.idx + genome.seq (src/GenomeSeqFetch.cpp:66-107) or <chrom>.seq files
                for cid, c in enumerate(self.chrom_names):
                    s.stage(cid, formats.read_seq_chrom(seq_dir, c, cid, gidx))
            except formats.FormatError as e:
                raise MishaError(str(e))
            self._seq = s
        return self._seq

    def _ivset(self, vt):
        if vt.name not in self._ivsets:
            iv = vt.src
            cid = self._cids(iv)
            s = gpu.IntervalSet(self.ctx)
            for c in range(len(self.chroms)):
                m = cid == c
                st, en = iv["start"][m], iv["end"][m]
                order = np.argsort(st, kind="stable")
                st, en = st[order], en[order]
                # sort + unify_overlaps (src/TrackExpressionVars.cpp:1299-1301, src/GIntervals.cpp:59-74)
                us, ue = [], []
                for a, b in zip(st, en):
                    if us and a <= ue[-1]:
                        ue[-1] = max(ue[-1], b)
                    else:
                        us.append(a)
                        ue.append(b)
                s.stage(c, np.array(us, np.int64), np.array(ue, np.int64))
            self._ivsets[vt.name] = s
        return self._ivsets[vt.name]

    def _array_track(self, track, slice_cols, slice_func, slice_percentile):
        """An array track under a slice (gvtrack.array.slice: columns + the function that folds them, R/vtrack.R; src/GenomeTrackArrays.cpp):
        every record presents ONE value -- get_sliced_val over its stored (value, column) pairs -- and from there on the track is a sparse
        track (calc_vals, src/GenomeTrackArrays.h:131-196, is GenomeTrackSparse's arithmetic). The fold runs on the host at staging time."""
        key = "\0array:%s:%s:%s:%r" % (track, ",".join(map(str, slice_cols)) if slice_cols is not None else "*", slice_func, slice_percentile)
        if key not in self._sparse:
            d, idx = self._track_dir(track), self._track_index(track)
            t = gpu.SparseTrack(self.ctx)
            try:
                for cid, c in enumerate(self.chrom_names):
                    st, en, rf, v, cols = formats.read_array_chrom(d, c, cid, idx)
                    t.stage(cid, st, en, formats.array_slice_values(rf, v, cols, slice_cols, slice_func, slice_percentile))
            except formats.FormatError as e:
                raise MishaError("Track %s: %s" % (track, e))
            self._sparse[key] = t
        return self._sparse[key]

    def gvtrack_array_slice(self, vtrack, slice=None, func="avg", params=None):
        """gvtrack.array.slice(vtrack, slice, func, params), R/vtrack.R: which columns (0-based indices) of the array track a virtual track
        reads and how they are folded into one value per record: avg / min / max / stddev / sum / quantile (params = percentile)."""
        if vtrack not in self.vtracks:
            raise MishaError("Virtual track %s does not exist" % vtrack)
        vt = self.vtracks[vtrack]
        if not isinstance(vt.src, str) or self._track_type(vt.src) != "array":
            raise MishaError("Virtual track %s: slices apply to array tracks only" % vtrack)
        if func not in formats.SLICE_FUNCS:
            raise MishaError("Virtual track %s: invalid slice function %s" % (vtrack, func))
        if func == "quantile" and (params is None or not (0 <= float(params) <= 1)):
            raise MishaError("Virtual track %s: slice function quantile requires a percentile in [0, 1]" % vtrack)
        vt.slice = None if slice is None else sorted({int(x) for x in np.atleast_1d(slice)})
        if vt.slice is not None and (not vt.slice or vt.slice[0] < 0):
            raise MishaError("Virtual track %s: invalid slice" % vtrack)
        vt.slice_func, vt.slice_percentile = func, float(params) if params is not None else 0.5

    def _value_track(self, vt):
        """The value-based source of a vtrack staged like a sparse track: sorted records, values narrowed to float (NA -> NaN)."""
        key = "\0value:" + vt.name
        if key not in self._sparse:
            iv = vt.src
            cid = self._cids(iv)
            vals = np.asarray(iv[_value_column(iv)], dtype=np.float64).astype(np.float32)
            t = gpu.SparseTrack(self.ctx)
            for c in range(len(self.chroms)):
                m = np.flatnonzero(cid == c)
                m = m[np.argsort(iv["start"][m], kind="stable")]
                t.stage(c, iv["start"][m], iv["end"][m], vals[m])
            self._sparse[key] = t
        return self._sparse[key]

    # ---- virtual tracks -----------------------------------------------------------------------------------
    def gvtrack_create(self, vtrack, src=None, func=None, params=None):
        """R/vtrack.R:1204. src: track name, Intervals (for 'coverage'), or None for sequence functions (pwm*)."""
        if not re.match(r"^[A-Za-z][A-Za-z0-9_.]*$", vtrack):
            raise MishaError("Invalid virtual track name %s" % vtrack)
        if isinstance(src, str):
            if not self.gtrack_exists(src):
                raise MishaError("Track %s does not exist" % src)
            func = func or "avg"
            if func not in TRACK_FUNCS:
                raise MishaError("Virtual track %s: function %s is not supported by this build" % (vtrack, func))
            if func == "quantile":
                if params is None or not (0 <= float(params) <= 1):
                    raise MishaError("Virtual track %s: function quantile requires a percentile in [0, 1]" % vtrack)
        elif isinstance(src, Intervals) and _value_column(src) is not None and (func or "avg") not in ("coverage", "distance", "distance.center", "neighbor.count"):
            # a VALUE-BASED track: intervals with one numeric value column act as an in-memory sparse track (GenomeTrackInMemory,
            # src/GenomeTrackInMemory.cpp:36-270 = GenomeTrackSparse's count-based arithmetic; detection as in
            # src/TrackExpressionVars.cpp:906-976: the first numeric column that is not chrom / start / end / strand / intervalID)
            func = func or "avg"
            if func not in TRACK_FUNCS:
                raise MishaError("Virtual track %s: invalid function %s used with value-based intervals" % (vtrack, func))
            if func == "quantile" and (params is None or not (0 <= float(params) <= 1)):
                raise MishaError("Virtual track %s: function quantile requires a percentile in [0, 1]" % vtrack)
            self._validate(src)
            cid = self._cids(src)
            order = np.lexsort((src["start"], cid))
            c, st, en = cid[order], src["start"][order], src["end"][order]
            if np.any((c[1:] == c[:-1]) & (st[1:] < en[:-1])):
                raise MishaError("Virtual track %s: value-based intervals must not overlap" % vtrack)
        elif isinstance(src, Intervals):
            func = func or "coverage"
            if func != "coverage":
                raise MishaError("Virtual track %s: function %s is not supported for interval sources by this build" % (vtrack, func))
        elif src is None and func in MASKED_FUNCS:  # no parameters (R/vtrack.R)
            params = {}
        elif src is None and func in KMER_FUNCS:  # .vtrack_params_kmer, R/vtrack.R:292-329
            p = dict(params) if isinstance(params, dict) else {"kmer": params}
            unknown = set(p) - {"kmer", "extend", "strand"}
            if unknown:
                raise MishaError("Virtual track %s: unknown parameter(s) for %s: %s" % (vtrack, func, ", ".join(sorted(unknown))))
            if not isinstance(p.get("kmer"), str):
                raise MishaError("kmer parameter must be a single string")
            if not p["kmer"]:
                raise MishaError("kmer sequence cannot be empty")
            params = dict(kmer=p["kmer"], extend=bool(p.get("extend", True)), strand=int(p.get("strand", 0)))
        elif src is None:
            if func not in PWM_FUNCS:
                raise MishaError("Virtual track %s: a NULL source requires a sequence function (pwm, pwm.max, pwm.max.pos, pwm.count, kmer.count, kmer.frac, masked.count, masked.frac)" % vtrack)
            p = dict(params) if isinstance(params, dict) else {"pssm": params}
            pssm = np.asarray(p.get("pssm"), dtype=np.float64)
            if pssm.ndim != 2 or pssm.shape[1] != 4:
                raise MishaError("Virtual track %s: PWM functions require a matrix with columns A, C, G, T" % vtrack)
            prior = float(p.get("prior", 0.01))  # defaults: R/vtrack.R:179-290
            if not (0 <= prior <= 1):
                raise MishaError("Virtual track %s: prior must be between 0 and 1" % vtrack)
            spat, spat_bin = p.get("spat_factor"), p.get("spat_bin")
            if spat is not None:  # R/vtrack.R:249-262
                spat = np.atleast_1d(np.asarray(spat, dtype=np.float64))
                if np.any(~(spat > 0)):
                    raise MishaError("spat_factor must be a numeric vector with all positive values")
                spat_bin = 1 if spat_bin is None else spat_bin
                if not spat_bin > 0:
                    raise MishaError("spat_bin must be a positive integer")
                if func == "pwm.max.pos":
                    raise MishaError("Virtual track %s: pwm.max.pos under spatial weights is not supported by this build" % vtrack)
                if p.get("spat_min") is not None or p.get("spat_max") is not None:
                    raise MishaError("Virtual track %s: spat_min / spat_max are not supported by this build" % vtrack)
            params = dict(pssm=pssm, bidirect=bool(p.get("bidirect", True)), prior=prior, extend=bool(p.get("extend", True)),
                          strand=int(p.get("strand", 1)), score_thresh=float(p.get("score.thresh", p.get("score_thresh", 0.0))),
                          spat_factor=None if spat is None else spat.astype(np.float32), spat_bin=int(spat_bin or 1))
        else:
            raise MishaError("Virtual track %s: unsupported source" % vtrack)
        self.vtracks[vtrack] = _VTrack(vtrack, src, func, params)
        self._ivsets.pop(vtrack, None)

    def gvtrack_iterator(self, vtrack, sshift=0, eshift=0):
        """R/vtrack.R:1370 (1-D form)."""
        if vtrack not in self.vtracks:
            raise MishaError("Virtual track %s does not exist" % vtrack)
        self.vtracks[vtrack].sshift, self.vtracks[vtrack].eshift = int(sshift), int(eshift)

    def gvtrack_rm(self, vtrack):
        self.vtracks.pop(vtrack, None)
        self._ivsets.pop(vtrack, None)

    def gvtrack_ls(self):
        return sorted(self.vtracks)

    # ---- scanner ------------------------------------------------------------------------------------------
    def _expr_vars(self, expr):
        """Names of tracks / virtual tracks an expression mentions (longest match first, as identifiers)."""
        names = sorted(set(self.gtrack_ls()) | set(self.vtracks), key=len, reverse=True)
        found = []
        for n in names:
            if re.search(r"(?<![A-Za-z0-9_.])" + re.escape(n) + r"(?![A-Za-z0-9_.])", expr):
                found.append(n)
        return found

    def _var_track(self, name):
        vt = self.vtracks.get(name)
        return vt.src if vt is not None and isinstance(vt.src, str) else (name if vt is None else None)

    def _rows(self, exprs, intervals, iterator, unify_scope):
        self._validate(intervals)
        cid = self._cids(intervals)
        order = np.lexsort((intervals["start"], cid))  # stable sort by (chromid, start), src/GenomeTrackExtract.cpp:557-562
        sc, ss, se = cid[order], intervals["start"][order], intervals["end"][order]
        if unify_scope and len(sc):  # gsummary / gdist / gscreen: sort + unify_overlaps
            keep_c, keep_s, keep_e = [sc[0]], [ss[0]], [se[0]]
            for c, a, b in zip(sc[1:], ss[1:], se[1:]):
                if c != keep_c[-1] or keep_e[-1] < a:
                    keep_c.append(c); keep_s.append(a); keep_e.append(b)
                elif keep_e[-1] < b:
                    keep_e[-1] = b
            sc, ss, se = np.array(keep_c, np.int32), np.array(keep_s, np.int64), np.array(keep_e, np.int64)
            order = np.arange(len(sc))
        if iterator is None:
            tracks = []
            for e in exprs:
                for v in self._expr_vars(e):
                    t = self._var_track(v)
                    if t is not None:
                        tracks.append(t)
            if not tracks:
                raise MishaError("Cannot implicitly determine iterator policy: track expressions do not contain any tracks.")
            types = {self._track_type(t) for t in tracks}
            if types == {"dense"}:
                sizes = {self._dense_track(t).bin_size for t in tracks}
                if len(sizes) != 1:
                    raise MishaError("Cannot implicitly determine iterator policy: track expressions contain tracks with different bin sizes.")
                iterator = sizes.pop()
            elif types == {"sparse"} and len(set(tracks)) == 1:
                chroms, st, en = [], [], []
                for c in range(len(self.chroms)):
                    a, b, _ = self._sparse_records(tracks[0], c)
                    chroms.append(np.full(len(a), c, np.int32)); st.append(a); en.append(b)
                iterator = ("records", np.concatenate(chroms), np.concatenate(st), np.concatenate(en))
            else:
                raise MishaError("Cannot implicitly determine iterator policy")
        if isinstance(iterator, (int, np.integer)):
            if iterator <= 0:
                raise MishaError("Bin size of a fixed bin iterator (%d) must be positive" % iterator)
            rows = self.ctx.rows_fixedbin(int(iterator), sc, ss, se)
        else:
            if isinstance(iterator, Intervals):
                self._validate(iterator)
                ic, is_, ie = self._cids(iterator), iterator["start"], iterator["end"]
            else:
                _, ic, is_, ie = iterator
            o = np.lexsort((is_, ic))
            ic, is_, ie = ic[o], is_[o], ie[o]
            if len(ic):  # unify_overlaps(false): touching intervals stay apart (src/TrackExpressionScanner.cpp:948)
                kc, ks, ke = [ic[0]], [is_[0]], [ie[0]]
                for c, a, b in zip(ic[1:], is_[1:], ie[1:]):
                    if c != kc[-1] or ke[-1] <= a:
                        kc.append(c); ks.append(a); ke.append(b)
                    elif ke[-1] < b:
                        ke[-1] = b
                ic, is_, ie = np.array(kc, np.int32), np.array(ks, np.int64), np.array(ke, np.int64)
            rows = self.ctx.rows_intervals(ic, is_, ie, sc, ss, se)
        return rows, order

    def _var_column(self, name, rows):
        """Device column (DevBuf) of one track / vtrack variable over the rows."""
        vt = self.vtracks.get(name)
        if vt is None:  # a bare track name: avg (src/TrackExpressionVars.cpp:1543-1546)
            vt = _VTrack(name, name, "avg", None)
        if isinstance(vt.src, str):
            f = _gpu_func(vt)
            kind = self._track_type(vt.src)
            if kind == "array":
                t = self._array_track(vt.src, vt.slice, vt.slice_func, vt.slice_percentile)
            else:
                t = self._dense_track(vt.src) if kind == "dense" else self._sparse_track(vt.src)
            return t.window_dev(rows, [f], vt.sshift, vt.eshift)[0]
        if isinstance(vt.src, Intervals) and vt.func != "coverage":  # value-based: the in-memory sparse track
            return self._value_track(vt).window_dev(rows, [_gpu_func(vt)], vt.sshift, vt.eshift)[0]
        if isinstance(vt.src, Intervals):
            return self._ivset(vt).coverage_dev(rows, vt.sshift, vt.eshift)
        p = vt.params
        if vt.func in MASKED_FUNCS:
            return self._sequence().masked_dev(rows, mode=vt.func, sshift=vt.sshift, eshift=vt.eshift)
        if vt.func in KMER_FUNCS:
            return self._sequence().kmer_dev(rows, p["kmer"], mode=vt.func, extend=p["extend"], strand=p["strand"], sshift=vt.sshift, eshift=vt.eshift)
        logp = gpu.pssm_logp(p["pssm"], p["prior"])
        return self._sequence().pwm_dev(rows, logp, bidirect=p["bidirect"], extend=p["extend"], mode=vt.func, strand=p["strand"],
                                        score_thresh=p["score_thresh"], sshift=vt.sshift, eshift=vt.eshift, spat_factor=p.get("spat_factor"),
                                        spat_bin=p.get("spat_bin", 1))

    def _eval(self, expr, rows, cache, logical=False):
        """Host value vector of a track expression over the rows. logical: a logical result is returned as RLogical (R's three-valued
        vector) instead of being stored as 1 / 0 / NA."""
        expr = expr.strip()
        names = self._expr_vars(expr)
        if not names:
            raise MishaError("Track expression \"%s\" does not contain any tracks or virtual tracks" % expr)
        env = {}
        for n in names:
            if n not in cache:
                cache[n] = self._var_column(n, rows).to_host()
            env[n] = cache[n]
        if expr in env:
            return env[expr]
        py = expr
        alias = {}
        for k, n in enumerate(names):  # names may contain dots: substitute safe identifiers
            a = "__v%d" % k
            py = re.sub(r"(?<![A-Za-z0-9_.])" + re.escape(n) + r"(?![A-Za-z0-9_.])", a, py)
            alias[a] = env[n]
        py = _r_logic_to_python(py).replace("^", "**")
        py = re.sub(r"\bTRUE\b", "True", re.sub(r"\bFALSE\b", "False", py))
        py = py.replace("is.na(", "is_na(")
        try:
            out = r_eval(py, alias)
        except Exception as e:  # noqa: BLE001
            raise MishaError("Evaluation of track expression \"%s\" failed: %s" % (expr, e))
        if isinstance(out, RLogical):
            return out if logical else out.as_numeric()
        if isinstance(out, RNumeric):
            out = out.v
        out = np.asarray(out)
        if out.shape == ():
            out = np.full(rows.n, out)
        return out

    # ---- entry points -------------------------------------------------------------------------------------
    def gtrack_create(self, track, expr, iterator=None, indexed=None):
        """gtrack.create(track, description, expr, iterator) -- src/GenomeTrackCreate.cpp:35-215: the expression evaluated over the
        whole genome; a fixed-bin iterator yields a dense track (uint32 bin size + one float32 per bin and chromosome, every
        chromosome written), an intervals iterator a sparse one (every iterator interval with its value, chromosomes without rows get
        a header-only file). The value is the expression's double narrowed to float, as write_next_bin / pack_record store it.
        indexed: write track.idx + track.dat (the reference does when the database is indexed; default: as the sequence is stored)."""
        if not re.match(r"^[A-Za-z][A-Za-z0-9_.]*$", track):
            raise MishaError("Invalid track name %s" % track)
        d = self._track_dir(track)
        if os.path.exists(d):
            raise MishaError("Track %s already exists" % track)
        scope = self.gintervals_all()
        rows, _ = self._rows([expr], scope, iterator, unify_scope=False)
        vals = np.asarray(self._eval(expr, rows, {}), dtype=np.float64).astype(np.float32)
        c, s, e, _ = rows.coords()
        kind = "dense" if isinstance(iterator, (int, np.integer)) or (iterator is None and rows_is_fixedbin(rows)) else "sparse"
        blobs = {}
        for cid in range(len(self.chrom_names)):
            m = c == cid
            if kind == "dense":
                bin_size = int(iterator) if iterator is not None else int(rows.binsize)
                if not m.any():
                    continue
                blobs[cid] = formats.dense_blob(vals[m], bin_size)
            else:
                blobs[cid] = formats.sparse_blob(s[m], e[m], vals[m])
        if indexed is None:
            indexed = os.path.exists(os.path.join(self.groot, "seq", "genome.idx"))
        if indexed:
            formats.write_indexed_track(d, kind, {cid: blobs.get(cid, b"") for cid in range(len(self.chrom_names))})
        else:
            os.makedirs(d)
            for cid, b in blobs.items():
                with open(os.path.join(d, self.chrom_names[cid]), "wb") as f:
                    f.write(b)
        self._track_idx.pop(track, None)
        return track

    def gtrack_rm(self, track):
        """gtrack.rm: the track's directory and whatever of it is staged."""
        import shutil
        d = self._track_dir(track)
        if not os.path.isdir(d):
            raise MishaError("Track %s does not exist" % track)
        for cache in (self._dense, self._sparse):
            t = cache.pop(track, None)
            if t is not None:
                t.free()
        self._track_idx.pop(track, None)
        shutil.rmtree(d)

    def gextract(self, *exprs, intervals=None, iterator=None, colnames=None):
        """R/compute-core.R:77. Returns a dict of columns: chrom, start, end, one per expression, intervalID."""
        if not exprs:
            raise MishaError("Usage: gextract([expr]+, intervals, colnames = NULL, iterator = NULL)")
        intervals = self.gintervals_all() if intervals is None else intervals
        if colnames is not None and len(colnames) != len(exprs):
            raise MishaError("Number of column names must match the number of track expressions")
        rows, order = self._rows(exprs, intervals, iterator, unify_scope=False)
        if rows.n == 0:
            return None  # R returns NULL
        c, s, e, k = rows.coords()
        res = {"chrom": np.array(self.chrom_names, dtype=object)[c], "start": s, "end": e}
        cache = {}
        for j, ex in enumerate(exprs):
            res[colnames[j] if colnames else ex] = np.asarray(self._eval(ex, rows, cache), dtype=np.float64)
        res["intervalID"] = (order[k] + 1).astype(np.int32)  # 1-based row of the scope data frame (src/IntervalConverter.cpp:357)
        return res

    def gscreen(self, expr, intervals=None, iterator=None):
        """R/compute-utils.R:216. Returns Intervals of merged TRUE runs (or None)."""
        intervals = self.gintervals_all() if intervals is None else intervals
        rows, _ = self._rows([expr], intervals, iterator, unify_scope=True)
        if rows.n == 0:
            return None
        flag = self._screen_flags_on_device(expr, rows)
        if flag is None:  # general expressions are evaluated by the host on the returned vectors, as in the reference
            v = self._eval(expr, rows, {}, logical=True)
            if not isinstance(v, RLogical):
                raise MishaError("Expression \"%s\" does not produce a logical result" % expr)  # src/TrackExpressionScanner.h:158-173
            flag = self.ctx.alloc(rows.n, np.int8).from_host(v.true().astype(np.int8))  # NA is not TRUE: the row is dropped
        c, s, e = gpu.screen_merge(self.ctx, rows, flag)
        if len(c) == 0:
            return None
        return Intervals(np.array(self.chrom_names, dtype=object)[c], s, e)

    _TERM = re.compile(r"^\s*([A-Za-z_.][A-Za-z0-9_.]*)\s*(>=|<=|==|!=|>|<)\s*([-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?)\s*$")

    def _screen_flags_on_device(self, expr, rows):
        """Predicates of the form `var op number` joined by all `&` or all `|`: the columns stay in HBM and only the flags'
        merge result leaves the device. None = not of that form."""
        for sep, is_or in (("&", False), ("|", True)):
            parts = expr.split(sep)
            if len(parts) > 8 or (sep == "|" and len(parts) == 1):
                continue
            terms = [self._TERM.match(p) for p in parts]
            if not all(terms):
                continue
            names = [t.group(1) for t in terms]
            if not all(n in self.vtracks or self.gtrack_exists(n) for n in names):
                return None
            cols = {}
            for n in names:
                if n not in cols:
                    cols[n] = self._var_column(n, rows)
            return gpu.screen_compare(self.ctx, [cols[n] for n in names], [t.group(2) for t in terms], [float(t.group(3)) for t in terms],
                                      is_or, rows.n)
        return None

    def gquantiles(self, expr, percentiles=0.5, intervals=None, iterator=None):
        """R/compute-distribution.R gquantiles -> C_gquantiles (src/GenomeTrackQuantiles.cpp:124-262): quantiles of the expression's
        values over all iterator rows. Returns a dict percentile -> value (the named vector R returns)."""
        pcts = np.atleast_1d(np.asarray(percentiles, dtype=np.float64))
        for p in pcts:
            if not (0 <= p <= 1):
                raise MishaError("Percentile (%g) is not in [0, 1] range\n" % p)
        intervals = self.gintervals_all() if intervals is None else intervals
        rows, _ = self._rows([expr], intervals, iterator, unify_scope=True)
        expr = expr.strip()
        if rows.n == 0:
            return {float(p): float("nan") for p in pcts}
        if expr in self.vtracks or self.gtrack_exists(expr):
            col = self._var_column(expr, rows)  # stays in HBM
        else:
            col = self.ctx.alloc(rows.n, np.float64).from_host(np.asarray(self._eval(expr, rows, {}), dtype=np.float64))
        q, _ = gpu.quantiles(self.ctx, col, rows.n, pcts)
        return {float(p): float(v) for p, v in zip(pcts, q)}

    def gintervals_quantiles(self, expr, percentiles=0.5, intervals=None, iterator=None):
        """R gintervals.quantiles -> gintervals_quantiles (src/GenomeTrackQuantiles.cpp:612-860, 1-D): per interval of the SORTED
        (not unified) intervals, the quantiles of the expression's non-NaN values over the iterator rows inside it -- a
        StreamPercentiler<double> fed the float-narrowed values, reset at every scope interval (:726-741), calc_medians' exact
        branch. Returns a dict: chrom / start / end of the sorted intervals and one float64 column per percentile (NaN where an
        interval holds no value). The value column stays in HBM; all intervals' slices go through one mg_quantiles_segmented call."""
        pcts = np.atleast_1d(np.asarray(percentiles, dtype=np.float64))
        for p in pcts:
            if not (0 <= p <= 1):
                raise MishaError("Percentile (%g) is not in [0, 1] range\n" % p)
        intervals = self.gintervals_all() if intervals is None else intervals
        rows, order = self._rows([expr], intervals, iterator, unify_scope=False)
        expr = expr.strip()
        n_int = len(order)
        out = {"chrom": np.asarray(intervals["chrom"])[order], "start": np.asarray(intervals["start"])[order],
               "end": np.asarray(intervals["end"])[order]}
        cols = np.full((len(pcts), n_int), np.nan)
        if rows.n:
            if expr in self.vtracks or self.gtrack_exists(expr):
                col = self._var_column(expr, rows)
            else:
                col = self.ctx.alloc(rows.n, np.float64).from_host(np.asarray(self._eval(expr, rows, {}), dtype=np.float64))
            scope = rows.coords()[3]  # index of each row's interval in the sorted scope; rows of one interval are contiguous
            cuts = np.flatnonzero(np.diff(scope)) + 1
            firsts = np.concatenate([[0], cuts])
            q, _ = gpu.quantiles_segmented(self.ctx, col, np.concatenate([firsts, [rows.n]]), pcts)  # NaN where nothing was kept
            cols[:, scope[firsts]] = q
        for k, p in enumerate(pcts):
            out[float(p)] = cols[k]
        return out

    def gsummary(self, expr, intervals=None, iterator=None):
        """R/compute-core.R:286. Returns the named 7-vector as a dict."""
        intervals = self.gintervals_all() if intervals is None else intervals
        rows, _ = self._rows([expr], intervals, iterator, unify_scope=True)
        part = gpu.summary_init(self.ctx)
        expr = expr.strip()
        vt = self.vtracks.get(expr)
        bare_dense = (vt is None and self.gtrack_exists(expr) and self._track_type(expr) == "dense") or \
                     (vt is not None and isinstance(vt.src, str) and self._track_type(vt.src) == "dense" and vt.func in FUSED_FUNCS)
        if rows.n:
            if bare_dense:  # fused on the device: values never leave the chip
                if vt is None:
                    self._dense_track(expr).summary_dev(rows, "avg", part)
                else:
                    f = _gpu_func(vt)
                    self._dense_track(vt.src).summary_dev(rows, f, part, vt.sshift, vt.eshift)
            else:
                v = np.asarray(self._eval(expr, rows, {}), dtype=np.float64)
                d = self.ctx.alloc(len(v), np.float64).from_host(v)
                gpu.summary_add(self.ctx, d, len(v), part)
        out = gpu.summary_finalize(part.to_host())
        return dict(zip(gpu.SUMMARY_NAMES, out))

    def gdist(self, *args, intervals=None, iterator=None, include_lowest=False):
        """R/compute-distribution.R:152: gdist(expr1, breaks1, [expr2, breaks2, ...]). Returns counts, one axis per expression."""
        if len(args) < 2 or len(args) % 2:
            raise MishaError("Usage: gdist([expr, breaks]+, intervals = ALLGENOME, include.lowest = FALSE, iterator = NULL)")
        exprs, breaks = [a.strip() for a in args[0::2]], [np.asarray(b, dtype=np.float64) for b in args[1::2]]
        intervals = self.gintervals_all() if intervals is None else intervals
        rows, _ = self._rows(exprs, intervals, iterator, unify_scope=True)
        shape = tuple(len(b) - 1 for b in breaks)
        counts = self.ctx.alloc(int(np.prod(shape)), np.uint64).zero()
        if rows.n:
            e0 = exprs[0]
            if len(exprs) == 1 and e0 not in self.vtracks and self.gtrack_exists(e0) and self._track_type(e0) == "dense":
                self._dense_track(e0).hist_dev(rows, "avg", breaks[0], counts, include_lowest)
            else:
                cache = {}
                cols = [self.ctx.alloc(rows.n, np.float64).from_host(np.asarray(self._eval(e, rows, cache), dtype=np.float64)) for e in exprs]
                gpu.hist_dev(self.ctx, cols, rows.n, breaks, counts, include_lowest)
        return counts.to_host().astype(np.float64).reshape(shape, order="F")

    def _column(self, expr, rows, cache):
        """The expression's value column in HBM: a vtrack / track name is computed there, anything else is evaluated on the host."""
        expr = expr.strip()
        if expr in self.vtracks or self.gtrack_exists(expr):
            return self._var_column(expr, rows)
        return self.ctx.alloc(rows.n, np.float64).from_host(np.asarray(self._eval(expr, rows, cache), dtype=np.float64))

    def gintervals_summary(self, expr, intervals=None, iterator=None):
        """R gintervals.summary -> gintervals_summary (src/GenomeTrackSummary.cpp:248-431, 1-D): the gsummary row of every interval of
        the SORTED intervals (an interval without iterator rows: 0 rows, NaN statistics). Returns chrom / start / end and the seven
        summary columns. One mg_summary_segmented call over the value column in HBM."""
        intervals = self.gintervals_all() if intervals is None else intervals
        rows, order = self._rows([expr], intervals, iterator, unify_scope=False)
        out = {"chrom": np.asarray(intervals["chrom"])[order], "start": np.asarray(intervals["start"])[order],
               "end": np.asarray(intervals["end"])[order]}
        res = np.tile(gpu.summary_finalize(np.array([0, 0, 0, 1.7976931348623157e308, -1.7976931348623157e308, 0.0])), (len(order), 1))
        if rows.n:
            col = self._column(expr, rows, {})
            scope = rows.coords()[3]
            firsts = np.concatenate([[0], np.flatnonzero(np.diff(scope)) + 1])
            res[scope[firsts]] = gpu.summary_segmented(self.ctx, col, np.concatenate([firsts, [rows.n]]))
        for k, name in enumerate(gpu.SUMMARY_NAMES):
            out[name] = res[:, k]
        return out

    def _gbins_args(self, args, usage):
        if len(args) < 2 or len(args) % 2:
            raise MishaError(usage)
        return [a.strip() for a in args[0::2]], [np.asarray(b, dtype=np.float64) for b in args[1::2]]

    def gbins_summary(self, *args, expr=None, intervals=None, iterator=None, include_lowest=False):
        """R gbins.summary(bin_expr1, breaks1, ..., expr=) -> gbins_summary (src/GenomeTrackSummary.cpp:433-506): the gsummary row of
        `expr` over the rows of every N-D bin of the bin expressions. Returns an array of shape (bins of dim 1, ..., 7)."""
        bexprs, breaks = self._gbins_args(args, "Usage: gbins.summary([expr, breaks]+, expr, intervals = ALLGENOME, include.lowest = FALSE, iterator = NULL)")
        intervals = self.gintervals_all() if intervals is None else intervals
        rows, _ = self._rows([expr] + bexprs, intervals, iterator, unify_scope=True)
        shape = tuple(len(b) - 1 for b in breaks)
        cache = {}
        n = rows.n
        col = self._column(expr, rows, cache) if n else self.ctx.alloc(1, np.float64)
        bcols = [self._column(e, rows, cache) if n else col for e in bexprs]
        res = gpu.bins_summary(self.ctx, col, bcols, n, breaks, include_lowest)
        return res.reshape(shape + (7,), order="F")

    def gbins_quantiles(self, *args, expr=None, percentiles=0.5, intervals=None, iterator=None, include_lowest=False):
        """R gbins.quantiles -> gbins_quantiles (src/GenomeTrackQuantiles.cpp:1199-1330): the quantiles of `expr` over the rows of every
        N-D bin of the bin expressions (NaN for a bin without values). Returns an array of shape (bins of dim 1, ..., percentiles)."""
        bexprs, breaks = self._gbins_args(args, "Usage: gbins.quantiles([expr, breaks]+, expr, percentiles = 0.5, intervals = ALLGENOME, include.lowest = FALSE, iterator = NULL)")
        pcts = np.atleast_1d(np.asarray(percentiles, dtype=np.float64))
        for p in pcts:
            if not (0 <= p <= 1):
                raise MishaError("Percentile (%g) is not in [0, 1] range\n" % p)
        intervals = self.gintervals_all() if intervals is None else intervals
        rows, _ = self._rows([expr] + bexprs, intervals, iterator, unify_scope=True)
        shape = tuple(len(b) - 1 for b in breaks)
        cache = {}
        n = rows.n
        col = self._column(expr, rows, cache) if n else self.ctx.alloc(1, np.float64)
        bcols = [self._column(e, rows, cache) if n else col for e in bexprs]
        q, _ = gpu.bins_quantiles(self.ctx, col, bcols, n, breaks, pcts, include_lowest)
        return q.T.reshape(shape + (len(pcts),), order="F")


def gdb_init(groot, device=0):
    return MishaDB(groot, device)
