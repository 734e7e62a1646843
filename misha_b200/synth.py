"""Seeded synthetic hg38-scale inputs in misha's exact on-disk formats.

Used by bench.py and the full-size property tests: there is no network for real tracks. Everything is generated with
numpy from one seed, chromosome by chromosome, so a rank can generate only the chromosomes it owns.
"""
import os

import numpy as np

SEED = 20261019

HG38 = {
    "chr1": 248956422, "chr2": 242193529, "chr3": 198295559, "chr4": 190214555, "chr5": 181538259, "chr6": 170805979,
    "chr7": 159345973, "chr8": 145138636, "chr9": 138394717, "chr10": 133797422, "chr11": 135086622, "chr12": 133275309,
    "chr13": 114364328, "chr14": 107043718, "chr15": 101991189, "chr16": 90338345, "chr17": 83257441, "chr18": 80373285,
    "chr19": 58617616, "chr20": 64444167, "chr21": 46709983, "chr22": 50818468, "chrX": 156040895, "chrY": 57227415,
}


def genome(scale=1.0):
    """[(name, size)] in chromid order = alphabetical names, as a per-chromosome DB orders them (R/db-root.R:157-162)."""
    names = sorted(HG38)
    return [(n, max(int(HG38[n] * scale), 1000)) for n in names]


def _rng(kind, chromid, seed=SEED):
    return np.random.default_rng([seed, {"dense": 1, "intervals": 2, "sparse": 3, "seq": 4, "pssm": 5}[kind], chromid])


def dense_bins(chromid, size, bin_size=20, seed=SEED):
    """float32[ceil(size/bin)]: exp(N(0,1)) clipped to [0,100], 5% NaN in runs (mean 50 bins), 0.01% +inf."""
    rng = _rng("dense", chromid, seed)
    n = -(-size // bin_size)
    v = rng.standard_normal(n, dtype=np.float32)
    np.exp(v, out=v)
    np.clip(v, 0, 100, out=v)
    starts = np.flatnonzero(rng.random(n, dtype=np.float32) < 0.001)
    lens = rng.geometric(1.0 / 50, size=len(starts))
    d = np.zeros(n + 1, dtype=np.int32)
    np.add.at(d, starts, 1)
    np.add.at(d, np.minimum(starts + lens, n), -1)
    v[np.cumsum(d[:n]) > 0] = np.nan
    v[rng.random(n, dtype=np.float32) < 1e-4] = np.inf
    return v


def intervals(chromid, size, n, lo=100, hi=300, seed=SEED):
    """n sorted, non-overlapping intervals with >= 1 bp gaps; lengths uniform{lo..hi}. Returns (start, end) int64."""
    rng = _rng("intervals", chromid, seed)
    if n == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    ln = rng.integers(lo, hi + 1, n, dtype=np.int64)
    free = size - int(ln.sum()) - n
    if free < 0:
        raise ValueError("chromosome too small for %d intervals" % n)
    g = rng.random(n + 1)
    gaps = np.floor(g / g.sum() * free).astype(np.int64)
    start = np.cumsum(gaps[:n] + 1) - 1 + np.concatenate([[0], np.cumsum(ln)[:-1]])
    return start, start + ln


def interval_counts(chroms, total):
    """Intervals per chromosome proportional to length, summing to `total` exactly."""
    sizes = np.array([s for _, s in chroms], dtype=np.float64)
    cnt = np.floor(sizes / sizes.sum() * total).astype(np.int64)
    cnt[np.argmax(sizes)] += total - cnt.sum()
    return cnt


def sparse_records(chromid, size, n, seed=SEED):
    """n sorted disjoint records, length uniform{20..500}, values U(0,1) float32 with 1% NaN."""
    rng = _rng("sparse", chromid, seed)
    start, end = intervals(chromid, size, n, 20, 500, seed + 1)
    val = rng.random(n, dtype=np.float32)
    val[rng.random(n, dtype=np.float32) < 0.01] = np.nan
    return start, end, val


def sequence(chromid, size, seed=SEED):
    """ASCII bytes: i.i.d. ACGT with GC = 0.41, 2% lowercase, 0.5% N in runs."""
    rng = _rng("seq", chromid, seed)
    u = rng.random(size, dtype=np.float32)
    out = np.empty(size, dtype=np.uint8)
    out[:] = ord("A")
    out[u >= 0.295] = ord("C")
    out[u >= 0.5] = ord("G")
    out[u >= 0.705] = ord("T")
    low = rng.random(size, dtype=np.float32) < 0.02
    out[low] |= 0x20
    starts = np.flatnonzero(rng.random(size, dtype=np.float32) < 0.005 / 40)
    lens = rng.geometric(1.0 / 40, size=len(starts))
    d = np.zeros(size + 1, dtype=np.int32)
    np.add.at(d, starts, 1)
    np.add.at(d, np.minimum(starts + lens, size), -1)
    out[np.cumsum(d[:size]) > 0] = ord("N")
    return out


def pssm(L=20, seed=SEED):
    return _rng("pssm", 0, seed).dirichlet([0.5] * 4, size=L)


# ---- on-disk writers (formats: SURVEY.md Appendix B) --------------------------------------------------------------
def write_chrom_sizes(groot, chroms):
    os.makedirs(groot, exist_ok=True)
    with open(os.path.join(groot, "chrom_sizes.txt"), "w") as f:
        for name, size in chroms:
            f.write("%s\t%d\n" % (name[3:] if name.startswith("chr") else name, size))


def track_dir(groot, track):
    return os.path.join(groot, "tracks", track.replace(".", "/") + ".track")


def write_dense_chrom(groot, track, chrom_name, bins, bin_size):
    d = track_dir(groot, track)
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, chrom_name), "wb") as f:
        f.write(np.uint32(bin_size).tobytes())
        f.write(np.ascontiguousarray(bins, dtype=np.float32).tobytes())


def write_sparse_chrom(groot, track, chrom_name, start, end, val):
    d = track_dir(groot, track)
    os.makedirs(d, exist_ok=True)
    rec = np.empty(len(start), dtype=np.dtype([("s", "<i8"), ("e", "<i8"), ("v", "<f4")]))
    rec["s"], rec["e"], rec["v"] = start, end, val
    with open(os.path.join(d, chrom_name), "wb") as f:
        f.write(np.int32(-1).tobytes())
        f.write(rec.tobytes())


def write_seq_chrom(groot, chrom_name, ascii_bytes):
    d = os.path.join(groot, "seq")
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, chrom_name + ".seq"), "wb") as f:
        f.write(np.ascontiguousarray(ascii_bytes, dtype=np.uint8).tobytes())


def lpt_shards(chroms, nranks):
    """Greedy longest-processing-time assignment of chromosomes to ranks (whole chromosomes per GPU)."""
    order = sorted(range(len(chroms)), key=lambda i: -chroms[i][1])
    load = [0] * nranks
    owner = [0] * len(chroms)
    for i in order:
        r = min(range(nranks), key=lambda k: load[k])
        owner[i] = r
        load[r] += chroms[i][1]
    return owner
