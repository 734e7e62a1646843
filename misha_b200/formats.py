"""misha's on-disk formats on the host side of staging: per-chromosome track files and the single-file INDEXED layouts.

  per-chromosome   <track>.track/<chrom>         dense : uint32 bin_size, float32[nbins]          (src/GenomeTrackFixedBin.cpp:1216-1260)
                                                  sparse: int32 -1, {int64 start, int64 end, float32 val}[n]   (src/GenomeTrackSparse.cpp:162-235)
                   seq/<chrom>.seq                ASCII bases                                          (src/GenomeSeqFetch.cpp:109-151)
  indexed          <track>.track/track.idx + track.dat   "MISHATDX" index (src/TrackIndex.h:7-22, src/TrackIndex.cpp:20-200): header
                                                  {magic[8], uint32 version = 1, uint32 type (0 dense, 1 sparse, 2 array), uint32 ncontigs,
                                                  uint64 flags (bit 0: little endian), uint64 CRC64-ECMA of the entries}, then per contig
                                                  {uint32 chromid, uint64 offset, uint64 length, uint32 reserved}; track.dat holds, at
                                                  [offset, offset + length), exactly the bytes of the per-chromosome file
                                                  (src/GenomeTrackFixedBin.cpp:1161-1212: bin size first, no other signature)
                   seq/genome.idx + genome.seq    "MISHAIDX" index (src/GenomeIndex.cpp:20-205): {magic[8], uint32 version = 1, uint32 ncontigs,
                                                  uint64 checksum}, then per contig {uint32 chromid, uint16 name length, name, uint64 offset,
                                                  uint64 length (bases), uint64 reserved}; genome.seq is the contigs' bases back to back
                                                  (src/GenomeSeqFetch.cpp:66-107)

Readers return numpy arrays ready for mg_*_stage; writers exist for the synthetic databases of the tests and the bench. Checksums are
verified on load and a mismatch is an error, as in the reference. numpy only: this is host IO, not the compute path.
"""
import os
import struct

import numpy as np

TRACK_MAGIC = b"MISHATDX"
GENOME_MAGIC = b"MISHAIDX"
TRACK_TYPES = {"dense": 0, "sparse": 1, "array": 2}
SPARSE_REC = np.dtype([("s", "<i8"), ("e", "<i8"), ("v", "<f4")])


class FormatError(RuntimeError):
    pass


# ---- CRC64-ECMA, reflected (src/CRC64.h:17-72) -------------------------------------------------------------------------------
_CRC_POLY = 0xC96C5795D7870F42
_crc_table = None


def _table():
    global _crc_table
    if _crc_table is None:
        t = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ _CRC_POLY if c & 1 else c >> 1
            t.append(c)
        _crc_table = t
    return _crc_table


def crc64(chunks):
    """CRC64 of the concatenation of byte strings (init ~0, final complement)."""
    t = _table()
    c = 0xFFFFFFFFFFFFFFFF
    for b in chunks:
        for byte in b:
            c = (c >> 8) ^ t[(c ^ byte) & 0xFF]
    return c ^ 0xFFFFFFFFFFFFFFFF


# ---- track.idx / track.dat -----------------------------------------------------------------------------------------------------
def read_track_index(track_dir):
    """{chromid: (offset, length)} and the track type name, or None when the track is stored per chromosome."""
    path = os.path.join(track_dir, "track.idx")
    if not os.path.exists(path):
        return None
    raw = open(path, "rb").read()
    if len(raw) < 36 or raw[:8] != TRACK_MAGIC:
        raise FormatError("Invalid index file header in %s (expected MISHATDX magic)" % path)
    version, ttype, n, flags, checksum = struct.unpack_from("<IIIQQ", raw, 8)
    if version != 1:
        raise FormatError("Index version %d not supported (expected 1) in %s" % (version, path))
    if ttype > 2:
        raise FormatError("Invalid track type %d in %s (expected 0-2)" % (ttype, path))
    if not flags & 1:
        raise FormatError("Index file %s has incompatible endianness" % path)
    if len(raw) < 36 + 24 * n:
        raise FormatError("Failed to read entry table of %s" % path)
    entries, chunks = {}, []
    for i in range(n):
        chromid, off, ln, _ = struct.unpack_from("<IQQI", raw, 36 + 24 * i)
        if chromid in entries:
            raise FormatError("Duplicate chromosome ID %d in track index %s" % (chromid, path))
        entries[chromid] = (off, ln)
        chunks += [struct.pack("<I", chromid), struct.pack("<Q", off), struct.pack("<Q", ln)]  # reserved is not hashed
    if crc64(chunks) != checksum:
        raise FormatError("Index file checksum mismatch in %s. Index may be corrupt." % path)
    return entries, [k for k, v in TRACK_TYPES.items() if v == ttype][0]


def write_indexed_track(track_dir, kind, blobs):
    """blobs: {chromid: bytes of the per-chromosome file, or b"" for an empty contig}; every chromosome of the genome gets an entry."""
    os.makedirs(track_dir, exist_ok=True)
    off, table, chunks = 0, [], []
    with open(os.path.join(track_dir, "track.dat"), "wb") as f:
        for chromid in sorted(blobs):
            b = blobs[chromid]
            f.write(b)
            table.append(struct.pack("<IQQI", chromid, off, len(b), 0))
            chunks += [struct.pack("<I", chromid), struct.pack("<Q", off), struct.pack("<Q", len(b))]
            off += len(b)
    with open(os.path.join(track_dir, "track.idx"), "wb") as f:
        f.write(TRACK_MAGIC + struct.pack("<IIIQQ", 1, TRACK_TYPES[kind], len(table), 1, crc64(chunks)))
        f.write(b"".join(table))


def _chrom_blob(track_dir, chrom_name, chromid, index):
    """The bytes of one chromosome's data, from track.dat (indexed) or the per-chromosome file; None when there is none."""
    if index is not None:
        ent = index[0].get(chromid)
        if ent is None or ent[1] == 0:
            return None
        return np.fromfile(os.path.join(track_dir, "track.dat"), dtype=np.uint8, count=ent[1], offset=ent[0])
    f = os.path.join(track_dir, chrom_name)
    if not os.path.exists(f):
        return None
    return np.fromfile(f, dtype=np.uint8)


def track_type(track_dir, chrom_names):
    """"dense" / "sparse" (src/GenomeTrack.cpp:227-250: the first int32 of a chromosome's data; the index says it for indexed tracks)."""
    index = read_track_index(track_dir)
    if index is not None:
        return index[1]
    for c in chrom_names:
        f = os.path.join(track_dir, c)
        if os.path.exists(f) and os.path.getsize(f) >= 4:
            sig = int(np.fromfile(f, dtype=np.int32, count=1)[0])
            if sig >= 1:
                return "dense"
            if sig == -1:
                return "sparse"
            if sig == ARRAY_SIGNATURE:
                return "array"
            raise FormatError("unsupported track type (signature %d)" % sig)
    return None


def read_dense_chrom(track_dir, chrom_name, chromid, index=None):
    """(bin_size, float32 bins) of one chromosome, or None."""
    raw = _chrom_blob(track_dir, chrom_name, chromid, index)
    if raw is None or len(raw) < 4:
        return None
    return int(raw[:4].view(np.uint32)[0]), raw[4:4 + (len(raw) - 4) // 4 * 4].view(np.float32)


def read_sparse_chrom(track_dir, chrom_name, chromid, index=None):
    """(start int64, end int64, val float32) of one chromosome (empty arrays when there is no data)."""
    raw = _chrom_blob(track_dir, chrom_name, chromid, index)
    if raw is None or len(raw) < 4:
        return np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.float32)
    n = (len(raw) - 4) // SPARSE_REC.itemsize
    rec = raw[4:4 + n * SPARSE_REC.itemsize].view(SPARSE_REC)
    return rec["s"].copy(), rec["e"].copy(), rec["v"].copy()


# ---- array tracks (src/GenomeTrackArrays.cpp) -----------------------------------------------------------------------------------
ARRAY_SIGNATURE = -8  # GenomeTrack::FORMAT_SIGNATURES[ARRAYS], src/GenomeTrack.cpp:23
SLICE_FUNCS = ("avg", "min", "max", "stddev", "sum", "quantile")  # GenomeTrackArrays::SLICE_FUNCTION_NAMES


def read_array_chrom(track_dir, chrom_name, chromid, index=None):
    """One chromosome of an array track: (start int64[n], end int64[n], rec_first int64[n + 1], vals float32[m], cols uint32[m]) -- every
    record's stored (value, column) pairs back to back (NaN values are never stored, src/GenomeTrackArrays.cpp:281-298). Layout: int32
    signature, int64 position of the interval table, the records' {uint32 count, count x {float32, uint32}} blocks, then the table
    {uint64 n, n x {int64 start, int64 end, int64 position of the record's block}}; positions are relative to the chromosome's data
    (:141-200)."""
    raw = _chrom_blob(track_dir, chrom_name, chromid, index)
    empty = (np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(1, np.int64), np.zeros(0, np.float32), np.zeros(0, np.uint32))
    if raw is None or len(raw) < 12:
        return empty
    if int(raw[:4].view(np.int32)[0]) != ARRAY_SIGNATURE:
        raise FormatError("Invalid arrays track header in %s" % os.path.join(track_dir, chrom_name))
    ipos = int(raw[4:12].view(np.int64)[0])
    n = int(raw[ipos:ipos + 8].view(np.uint64)[0])
    tab = raw[ipos + 8:ipos + 8 + 24 * n].view(np.int64).reshape(n, 3)
    start, end, vpos = tab[:, 0].copy(), tab[:, 1].copy(), tab[:, 2]
    if n and (np.any(start < 0) or np.any(start >= end) or np.any(start[1:] < end[:-1]) or np.any(vpos < 0) or np.any(vpos >= len(raw)) or np.any(np.diff(vpos) <= 0)):
        raise FormatError("Invalid format of arrays track file %s" % os.path.join(track_dir, chrom_name))
    counts = np.array([int(raw[p:p + 4].view(np.uint32)[0]) for p in vpos], dtype=np.int64)
    rec_first = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    vals, cols = np.empty(rec_first[-1], np.float32), np.empty(rec_first[-1], np.uint32)
    for r, p in enumerate(vpos):
        blk = raw[p + 4:p + 4 + 8 * counts[r]]
        pair = blk.view(np.dtype([("v", "<f4"), ("i", "<u4")]))
        vals[rec_first[r]:rec_first[r + 1]], cols[rec_first[r]:rec_first[r + 1]] = pair["v"], pair["i"]
    return start, end, rec_first, vals, cols


def array_slice_values(rec_first, vals, cols, slice_cols=None, func="avg", percentile=0.5):
    """The value every record of an array track presents under gvtrack.array.slice(slice, func): GenomeTrackArrays::get_sliced_val
    (src/GenomeTrackArrays.cpp:372-556), narrowed to float as its return type does. slice_cols None / empty = every stored value.
    Accumulations run in double in column order; squares of the stddev are float products as in the reference (v * v on floats)."""
    n = len(rec_first) - 1
    out = np.full(n, np.nan, dtype=np.float32)
    sl = None if slice_cols is None or len(slice_cols) == 0 else np.unique(np.asarray(slice_cols, dtype=np.uint32))
    for r in range(n):
        v, c = vals[rec_first[r]:rec_first[r + 1]], cols[rec_first[r]:rec_first[r + 1]]
        if sl is not None:
            keep = np.isin(c, sl)
            order = np.argsort(c[keep], kind="stable")
            v = v[keep][order]
            v = v[~np.isnan(v)]
            if func in ("avg", "sum", "min", "max", "quantile") and len(v) == 0:
                continue
        N = len(v)
        with np.errstate(all="ignore"):
            if func == "avg":
                out[r] = np.float32(_seq_sum(v) / N) if N else np.float32(np.nan)
            elif func == "sum":
                out[r] = np.float32(_seq_sum(v))
            elif func == "min":
                out[r] = v.min() if N else np.float32(np.finfo(np.float32).max)
            elif func == "max":
                out[r] = v.max() if N else np.float32(-np.finfo(np.float32).max)
            elif func == "stddev":
                if sl is not None and N <= 1:
                    continue
                s, msq = _seq_sum(v), _seq_sum((v * v).astype(np.float32))
                avg = s / N if N else np.nan
                out[r] = np.float32(np.sqrt(msq / (N - 1) - (avg * avg) * (N / (N - 1)))) if N != 1 else np.float32(np.nan)
            elif func == "quantile":
                srt = np.sort(v)
                idx = (N - 1) * min(max(percentile, 0.0), 1.0)  # StreamPercentiler exact regime, src/StreamPercentiler.h:191-217
                i1, i2 = int(np.floor(idx)), int(np.ceil(idx))
                w = idx - i1
                out[r] = np.float32(float(srt[i1]) * (1.0 - w) + float(srt[i2]) * w)
            else:
                raise FormatError("Unrecognized slice function %s" % func)
    return out


def _seq_sum(v):
    s = 0.0
    for x in v:
        s += float(x)
    return s


def array_blob(start, end, rec_first, vals, cols):
    """The bytes of one array-track chromosome (tests / synthetic databases); NaN values are dropped, records without values skipped."""
    body, table = [], []
    pos = 12
    for r in range(len(start)):
        v, c = np.asarray(vals[rec_first[r]:rec_first[r + 1]], np.float32), np.asarray(cols[rec_first[r]:rec_first[r + 1]], np.uint32)
        keep = ~np.isnan(v)
        if not keep.any():
            continue
        pair = np.empty(int(keep.sum()), dtype=np.dtype([("v", "<f4"), ("i", "<u4")]))
        pair["v"], pair["i"] = v[keep], c[keep]
        blk = np.uint32(len(pair)).tobytes() + pair.tobytes()
        table.append((int(start[r]), int(end[r]), pos))
        body.append(blk)
        pos += len(blk)
    tab = np.array(table, dtype=np.int64).reshape(-1, 3)
    return np.int32(ARRAY_SIGNATURE).tobytes() + np.int64(pos).tobytes() + b"".join(body) + np.uint64(len(table)).tobytes() + tab.tobytes()


def dense_blob(bins, bin_size):
    return np.uint32(bin_size).tobytes() + np.ascontiguousarray(bins, dtype=np.float32).tobytes()


def sparse_blob(start, end, val):
    rec = np.empty(len(start), dtype=SPARSE_REC)
    rec["s"], rec["e"], rec["v"] = start, end, val
    return np.int32(-1).tobytes() + rec.tobytes()


# ---- genome.idx / genome.seq -----------------------------------------------------------------------------------------------------
def read_genome_index(seq_dir):
    """{chromid: (name, offset, length)} or None when the sequence is stored per chromosome."""
    path = os.path.join(seq_dir, "genome.idx")
    if not os.path.exists(path):
        return None
    raw = open(path, "rb").read()
    if len(raw) < 24 or raw[:8] != GENOME_MAGIC:
        raise FormatError("Invalid index file header in %s" % path)
    version, n, checksum = struct.unpack_from("<IIQ", raw, 8)
    if version != 1:
        raise FormatError("Index version %d not supported (expected 1) in %s" % (version, path))
    pos, entries, chunks = 24, {}, []
    for i in range(n):
        if pos + 6 > len(raw):
            raise FormatError("Failed to read chromid at entry %d in %s" % (i, path))
        chromid, nl = struct.unpack_from("<IH", raw, pos)
        pos += 6
        if nl > 1024:
            raise FormatError("Contig name length %d exceeds maximum (1024) at entry %d in %s" % (nl, i, path))
        name = raw[pos:pos + nl]
        pos += nl
        if pos + 24 > len(raw):
            raise FormatError("Failed to read offset/length/reserved at entry %d in %s" % (i, path))
        off, ln, _ = struct.unpack_from("<QQQ", raw, pos)
        pos += 24
        if chromid in entries:
            raise FormatError("Duplicate chromosome ID %d (contig '%s') in index %s" % (chromid, name.decode(), path))
        entries[chromid] = (name.decode(), off, ln)
        chunks += [struct.pack("<I", chromid), name, struct.pack("<Q", off), struct.pack("<Q", ln)]
    if crc64(chunks) != checksum:
        raise FormatError("Index file checksum mismatch in %s. Index may be corrupt." % path)
    return entries


def write_indexed_genome(seq_dir, contigs):
    """contigs: [(chromid, name, uint8 bases)] -> genome.seq + genome.idx."""
    os.makedirs(seq_dir, exist_ok=True)
    off, body, chunks = 0, [], []
    with open(os.path.join(seq_dir, "genome.seq"), "wb") as f:
        for chromid, name, bases in contigs:
            b = np.frombuffer(bases, dtype=np.uint8) if isinstance(bases, (bytes, bytearray)) else np.ascontiguousarray(bases, dtype=np.uint8)
            f.write(b.tobytes())
            nm = name.encode()
            body.append(struct.pack("<IH", chromid, len(nm)) + nm + struct.pack("<QQQ", off, len(b), 0))
            chunks += [struct.pack("<I", chromid), nm, struct.pack("<Q", off), struct.pack("<Q", len(b))]
            off += len(b)
    with open(os.path.join(seq_dir, "genome.idx"), "wb") as f:
        f.write(GENOME_MAGIC + struct.pack("<IIQ", 1, len(contigs), crc64(chunks)))
        f.write(b"".join(body))


def read_seq_chrom(seq_dir, chrom_name, chromid, index=None):
    """uint8 bases of one chromosome from genome.seq (indexed) or <chrom>.seq / chr<chrom>.seq (src/GenomeSeqFetch.cpp:112-130)."""
    if index is not None:
        ent = index.get(chromid)
        if ent is None:
            raise FormatError("Contig ID %d not found in genome index" % chromid)
        return np.fromfile(os.path.join(seq_dir, "genome.seq"), dtype=np.uint8, count=ent[2], offset=ent[1])
    for cand in (chrom_name + ".seq", "chr" + chrom_name + ".seq"):
        f = os.path.join(seq_dir, cand)
        if os.path.exists(f):
            return np.fromfile(f, dtype=np.uint8)
    raise FormatError("Failed to open sequence file for chromosome %s in %s" % (chrom_name, seq_dir))
