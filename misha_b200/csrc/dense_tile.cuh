// dense_tile.cuh -- k_dense_tile: avg / sum / nearest / size / exists, min / max and quantiles of dense windows, one THREAD BLOCK
// per tile of MG_TL_THREADS consecutive rows, everything computed out of shared memory.
//
// The rows the scanner produces are sorted, so the windows of a tile overlap almost entirely (C3: 256 rows 15 bins apart with
// 510-bin windows span ~4500 bins). Instead of every row gathering its own sectors of bins, prefix entries, sparse-table entries
// and wavelet levels from L1 / L2 (k_dense_rowquery: 2.1x the algorithmic DRAM bytes, 1310 instructions per row, bound by the
// latency of dependent global loads), the thread block
//   1. reduces its rows' bin range [b0, b1) and block-record range (warp redux + one barrier),
//   2. pulls the bins of that range and the wavelet levels of the block records with bulk copies (cp.async.bulk, one mbarrier),
//   3. builds, cooperatively and in shared memory, what the rows need:
//        - per chunk of 20 bins (one thread each): sum, non-NaN mask, max / min, sum |v|; the chunk-local prefix before each
//          quad of 4 bins; suffix / prefix extrema of the quads inside their chunk;
//        - a block scan of the chunk sums and counts (the tile-local prefix: no chromosome-wide prefix array is read, and its
//          rounding error is relative to the TILE's magnitude, not the chromosome's);
//        - a sparse table over the chunk extrema;
//        - the wavelet levels re-laid as {ones before, 32 bits} units, one 8-byte shared-memory load per rank query;
//   4. answers every row from those: sum = P(e) - P(s) with P(i) = chunk base + quad prefix + <= 3 bins; extrema = <= 3 + <= 3 bins,
//      two quad entries, two sparse-table entries; quantile = the 11-level walk on shared memory, the value read from the staged bins.
// DRAM traffic: the track once, the levels once (2.75 B/bin), rows in, columns out, one pos_of_rank sector per row.
//
// avg / sum are bit-identical to the reference's sequential double sum (src/GenomeTrackFixedBin.cpp:881-970): the tile knows
// the sum of |v| over its bins, hence a rigorous bound on |S_tile - S_sequential|; a row whose double result lies closer than
// that bound to a float rounding boundary is re-summed in bin order (a few rows per 100 000).
//
// A tile that does not fit (rows on two chromosomes, rows too far apart, windows longer than a block record serves) falls back to
// the per-row index path (mg_rq_row, the arithmetic of k_dense_rowquery without staging): correctness never depends on the order
// or density of the rows.
#pragma once
#include "dense_index.cuh"

#ifndef MG_TL_THREADS
#define MG_TL_THREADS 256
#endif
#define MG_TL_CHUNK 20                               // bins per thread in the cooperative pass: 5 float4, stride 20 words = conflict-free LDS.128
#define MG_TL_QPC (MG_TL_CHUNK / 4)                  // quads per chunk
#define MG_TL_CAP (MG_TL_THREADS * MG_TL_CHUNK)      // bins a tile can hold
#define MG_TL_NQD (MG_TL_CAP / 4)
#define MG_TL_SLOTS ((MG_TL_CAP - 1) / MG_BW_STRIDE + 2) // block records a tile can need
#define MG_TL_TAB_LEVELS 9                           // sparse table over <= 256 chunk extrema: a run of 2^8 chunks is level 8
#define MG_TL_NW (MG_TL_THREADS / 32)

static_assert(MG_TL_CHUNK % 4 == 0 && MG_TL_CAP < 43690, "chunk index by multiply-shift needs i < 43690");
static_assert((1 << (MG_TL_TAB_LEVELS - 1)) >= MG_TL_THREADS, "sparse table levels");

// shared-memory layout (byte offsets; every region 16-byte aligned)
#define MG_TL_OFF_VALS 0
#define MG_TL_SZ_VALS ((MG_TL_CAP + 4) * 4)
#define MG_TL_OFF_PL4 (MG_TL_OFF_VALS + MG_TL_SZ_VALS)
#define MG_TL_SZ_PL4 (((MG_TL_NQD + 1) * 8 + 15) & ~15)
#define MG_TL_OFF_PB (MG_TL_OFF_PL4 + MG_TL_SZ_PL4)
#define MG_TL_SZ_PB (((MG_TL_THREADS + 1) * 8 + 15) & ~15)
#define MG_TL_OFF_CM (MG_TL_OFF_PB + MG_TL_SZ_PB)
#define MG_TL_OFF_EXT (MG_TL_OFF_CM + MG_TL_SZ_PB)   // per requested extremum: quad suffix, quad prefix, sparse table
#define MG_TL_SZ_Q1 (MG_TL_NQD * 4)
#define MG_TL_SZ_TAB (MG_TL_TAB_LEVELS * MG_TL_THREADS * 4)
#define MG_TL_SZ_EXT (2 * MG_TL_SZ_Q1 + MG_TL_SZ_TAB)

struct TileLayout {
	int off_lv, total;
};
// next: extrema requested (0..2); Q: the walked levels of the block records are staged
__host__ __device__ constexpr TileLayout mg_tile_layout(int next, bool Q)
{
	TileLayout L{};
	L.off_lv = MG_TL_OFF_EXT + next * MG_TL_SZ_EXT;
	L.total = L.off_lv + (Q ? MG_TL_SLOTS * MG_BW_STAGE_BYTES : 0);
	return L;
}

// ---- the per-row index path (fallback of a tile that does not fit): mg_rq_row of dense_index.cuh, kept out of line ----------
template <bool SUM, bool MM, bool Q>
__device__ __noinline__ void mg_tile_slow_row(const DenseQuery &q, int64_t i)
{
	DenseWin w;
	const bool ok = mg_dense_win(q, q.first + i, w);
	mg_rq_row<SUM, MM, Q, false>(q, i, w, ok);
}

// x when it is a number, else 0; bit (a constant) is OR-ed into mask when it is. Inline PTX: left to the compiler, the select moves
// behind the conversion to double and becomes two selects on the halves.
template <uint32_t BIT>
__device__ __forceinline__ float mg_clean_bit(float x, uint32_t &mask)
{
	float c;
	uint32_t b;
	asm("{\n\t.reg .pred p;\n\tsetp.num.f32 p, %2, %2;\n\tselp.f32 %0, %2, 0f00000000, p;\n\tselp.u32 %1, %3, 0, p;\n\t}" : "=f"(c), "=r"(b) : "f"(x), "n"(BIT));
	mask |= b;
	return c;
}
// x when it is a number and k > IDX, else 0
template <uint32_t IDX>
__device__ __forceinline__ float mg_clean_below(float x, uint32_t k)
{
	float c;
	asm("{\n\t.reg .pred p, q;\n\tsetp.gt.u32 q, %2, %3;\n\tsetp.num.and.f32 p, %1, %1, q;\n\tselp.f32 %0, %1, 0f00000000, p;\n\t}" : "=f"(c) : "f"(x), "r"(k), "n"(IDX));
	return c;
}

// ---- tile helpers ------------------------------------------------------------------------------------------------------
struct TileSmem {
	const float *vals;  // [CAP + 4] the tile's bins, NaN beyond the staged range
	const double *pl4;  // [NQD + 1] chunk-local prefix sum before each quad
	const double *pb;   // [THREADS + 1] tile prefix sum before each chunk
	const uint2 *cm;    // [THREADS + 1] {non-NaN bins before the chunk, non-NaN mask of the chunk's 20 bins}
};
struct TileExt {
	const float *qsuf; // [NQD] extremum of the quads from this one to the end of its chunk
	const float *qpre; // [NQD] ... from the start of its chunk to this one
	const float *tab;  // [TAB_LEVELS][THREADS] sparse table over the chunk extrema
};

__device__ __forceinline__ uint32_t mg_tl_chunk_of(uint32_t i) { return (i * 52429u) >> 20; } // i / 20 for i < 43690

// P(i) = sum of the non-NaN bins [0, i) of the tile, C(i) = their count; v = the quad holding bin i (reused by the extrema)
template <bool SUM>
__device__ __forceinline__ void mg_tl_prefix(const TileSmem &s, uint32_t i, double &P, uint32_t &C, float4 &v)
{
	const uint32_t c = mg_tl_chunk_of(i), j = i - c * MG_TL_CHUNK, qd = i >> 2, k = i & 3u;
	const uint2 m = s.cm[c];
	C = m.x + __popc(m.y & ((1u << j) - 1u));
	v = reinterpret_cast<const float4 *>(s.vals)[qd];
	if (SUM) {
		// the <= 3 bins of the quad below i, NaN as 0: float selects, then exact conversions
		const float x0 = mg_clean_below<0>(v.x, k), x1 = mg_clean_below<1>(v.y, k), x2 = mg_clean_below<2>(v.z, k);
		P = ((s.pb[c] + s.pl4[qd]) + (double)x0) + ((double)x1 + (double)x2);
	}
}

template <bool IS_MAX>
__device__ __forceinline__ float mg_tl_ext2(float a, float b)
{
	return IS_MAX ? fmaxf(a, b) : fminf(a, b); // NaN operands are skipped
}

// extremum of the tile's bins [s, e), s < e; vs / ve = the quads holding bins s and e (from mg_tl_prefix). -inf / +inf when the
// window holds no value.
template <bool IS_MAX>
__device__ __forceinline__ float mg_tl_extremum(const TileSmem &sm, const TileExt &x, uint32_t s, uint32_t e, const float4 &vs, const float4 &ve)
{
	float acc = IS_MAX ? __int_as_float(0xff800000) : __int_as_float(0x7f800000);
	const uint32_t qs = s >> 2, qe = e >> 2, ks = s & 3u, ke = e & 3u;
	const float hs[4] = {vs.x, vs.y, vs.z, vs.w}, he[4] = {ve.x, ve.y, ve.z, ve.w};
	if (qs == qe) { // inside one quad
#pragma unroll
		for (int k = 0; k < 4; ++k)
			if ((uint32_t)k >= ks && (uint32_t)k < ke)
				acc = mg_tl_ext2<IS_MAX>(acc, hs[k]);
		return acc;
	}
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		if ((uint32_t)k >= ks)
			acc = mg_tl_ext2<IS_MAX>(acc, hs[k]);
		if ((uint32_t)k < ke)
			acc = mg_tl_ext2<IS_MAX>(acc, he[k]);
	}
	const uint32_t m0 = qs + 1, m1 = qe; // whole quads [m0, m1)
	if (m0 >= m1)
		return acc;
	const uint32_t c0 = mg_tl_chunk_of(4 * m0 + MG_TL_CHUNK - 4), c1 = mg_tl_chunk_of(4 * m1); // whole chunks [c0, c1)
	if (c0 > c1) { // no chunk boundary inside: <= 4 quads of one chunk
		for (uint32_t qd = m0; qd < m1; ++qd) {
			const float4 v = reinterpret_cast<const float4 *>(sm.vals)[qd];
			acc = mg_tl_ext2<IS_MAX>(acc, mg_tl_ext2<IS_MAX>(mg_tl_ext2<IS_MAX>(v.x, v.y), mg_tl_ext2<IS_MAX>(v.z, v.w)));
		}
		return acc;
	}
	if (m0 < c0 * MG_TL_QPC) // quads from m0 to the end of their chunk
		acc = mg_tl_ext2<IS_MAX>(acc, x.qsuf[m0]);
	if (m1 > c1 * MG_TL_QPC) // quads from the start of the last chunk up to m1 - 1
		acc = mg_tl_ext2<IS_MAX>(acc, x.qpre[m1 - 1]);
	if (c0 < c1) {
		const uint32_t len = c1 - c0;
		const int j = 31 - __clz((int)len);
		const float *t = x.tab + j * MG_TL_THREADS;
		acc = mg_tl_ext2<IS_MAX>(acc, mg_tl_ext2<IS_MAX>(t[c0], t[c1 - (1u << j)]));
	}
	return acc;
}

// double S rounds to the float the reference's sequentially accumulated S_ref rounds to, given |S - S_ref| <= bound?
__device__ __forceinline__ bool mg_round_safe(double S, double bound)
{
	const float f = __double2float_rn(S);
	const uint32_t fb = __float_as_uint(f);
	const uint32_t ex = (fb >> 23) & 0xffu;
	if (ex == 0u || ex == 0xffu || (fb & 0x7fffffu) == 0u) // zero / subnormal / overflow / a power of two (uneven neighbours): no shortcut
		return false;
	const double half_ulp = __hiloint2double((int)((ex + 872u) << 20), 0); // 2^(ex - 127 - 24)
	return fabs(S - (double)f) + bound < half_ulp;
}

// ---- the kernel ----------------------------------------------------------------------------------------------------------
// The window of one row in 32-bit bin coordinates. ok: the window exists; reg: it holds bins; bad: not expressible here
// (coordinates beyond 32 bits): the tile takes the per-row path. w is filled for that path.
struct TileRow {
	uint32_t sb, eb, chrom;
	bool ok, reg, bad;
};

__device__ __forceinline__ void mg_tile_row(const DenseQuery &q, int64_t row, bool live, TileRow &r)
{
	r.sb = r.eb = r.chrom = 0;
	r.ok = r.reg = r.bad = false;
	if (!live)
		return;
	int chrom;
	int64_t s, e, scope;
	mg_get_row(q.rows, row, chrom, s, e, scope);
	const bool inl = q.inl_n > 0; // sizes and bin counts from the kernel parameters (see DenseQuery)
	const int64_t csize = inl ? (int64_t)q.inl_csize[chrom] : __ldg(q.chrom_sizes + chrom);
	int64_t ws = s + q.sshift, we = e + q.eshift;
	ws = ws < 0 ? 0 : ws;
	we = we > csize ? csize : we;
	r.ok = ws < we;
	if (!r.ok)
		return;
	if ((uint64_t)csize <= 0x7fffffffull - q.bin_size && q.bin_magic) { // every coordinate and bin of this chromosome fits 31 bits
		const uint32_t nbins = inl ? q.inl_nbins[chrom] : (uint32_t)q.table[chrom].nbins;
		const uint32_t a = (uint32_t)ws, b = (uint32_t)we + q.bin_size - 1u;
		const uint32_t sb = __umulhi(a, q.bin_magic32) >> q.bin_shift32; // src/GenomeTrackFixedBin.cpp:675 (a, b < 2^31)
		uint32_t eb = __umulhi(b, q.bin_magic32) >> q.bin_shift32;       // :676
		eb = eb < nbins ? eb : nbins;                                      // read_bins_bulk clamp :1036-1041
		r.sb = sb;
		r.eb = eb;
		r.chrom = (uint32_t)chrom;
		r.reg = eb > sb;
	} else
		r.bad = true;
}

// WMAX / WMIN: the max / min columns are requested
template <bool SUM, bool WMAX, bool WMIN, bool Q>
__global__ void __launch_bounds__(MG_TL_THREADS, (Q && WMAX && WMIN) ? 2 : 3) k_dense_tile(const __grid_constant__ DenseQuery q)
{
	constexpr bool MM = WMAX || WMIN;
	constexpr TileLayout L = mg_tile_layout((WMAX ? 1 : 0) + (WMIN ? 1 : 0), Q);
	extern __shared__ __align__(128) uint8_t smem[];
	__shared__ uint64_t s_bar;
	__shared__ __align__(16) uint32_t s_acc[8]; // chrom min, chrom max, min sbin, max ebin, max sbin, max window, any bad row, -
	__shared__ double s_wsum[MG_TL_NW];
	__shared__ uint32_t s_wcnt[MG_TL_NW];
	__shared__ float s_abs;
	const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
	const int64_t i = (int64_t)blockIdx.x * MG_TL_THREADS + t;
	const bool live = i < q.count;
	if (t == 0) {
		mg_mbar_init(&s_bar, 1);
		s_abs = 0.f;
	}
	if (t < 8)
		s_acc[t] = (t == 0 || t == 2 || t == 4) ? 0xffffffffu : 0u;

	// ---- 1. the rows' windows; the tile's bin range, chromosome and record range ----
	TileRow r;
	mg_tile_row(q, q.first + i, live, r);
	const uint32_t sb = r.sb, eb = r.eb, nb = eb - sb;
	const bool ok = r.ok, reg = r.reg;
	__syncthreads(); // the accumulators are initialised
	{
		const uint32_t r_cmin = __reduce_min_sync(0xffffffffu, reg ? r.chrom : 0xffffffffu), r_cmax = __reduce_max_sync(0xffffffffu, r.chrom);
		const uint32_t r_bmin = __reduce_min_sync(0xffffffffu, reg ? sb : 0xffffffffu), r_bmax = __reduce_max_sync(0xffffffffu, eb);
		const uint32_t r_nbmax = MM ? __reduce_max_sync(0xffffffffu, nb) : 0u;
		const bool r_bad = __any_sync(0xffffffffu, r.bad || (Q && nb > MG_BW_S));
		if (lane == 0) {
			atomicMin(&s_acc[0], r_cmin);
			atomicMax(&s_acc[1], r_cmax);
			atomicMin(&s_acc[2], r_bmin);
			atomicMax(&s_acc[3], r_bmax);
			if (MM)
				atomicMax(&s_acc[5], r_nbmax);
			if (r_bad)
				s_acc[6] = 1u;
		}
	}
	__syncthreads();
	const uint4 acc0 = *reinterpret_cast<const uint4 *>(s_acc), acc1 = *reinterpret_cast<const uint4 *>(s_acc + 4);
	const uint32_t cmin = acc0.x, cmax = acc0.y, bmin = acc0.z, bmax = acc0.w, nbmax = acc1.y, anybad = acc1.z;
	const bool anyreg = cmin != 0xffffffffu;
	const uint32_t b0 = bmin & ~3u;                                // 16-byte aligned start of the staged bins
	const uint32_t span4 = anyreg ? ((bmax - b0 + 3u) & ~3u) : 0u; // staged bins (a multiple of 4; the track is padded with NaN)
	// records: from the first window's to the one the last bin of the range would start in (>= the last window's: one reduction less)
	const uint32_t blk0 = bmin / MG_BW_STRIDE, nrec = anyreg ? (bmax - 1u) / MG_BW_STRIDE - blk0 + 1u : 0u;
	const DenseChrom *ch = q.table + cmin;
	bool fast = anyreg && !anybad && cmin == cmax && span4 <= MG_TL_CAP;
	if (Q)
		fast = fast && q.use_bw && nrec <= MG_TL_SLOTS;
	if (!fast) {
		if (live)
			mg_tile_slow_row<SUM, MM, Q>(q, i); // rows without bins included (they report NaN / 0)
		return;
	}

	// ---- 2. bulk copies: the bins, the walked levels of the block records ----
	float *s_vals = reinterpret_cast<float *>(smem + MG_TL_OFF_VALS);
	double *s_pl4 = reinterpret_cast<double *>(smem + MG_TL_OFF_PL4);
	double *s_pb = reinterpret_cast<double *>(smem + MG_TL_OFF_PB);
	uint2 *s_cm = reinterpret_cast<uint2 *>(smem + MG_TL_OFF_CM);
	float *s_ext_mx = reinterpret_cast<float *>(smem + MG_TL_OFF_EXT);
	float *s_ext_mn = reinterpret_cast<float *>(smem + MG_TL_OFF_EXT + (WMAX ? MG_TL_SZ_EXT : 0));
	uint8_t *s_lv = smem + L.off_lv;
	if (t == 0) {
		const float *vsrc = ch->vals;
		const uint32_t vbytes = span4 * 4u;
		mg_mbar_expect_tx(&s_bar, vbytes + (Q ? nrec * MG_BW_STAGE_BYTES : 0u));
		mg_bulk_g2s(s_vals, vsrc + b0, vbytes, &s_bar);
		if (Q) {
			const uint8_t *src = ch->bw + (size_t)blk0 * MG_BW_RECORD_BYTES;
			for (uint32_t sl = 0; sl < nrec; ++sl)
				mg_bulk_g2s(s_lv + sl * MG_BW_STAGE_BYTES, src + (size_t)sl * MG_BW_RECORD_BYTES, MG_BW_STAGE_BYTES, &s_bar);
		}
	}
	for (uint32_t k = span4 + t; k < MG_TL_CAP + 4; k += MG_TL_THREADS) // bins beyond the staged range hold no value
		s_vals[k] = mg_nanf();
	// one warp watches the barrier, the others park at the thread-block barrier (no issue slots spent on polling)
	if (warp == 0)
		mg_mbar_wait(&s_bar, 0);
	__syncthreads();

	// ---- 3a. one chunk of 20 bins per thread ----
	double run = 0;
	uint32_t vmask = 0;
	float cmx = __int_as_float(0xff800000), cmn = __int_as_float(0x7f800000), amax = 0.f;
	{
		const float4 *v4 = reinterpret_cast<const float4 *>(s_vals) + t * MG_TL_QPC;
		float qmx[MG_TL_QPC], qmn[MG_TL_QPC];
#pragma unroll
		for (int qd = 0; qd < MG_TL_QPC; ++qd) {
			const float4 v = v4[qd];
			if (SUM)
				s_pl4[t * MG_TL_QPC + qd] = run;
			const float x[4] = {v.x, v.y, v.z, v.w};
			uint32_t qm = 0;
			const float c0 = mg_clean_bit<1u>(x[0], qm), c1 = mg_clean_bit<2u>(x[1], qm), c2 = mg_clean_bit<4u>(x[2], qm), c3 = mg_clean_bit<8u>(x[3], qm);
			vmask |= qm << (4 * qd);
			if (SUM)
				run = (((run + (double)c0) + (double)c1) + (double)c2) + (double)c3;
			// NaN when the quad holds no value: skipped by every later fmaxf / fminf
			if (WMAX)
				qmx[qd] = fmaxf(fmaxf(fmaxf(x[0], x[1]), x[2]), x[3]);
			if (WMIN)
				qmn[qd] = fminf(fminf(fminf(x[0], x[1]), x[2]), x[3]);
			if (SUM)
				amax = fmaxf(fmaxf(amax, fmaxf(fmaxf(fabsf(x[0]), fabsf(x[1])), fabsf(x[2]))), fabsf(x[3]));
		}
		if (WMAX) {
			float *qsuf = s_ext_mx, *qpre = s_ext_mx + MG_TL_NQD;
			float a = __int_as_float(0xff800000);
#pragma unroll
			for (int qd = 0; qd < MG_TL_QPC; ++qd) {
				a = fmaxf(a, qmx[qd]);
				qpre[t * MG_TL_QPC + qd] = a;
			}
			cmx = a;
			a = __int_as_float(0xff800000);
#pragma unroll
			for (int qd = MG_TL_QPC - 1; qd >= 0; --qd) {
				a = fmaxf(a, qmx[qd]);
				qsuf[t * MG_TL_QPC + qd] = a;
			}
		}
		if (WMIN) {
			float *qsuf = s_ext_mn, *qpre = s_ext_mn + MG_TL_NQD;
			float a = __int_as_float(0x7f800000);
#pragma unroll
			for (int qd = 0; qd < MG_TL_QPC; ++qd) {
				a = fminf(a, qmn[qd]);
				qpre[t * MG_TL_QPC + qd] = a;
			}
			cmn = a;
			a = __int_as_float(0x7f800000);
#pragma unroll
			for (int qd = MG_TL_QPC - 1; qd >= 0; --qd) {
				a = fminf(a, qmn[qd]);
				qsuf[t * MG_TL_QPC + qd] = a;
			}
		}
	}
	// ---- 3b. block scan of the chunk sums / counts; a bound on the tile's sum |v| ----
	const uint32_t cnt = __popc(vmask);
	double incl = run;
	uint32_t cincl = cnt;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const double o = __shfl_up_sync(0xffffffffu, incl, d);
		const uint32_t oc = __shfl_up_sync(0xffffffffu, cincl, d);
		if (lane >= d) {
			incl += o;
			cincl += oc;
		}
	}
	if (SUM) { // sum |v| <= 20 x (largest |v| of the chunk), summed over the chunks
		float a = amax;
#pragma unroll
		for (int m = 16; m > 0; m >>= 1)
			a += __shfl_xor_sync(0xffffffffu, a, m);
		if (lane == 0)
			atomicAdd(&s_abs, a);
	}
	double excl = __shfl_up_sync(0xffffffffu, incl, 1);
	excl = lane ? excl : 0.;
	if (lane == 31) {
		s_wsum[warp] = incl;
		s_wcnt[warp] = cincl;
	}
	__syncthreads();
	double base = 0;
	uint32_t cbase = 0;
	for (int k = 0; k < warp; ++k) { // warp-uniform trip count
		base += s_wsum[k];
		cbase += s_wcnt[k];
	}
	s_pb[t] = base + excl;
	s_cm[t] = make_uint2(cbase + cincl - cnt, vmask);
	if (t == MG_TL_THREADS - 1) {
		s_pb[MG_TL_THREADS] = base + incl;
		s_cm[MG_TL_THREADS] = make_uint2(cbase + cincl, 0u);
		s_pl4[MG_TL_NQD] = 0.;
	}
	// ---- 3c. sparse tables over the chunk extrema: level j entry c = extremum of chunks [c, c + 2^j) ----
	if (MM) {
		// levels needed: a window of nbmax bins holds at most nbmax / 20 whole chunks
		const uint32_t maxrun = nbmax / MG_TL_CHUNK;
		const int nlev = maxrun ? 32 - __clz((int)maxrun) : 1; // floor(log2(maxrun)) + 1
		float *tmx = s_ext_mx + 2 * MG_TL_NQD, *tmn = s_ext_mn + 2 * MG_TL_NQD;
		if (WMAX)
			tmx[t] = cmx;
		if (WMIN)
			tmn[t] = cmn;
		for (int j = 1; j < nlev; ++j) {
			__syncthreads();
			const int h = 1 << (j - 1);
			if (WMAX) {
				const float a = tmx[(j - 1) * MG_TL_THREADS + t];
				tmx[j * MG_TL_THREADS + t] = t + h < MG_TL_THREADS ? fmaxf(a, tmx[(j - 1) * MG_TL_THREADS + t + h]) : a;
			}
			if (WMIN) {
				const float a = tmn[(j - 1) * MG_TL_THREADS + t];
				tmn[j * MG_TL_THREADS + t] = t + h < MG_TL_THREADS ? fminf(a, tmn[(j - 1) * MG_TL_THREADS + t + h]) : a;
			}
		}
	}
	__syncthreads();
	if (!live)
		return;

	// ---- 4. the rows ----
	TileSmem sm;
	sm.vals = s_vals;
	sm.pl4 = s_pl4;
	sm.pb = s_pb;
	sm.cm = s_cm;
	const uint32_t om = q.out_mask; // bit f: column f is requested
	double avg = mg_nan(), sum = mg_nan(), size = mg_nan(), exists = mg_nan(), mn = mg_nan(), mx = mg_nan();
	uint32_t n = 0;
	const uint32_t s = sb - b0, e = eb - b0; // tile-local window (reg rows)
	if (ok) {
		size = 0;
		exists = 0;
	}
	if (reg) {
		double Ps, Pe;
		uint32_t Cs, Ce;
		float4 vs, ve;
		mg_tl_prefix<SUM>(sm, s, Ps, Cs, vs);
		mg_tl_prefix<SUM>(sm, e, Pe, Ce, ve);
		n = Ce - Cs;
		if (SUM) {
			size = (double)n;
			exists = n > 0 ? 1. : 0.;
			if (n > 0) {
				double S = Pe - Ps;
				// |S - S_ref| <= 2^-53 * sum|v| * (adds on our side + adds of the sequential sum); 100: two prefixes through the scan tree,
				// the chunk and the quad, with slack; s_abs >= sum |v| of the tile (float additions: 1.001 covers their rounding)
				const double bound = (double)s_abs * (MG_TL_CHUNK * 1.001 * 1.1102230246251565e-16) * (double)(100u + nb);
				double A = S / (double)n;
				bool safe = true;
				if (om & (1u << MG_SUM))
					safe = mg_round_safe(S, bound);
				if (om & ((1u << MG_AVG) | (1u << MG_NEAREST))) // the quotient moves by bound / n, plus one rounding of the division on either side
					safe = safe && mg_round_safe(A, bound / (double)n + fabs(A) * 2.2204460492503131e-16);
				if (!safe) { // the reference's arithmetic: bins in order (src/GenomeTrackFixedBin.cpp:881-882)
					S = 0;
					for (uint32_t b = s; b < e; ++b) {
						const float v = s_vals[b];
						if (v == v)
							S += (double)v;
					}
					A = S / (double)n;
				}
				avg = mg_f2d(A);
				sum = mg_f2d(S);
			}
		}
		if (MM && n > 0) {
			if (WMAX) {
				TileExt x{s_ext_mx, s_ext_mx + MG_TL_NQD, s_ext_mx + 2 * MG_TL_NQD};
				mx = (double)mg_tl_extremum<true>(sm, x, s, e, vs, ve);
			}
			if (WMIN) {
				TileExt x{s_ext_mn, s_ext_mn + MG_TL_NQD, s_ext_mn + 2 * MG_TL_NQD};
				mn = (double)mg_tl_extremum<false>(sm, x, s, e, vs, ve);
			}
		}
	}
	if (SUM) {
		if (om & (1u << MG_AVG)) mg_st_stream(q.out[MG_AVG] + i, avg);
		if (om & (1u << MG_NEAREST)) mg_st_stream(q.out[MG_NEAREST] + i, avg);
		if (om & (1u << MG_SUM)) mg_st_stream(q.out[MG_SUM] + i, sum);
		if (om & (1u << MG_SIZE)) mg_st_stream(q.out[MG_SIZE] + i, size);
		if (om & (1u << MG_EXISTS)) mg_st_stream(q.out[MG_EXISTS] + i, exists);
	}
	if (WMIN) mg_st_stream(q.out[MG_MIN] + i, mn);
	if (WMAX) mg_st_stream(q.out[MG_MAX] + i, mx);
	if (Q) {
		const uint32_t blk = sb / MG_BW_STRIDE, l = sb - blk * MG_BW_STRIDE;
		const uint8_t *lv = s_lv + (size_t)(blk - blk0) * MG_BW_STAGE_BYTES;
		const uint16_t *pos = reinterpret_cast<const uint16_t *>(ch->bw + (size_t)blk * MG_BW_RECORD_BYTES + MG_BW_POS_OFFSET);
		const float *vblk = s_vals + ((int)(blk * MG_BW_STRIDE) - (int)b0); // block-local position -> tile bin
		for (int k = 0; k < q.nq; ++k) {
			double qv = mg_nan();
			if (n > 0) {
				const double index = __dmul_rn((double)(n - 1), q.q_pct[k]); // q_pct is clamped to [0, 1] by the host
				const double fl = floor(index);
				const bool want2 = index != fl;
				float x1, x2;
				mg_bw_select<false>(lv, pos, vblk, l, l + nb, (uint32_t)fl, want2, x1, x2);
				qv = mg_interp(x1, x2, __dsub_rn(index, fl));
			}
			mg_st_stream(q.q_out[k] + i, qv);
		}
	}
}

// ---- k_dense_thin: the staging of k_dense_tile without its cooperative passes ---------------------------------------------------
// The thread block stages its rows' bins and block records exactly as k_dense_tile does, but builds nothing: a row sums its two
// partial groups of 8 bins out of shared memory and takes the whole groups in between from the chromosome's p8 prefix and sparse
// tables (dense.cuh / dense_index.cuh) -- four independent global loads issued BEFORE the wait for the bulk copies, so they fly with
// them. After the wait a row touches global memory once more (the 64 bytes of pos_of_rank behind its quantile); the bins under the
// quantile come from the staged copy. Against k_dense_rowquery: three dependent global round trips per row instead of six, no
// sector over-fetch on the track; against k_dense_tile: ~350 fewer instructions per row and half the shared memory (5-6 resident
// thread blocks per SM), at the price of the p8 / sparse-table traffic (2.5 B/bin) and of sums that are not re-summed to the
// reference's bits (they are within 1e-6, like k_dense_rowquery's).
// Measured and dropped: an L2 prefetch hint (one warp pulling the rows, bins, records and prefix entries of the tile one or two
// waves of resident thread blocks ahead into L2): 0.402 -> 0.485 / 0.551 ms for the C3 step -- the kernel is short of DRAM bandwidth,
// not of latency tolerance, and the hint's own traffic and issue slots cost more than the shorter round trips return.
#ifndef MG_TH_MB
#define MG_TH_MB 6 // measured on B200 at the C3 shape: 4 / 5 / 6 resident thread blocks per SM -> 0.456 / 0.412 / 0.398 ms
#endif
#define MG_TH_SZ_VALS ((MG_TL_CAP + 8) * 4)

template <bool SUM, bool MM>
__device__ __forceinline__ void mg_part8_smem(const float *__restrict__ svals /* staged bins at the group's first bin */, int lo, int hi, double &S, int &n,
											  float &mx, float &mn)
{
	const float4 a = reinterpret_cast<const float4 *>(svals)[0], b = reinterpret_cast<const float4 *>(svals)[1];
	const uint32_t sel = ((1u << hi) - 1u) & ~((1u << lo) - 1u); // bins lo .. hi - 1
	const float c0 = mg_take<0>(a.x, sel, n), c1 = mg_take<1>(a.y, sel, n), c2 = mg_take<2>(a.z, sel, n), c3 = mg_take<3>(a.w, sel, n);
	const float c4 = mg_take<4>(b.x, sel, n), c5 = mg_take<5>(b.y, sel, n), c6 = mg_take<6>(b.z, sel, n), c7 = mg_take<7>(b.w, sel, n);
	if (SUM) // bin order
		S = (((((((S + (double)c0) + (double)c1) + (double)c2) + (double)c3) + (double)c4) + (double)c5) + (double)c6) + (double)c7;
	if (MM) {
		const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			const bool in = (sel >> j) & 1u; // NaN operands are skipped by fmaxf / fminf
			mx = fmaxf(mx, in ? v[j] : __int_as_float(0xff800000));
			mn = fminf(mn, in ? v[j] : __int_as_float(0x7f800000));
		}
	}
}

// the same with the selection already made: bit j of m = bin j of the group is inside the window and holds a value.
// Inline PTX for the reason mg_take gives: a predicate from the mask bit and selects on the FLOAT, before the conversion to double.
template <uint32_t J>
__device__ __forceinline__ void mg_pick(float x, uint32_t m, float &zero_or_x, float &ninf_or_x, float &pinf_or_x)
{
	asm("{\n\t.reg .pred p;\n\t.reg .u32 t;\n\tand.b32 t, %4, %5;\n\tsetp.ne.u32 p, t, 0;\n\tselp.f32 %0, %3, 0f00000000, p;\n\t"
		"selp.f32 %1, %3, 0fFF800000, p;\n\tselp.f32 %2, %3, 0f7F800000, p;\n\t}"
		: "=f"(zero_or_x), "=f"(ninf_or_x), "=f"(pinf_or_x)
		: "f"(x), "r"(m), "n"(1u << J));
}
template <bool SUM, bool MM>
__device__ __forceinline__ void mg_part8_mask(const float *__restrict__ svals, uint32_t m, double &S, float &mx, float &mn)
{
	const float4 a = reinterpret_cast<const float4 *>(svals)[0], b = reinterpret_cast<const float4 *>(svals)[1];
	float c[8], h[8], l[8];
	mg_pick<0>(a.x, m, c[0], h[0], l[0]);
	mg_pick<1>(a.y, m, c[1], h[1], l[1]);
	mg_pick<2>(a.z, m, c[2], h[2], l[2]);
	mg_pick<3>(a.w, m, c[3], h[3], l[3]);
	mg_pick<4>(b.x, m, c[4], h[4], l[4]);
	mg_pick<5>(b.y, m, c[5], h[5], l[5]);
	mg_pick<6>(b.z, m, c[6], h[6], l[6]);
	mg_pick<7>(b.w, m, c[7], h[7], l[7]);
	if (SUM) // bin order
		S = (((((((S + (double)c[0]) + (double)c[1]) + (double)c[2]) + (double)c[3]) + (double)c[4]) + (double)c[5]) + (double)c[6]) + (double)c[7];
	if (MM) {
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			mx = fmaxf(mx, h[j]);
			mn = fminf(mn, l[j]);
		}
	}
}

#ifdef MG_DBG_BLOCKTIME
// development build only (tools/dbg_slow_blocks.py): the longest thread lifetime (ns) of every thread block of the last k_dense_thin launch
__device__ unsigned long long g_dbg_blocktime[1 << 16];
struct MgDbgTimer {
	unsigned long long t0;
	__device__ MgDbgTimer() { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0)); }
	__device__ ~MgDbgTimer()
	{
		unsigned long long t1;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
		if (blockIdx.x < (1u << 16))
			atomicMax(&g_dbg_blocktime[blockIdx.x], t1 - t0);
	}
};
#endif
template <bool SUM, bool WMAX, bool WMIN, bool Q>
__global__ void __launch_bounds__(MG_TL_THREADS, MG_TH_MB) k_dense_thin(const __grid_constant__ DenseQuery q)
{
	constexpr bool MM = WMAX || WMIN;
#ifdef MG_DBG_BLOCKTIME
	MgDbgTimer dbg_timer;
#endif
	extern __shared__ __align__(128) uint8_t smem[];
	__shared__ uint64_t s_bar;
	__shared__ __align__(16) uint32_t s_acc[8]; // chrom min, chrom max, min sbin, max ebin, shortest run of whole groups, -, any bad row, -
	const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
	const int64_t i = (int64_t)blockIdx.x * MG_TL_THREADS + t;
	const bool live = i < q.count;
	if (t == 0)
		mg_mbar_init(&s_bar, 1);
	if (t < 8)
		s_acc[t] = (t == 0 || t == 2 || t == 4) ? 0xffffffffu : 0u;

	// ---- 1. the rows' windows; the tile's bin range, chromosome and record range ----
	TileRow r;
	mg_tile_row(q, q.first + i, live, r);
	const uint32_t sb = r.sb, eb = r.eb, nb = eb - sb;
	const bool ok = r.ok, reg = r.reg;
	const uint32_t g0 = sb >> 3, g1 = reg ? (eb - 1u) >> 3 : 0u;
	const bool multi = reg && g0 != g1;
	const uint32_t gs = g0 + 1, len = multi ? g1 - gs : 0u; // the whole groups of 8 inside the window: [gs, g1)
	__syncthreads(); // the accumulators are initialised
	{
		// every thread hands its values to the shared-memory reductions: the compiler folds a warp's 32 into one (CREDUX + one elected
		// lane's ATOMS) -- a reduction written out here ahead of a single lane's atomic would be folded a second time
		atomicMin(&s_acc[0], reg ? r.chrom : 0xffffffffu);
		atomicMax(&s_acc[1], r.chrom);
		atomicMin(&s_acc[2], reg ? sb : 0xffffffffu);
		atomicMax(&s_acc[3], eb);
		if (MM)
			atomicMin(&s_acc[4], len > 0 ? len : 0xffffffffu);
		if (r.bad || (Q && nb > MG_BW_S))
			s_acc[6] = 1u;
	}
	__syncthreads();
	const uint4 acc0 = *reinterpret_cast<const uint4 *>(s_acc);
	const uint32_t cmin = acc0.x, cmax = acc0.y, bmin = acc0.z, bmax = acc0.w, lmin = s_acc[4], anybad = s_acc[6];
	const bool anyreg = cmin != 0xffffffffu;
	const uint32_t b0 = bmin & ~7u;                                // group-aligned start of the staged bins
	const uint32_t span8 = anyreg ? ((bmax - b0 + 7u) & ~7u) : 0u; // staged bins: whole groups of 8 (the track is padded with NaN)
	const uint32_t blk0 = bmin / MG_BW_STRIDE, nrec = anyreg ? (bmax - 1u) / MG_BW_STRIDE - blk0 + 1u : 0u;
	const DenseChrom *ch = q.table + cmin;
	const bool inl = q.inl_n > 0 && anyreg;
	const uint8_t *bw_base = Q ? (inl ? q.inl_bw[cmin] : (anyreg ? ch->bw : nullptr)) : nullptr;
	const Pref *p8_base = inl ? q.inl_p8[cmin] : (anyreg ? ch->p8 : nullptr);
	bool fast = anyreg && !anybad && cmin == cmax && span8 <= MG_TL_CAP;
	if (Q)
		fast = fast && q.use_bw && nrec <= MG_TL_SLOTS;
	if (!fast) {
		if (live)
			mg_tile_slow_row<SUM, MM, Q>(q, i);
		return;
	}

	// ---- 2. bulk copies: the bins, the walked levels of the block records ----
	float *s_vals = reinterpret_cast<float *>(smem);
	uint8_t *s_lv = smem + MG_TH_SZ_VALS;
	if (t == 0) {
		const float *vsrc = inl ? q.inl_vals[cmin] : ch->vals;
		const uint32_t vbytes = span8 * 4u;
		mg_mbar_expect_tx(&s_bar, vbytes + (Q ? nrec * MG_BW_STAGE_BYTES : 0u));
		mg_bulk_g2s(s_vals, vsrc + b0, vbytes, &s_bar);
		if (Q) {
			const uint8_t *src = bw_base + (size_t)blk0 * MG_BW_RECORD_BYTES;
			for (uint32_t sl = 0; sl < nrec; ++sl)
				mg_bulk_g2s(s_lv + sl * MG_BW_STAGE_BYTES, src + (size_t)sl * MG_BW_RECORD_BYTES, MG_BW_STAGE_BYTES, &s_bar);
		}
	}
	// what the whole groups of the window contribute, from the chromosome's structures: loads that fly with the copies (their values are
	// looked at after the wait). Sums and counts: two p8 entries. Extrema: the sparse table, at ONE level for the whole tile -- the level of
	// the tile's shortest run, 2 to 4 lookups a row: rows whose runs straddle a power of two (C3: 63 - 65 groups) would otherwise stream
	// two levels of the table through the device.
	int4 pa_raw = make_int4(0, 0, 0, (int)0xffff0000u), pb_raw = pa_raw; // entry g0 + 1 also says which bins of group g0 hold no value
	float mx0 = __int_as_float(0xff800000), mx1 = mx0, mx2 = mx0, mx3 = mx0;
	float mn0 = __int_as_float(0x7f800000), mn1 = mn0, mn2 = mn0, mn3 = mn0;
	double abs_sum = 0; // scale of the cancellation guard below
	if (reg) {
		pa_raw = __ldg(reinterpret_cast<const int4 *>(p8_base + g0 + 1));
		if (multi) {
			pb_raw = __ldg(reinterpret_cast<const int4 *>(p8_base + g1));
			if (MM && len > 0) { // lmin <= len < 2^MG_ST_LEVELS: the tile holds MG_TL_CAP bins
				int j = 31 - __clz((int)lmin);
				uint32_t nl = (len + (1u << j) - 1u) >> j;
				if (nl > 4u) { // far longer than the tile's shortest: the row's own level
					j = 31 - __clz((int)len);
					nl = 2u;
				}
				const uint32_t w = 1u << j;
				// the level's address from the kernel parameters when they hold the table: no pointer load ahead of the entries
				if (WMAX) {
					const float *lvl = inl ? q.inl_smax[cmin] + (size_t)j * q.inl_stpitch[cmin] : ch->smax[j];
					mx0 = __ldg(lvl + gs);
					mx1 = __ldg(lvl + g1 - w);
					if (nl > 2u)
						mx2 = __ldg(lvl + gs + w);
					if (nl > 3u)
						mx3 = __ldg(lvl + gs + 2u * w);
				}
				if (WMIN) {
					const float *lvl = inl ? q.inl_smin[cmin] + (size_t)j * q.inl_stpitch[cmin] : ch->smin[j];
					mn0 = __ldg(lvl + gs);
					mn1 = __ldg(lvl + g1 - w);
					if (nl > 2u)
						mn2 = __ldg(lvl + gs + w);
					if (nl > 3u)
						mn3 = __ldg(lvl + gs + 2u * w);
				}
			}
			if (SUM)
				abs_sum = ch->abs_sum;
		}
	}
	// one warp watches the barrier, the others park at the thread-block barrier (no issue slots spent on polling)
	if (warp == 0)
		mg_mbar_wait(&s_bar, 0);
	__syncthreads();
	if (!live)
		return;

	// ---- 3. the row ----
	// the window's values are counted from the two entries: their difference and the no-value masks of the two partial groups
	const double pa_sum = __hiloint2double(pa_raw.y, pa_raw.x), pb_sum = __hiloint2double(pb_raw.y, pb_raw.x);
	const uint32_t sel_h = reg ? (((1u << (multi ? 8u : eb - 8u * g0)) - 1u) & ~((1u << (sb & 7u)) - 1u)) : 0u; // bins of group g0 inside the window
	const uint32_t sel_t = multi ? (1u << (eb - 8u * g1)) - 1u : 0u;                                            // bins of group g1
	const uint32_t mh = sel_h & ~((uint32_t)pa_raw.w >> 24), mt = sel_t & ~((uint32_t)pb_raw.w >> 16); // ... that hold a value
	const uint32_t n = __popc(mh) + __popc(mt) + (multi ? (uint32_t)pb_raw.z - (uint32_t)pa_raw.z : 0u);
	const uint32_t om = q.out_mask; // bit f: column f is requested
	double avg = mg_nan(), sum = mg_nan(), size = mg_nan(), exists = mg_nan(), mn = mg_nan(), mx = mg_nan();
	if (ok) {
		size = 0;
		exists = 0;
	}
	if (reg && (SUM || MM)) {
		double S = 0;
		float pmx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)), pmn = fminf(fminf(mn0, mn1), fminf(mn2, mn3));
		mg_part8_mask<SUM, MM>(s_vals + (8u * g0 - b0), mh, S, pmx, pmn);
		if (multi) {
			double St = 0;
			mg_part8_mask<SUM, MM>(s_vals + (8u * g1 - b0), mt, St, pmx, pmn);
			S = (S + (pb_sum - pa_sum)) + St;
			// the prefix difference has lost too many digits against the chromosome's running sum: the bins in order, from the staged copy
			if (SUM && g0 + 1 < g1 && n > 0 && fabs(S) < abs_sum * 1.4901161193847656e-08 /* 2^-26 */) {
				S = 0;
#pragma unroll 8
				for (uint32_t b = sb - b0; b < eb - b0; ++b) {
					const float v = s_vals[b];
					if (v == v)
						S += (double)v;
				}
			}
		}
		if (SUM) {
			size = (double)n;
			exists = n > 0 ? 1. : 0.;
			if (n > 0) {
				avg = mg_f2d(S / (double)n);
				sum = mg_f2d(S);
			}
		}
		if (MM && n > 0) {
			mx = (double)pmx;
			mn = (double)pmn;
		}
	}
	if (SUM) {
		constexpr uint32_t GROUP = (1u << MG_AVG) | (1u << MG_NEAREST) | (1u << MG_SUM) | (1u << MG_SIZE) | (1u << MG_EXISTS);
		if ((om & GROUP) == (1u << MG_AVG)) // the usual request, without the other four columns' tests and addresses in its way
			mg_st_stream(q.out[MG_AVG] + i, avg);
		else {
			if (om & (1u << MG_AVG)) mg_st_stream(q.out[MG_AVG] + i, avg);
			if (om & (1u << MG_NEAREST)) mg_st_stream(q.out[MG_NEAREST] + i, avg);
			if (om & (1u << MG_SUM)) mg_st_stream(q.out[MG_SUM] + i, sum);
			if (om & (1u << MG_SIZE)) mg_st_stream(q.out[MG_SIZE] + i, size);
			if (om & (1u << MG_EXISTS)) mg_st_stream(q.out[MG_EXISTS] + i, exists);
		}
	}
	if (WMIN) mg_st_stream(q.out[MG_MIN] + i, mn);
	if (WMAX) mg_st_stream(q.out[MG_MAX] + i, mx);
	if (Q) {
		const uint32_t blk = sb / MG_BW_STRIDE, l = sb - blk * MG_BW_STRIDE;
		const uint8_t *lv = s_lv + (size_t)(blk - blk0) * MG_BW_STAGE_BYTES;
		const uint16_t *pos = reinterpret_cast<const uint16_t *>(bw_base + (size_t)blk * MG_BW_RECORD_BYTES + MG_BW_POS_OFFSET);
		const float *vblk = s_vals + ((int)(blk * MG_BW_STRIDE) - (int)b0); // block-local position -> staged bin
		for (int k = 0; k < q.nq; ++k) {
			double qv = mg_nan();
			if (n > 0) {
				const double index = __dmul_rn((double)(n - 1), q.q_pct[k]); // q_pct is clamped to [0, 1] by the host
				const double fl = floor(index);
				const bool want2 = index != fl;
				float x1, x2;
				mg_bw_select<false>(lv, pos, vblk, l, l + nb, (uint32_t)fl, want2, x1, x2);
				qv = mg_interp(x1, x2, __dsub_rn(index, fl));
			}
			mg_st_stream(q.q_out[k] + i, qv);
		}
	}
}

// ---- k_dense_pipe: k_dense_thin as a persistent, software-pipelined thread block ---------------------------------------------------
// k_dense_thin's thread block lives through four dependent memory round trips (rows, bulk copies + prefix entries, pos_of_rank, stores)
// and only the blocks that happen to sit in their copy phase keep bytes in flight: ~50 KB per SM, which at the loaded DRAM latency
// is ~5 TB/s -- the measured 4.85. Here a thread block walks tiles blockIdx.x, + gridDim.x, ... with TWO stage buffers: while the
// rows of tile t are computed out of one buffer, the bulk copies of tile t + G land in the other, the prefix / sparse-table entries
// of tile t + G and the rows of tile t + 2G are in flight in registers. Every resident thread block always has a tile's worth of
// copies outstanding, and never waits for one it has just issued.
// resident thread blocks per SM: DIST + 1 stage buffers each, DIST tiles of copies in flight per thread block
#define MG_PP_BPS(DIST) ((DIST) == 1 ? 3 : 2)
#define MG_PP_STAGE(Q) (MG_TH_SZ_VALS + ((Q) ? MG_TL_SLOTS * MG_BW_STAGE_BYTES : 0))

struct PipeRaw { // a row as loaded, before the window transform
	int32_t chrom;
	int64_t s, e;
	bool live;
};
struct PipeGeo { // what a tile's rows share (uniform over the thread block)
	uint32_t b0, blk0, cmin;
	bool fast;
};
struct PipePre { // a row's whole groups from the chromosome's structures, AS LOADED: combining them where they are issued would wait for
	int4 a, b;    // the loads there, one DRAM round trip per tile; they are consumed an iteration later
	float mx0, mx1, mn0, mn1;
};
__device__ __forceinline__ void mg_pipe_pre_clear(PipePre &p)
{
	p.a = p.b = make_int4(0, 0, 0, 0);
	p.mx0 = p.mx1 = __int_as_float(0xff800000);
	p.mn0 = p.mn1 = __int_as_float(0x7f800000);
}

__device__ __forceinline__ void mg_pipe_load(const DenseQuery &q, int64_t i, PipeRaw &r)
{
	r.live = i < q.count;
	r.chrom = 0;
	r.s = r.e = 0;
	if (r.live) {
		int64_t scope;
		int chrom;
		mg_get_row(q.rows, q.first + i, chrom, r.s, r.e, scope);
		r.chrom = chrom;
	}
}

__device__ __forceinline__ void mg_pipe_window(const DenseQuery &q, const PipeRaw &raw, TileRow &r)
{
	r.sb = r.eb = r.chrom = 0;
	r.ok = r.reg = r.bad = false;
	if (!raw.live)
		return;
	const int64_t csize = __ldg(q.chrom_sizes + raw.chrom);
	int64_t ws = raw.s + q.sshift, we = raw.e + q.eshift;
	ws = ws < 0 ? 0 : ws;
	we = we > csize ? csize : we;
	r.ok = ws < we;
	if (!r.ok)
		return;
	if ((uint64_t)csize <= 0x7fffffffull - q.bin_size && q.bin_magic) {
		const uint32_t nbins = (uint32_t)q.table[raw.chrom].nbins;
		const uint32_t a = (uint32_t)ws, b = (uint32_t)we + q.bin_size - 1u;
		const uint32_t sb = (uint32_t)__umul64hi((uint64_t)a, q.bin_magic);
		uint32_t eb = (uint32_t)__umul64hi((uint64_t)b, q.bin_magic);
		eb = eb < nbins ? eb : nbins;
		r.sb = sb;
		r.eb = eb;
		r.chrom = (uint32_t)raw.chrom;
		r.reg = eb > sb;
	} else
		r.bad = true;
}

template <bool SUM, bool WMAX, bool WMIN, bool Q, int DIST>
__global__ void __launch_bounds__(MG_TL_THREADS, MG_PP_BPS(DIST)) k_dense_pipe(const __grid_constant__ DenseQuery q)
{
	constexpr bool MM = WMAX || WMIN;
	constexpr uint32_t STAGE = MG_PP_STAGE(Q);
	constexpr int NBUF = DIST + 1;
	extern __shared__ __align__(128) uint8_t smem[];
	__shared__ uint64_t s_bar[NBUF];
	__shared__ __align__(16) uint32_t s_acc[2][8]; // chrom min, chrom max, min sbin, max ebin, -, -, any bad row, -
	const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
	const int64_t ntiles = (q.count + MG_TL_THREADS - 1) / MG_TL_THREADS;
	const int64_t G = gridDim.x;
	if (t == 0) {
#pragma unroll
		for (int b = 0; b < NBUF; ++b)
			mg_mbar_init(&s_bar[b], 1);
	}
	if (t < 16)
		(&s_acc[0][0])[t] = ((t & 7) == 0 || (t & 7) == 2) ? 0xffffffffu : 0u;
	__syncthreads();

	uint32_t phase = 0; // bit b: the parity the next wait on buffer b uses
	// what travels from one iteration to the next: the windows / geometry / prefix entries of the DIST tiles whose copies are in
	// flight (slot 0 = the tile computed next), and the rows of the tile after those
	TileRow rn[DIST];
	PipeGeo gn[DIST];
	PipePre pn[DIST];
	PipeRaw raw;
#pragma unroll
	for (int d = 0; d < DIST; ++d) {
		rn[d].sb = rn[d].eb = rn[d].chrom = 0;
		rn[d].ok = rn[d].reg = rn[d].bad = false;
		gn[d].b0 = gn[d].blk0 = gn[d].cmin = 0;
		gn[d].fast = false;
		mg_pipe_pre_clear(pn[d]);
	}
	mg_pipe_load(q, (int64_t)blockIdx.x * MG_TL_THREADS + t, raw);
	int buf = 0, ibuf = 0; // buffer of the tile computed in this iteration / of the tile whose copies this iteration issues

	// iterations -DIST .. -1 are the prologue: they only issue copies
	for (int it = -DIST;; ++it) {
		const int64_t tile = (int64_t)blockIdx.x + (int64_t)it * G, itile = tile + (int64_t)DIST * G;
		if (it >= 0 && tile >= ntiles)
			break;
		const TileRow r = rn[0];
		const PipeGeo g = gn[0];
		const PipePre p = pn[0];
#pragma unroll
		for (int d = 0; d + 1 < DIST; ++d) {
			rn[d] = rn[d + 1];
			gn[d] = gn[d + 1];
			pn[d] = pn[d + 1];
		}
		// ---- A. tile itile: windows, geometry, bulk copies into buffer ibuf, prefix / sparse-table loads ----
		if (itile < ntiles) {
			TileRow &rr = rn[DIST - 1];
			PipeGeo &gg = gn[DIST - 1];
			PipePre &pp = pn[DIST - 1];
			mg_pipe_window(q, raw, rr);
			const bool reg = rr.reg;
			const uint32_t nb = rr.eb - rr.sb;
			uint32_t *acc = s_acc[it & 1];
			{
				const uint32_t r_cmin = __reduce_min_sync(0xffffffffu, reg ? rr.chrom : 0xffffffffu), r_cmax = __reduce_max_sync(0xffffffffu, rr.chrom);
				const uint32_t r_bmin = __reduce_min_sync(0xffffffffu, reg ? rr.sb : 0xffffffffu), r_bmax = __reduce_max_sync(0xffffffffu, rr.eb);
				const bool r_bad = __any_sync(0xffffffffu, rr.bad || (Q && nb > MG_BW_S));
				if (lane == 0) {
					atomicMin(&acc[0], r_cmin);
					atomicMax(&acc[1], r_cmax);
					atomicMin(&acc[2], r_bmin);
					atomicMax(&acc[3], r_bmax);
					if (r_bad)
						acc[6] = 1u;
				}
			}
			__syncthreads(); // B1: also, every thread is through with the tile that used buffer ibuf
			const uint4 acc0 = *reinterpret_cast<const uint4 *>(acc);
			const uint32_t cmin = acc0.x, cmax = acc0.y, bmin = acc0.z, bmax = acc0.w, anybad = acc[6];
			if (t < 8) // the other set serves the next iteration; its last readers are behind B1
				s_acc[(it & 1) ^ 1][t] = (t == 0 || t == 2) ? 0xffffffffu : 0u;
			const bool anyreg = cmin != 0xffffffffu;
			const uint32_t b0 = bmin & ~7u;
			const uint32_t span8 = anyreg ? ((bmax - b0 + 7u) & ~7u) : 0u;
			const uint32_t blk0 = bmin / MG_BW_STRIDE, nrec = anyreg ? (bmax - 1u) / MG_BW_STRIDE - blk0 + 1u : 0u;
			bool fast = anyreg && !anybad && cmin == cmax && span8 <= MG_TL_CAP;
			if (Q)
				fast = fast && q.use_bw && nrec <= MG_TL_SLOTS;
			gg.b0 = b0;
			gg.blk0 = blk0;
			gg.cmin = cmin;
			gg.fast = fast;
			mg_pipe_pre_clear(pp);
			if (fast) {
				const DenseChrom *ch = q.table + cmin;
				if (t == 0) {
					uint8_t *dst = smem + (size_t)ibuf * STAGE;
					const uint32_t vbytes = span8 * 4u;
					mg_mbar_expect_tx(&s_bar[ibuf], vbytes + (Q ? nrec * MG_BW_STAGE_BYTES : 0u));
					mg_bulk_g2s(dst, ch->vals + b0, vbytes, &s_bar[ibuf]);
					if (Q) {
						const uint8_t *src = ch->bw + (size_t)blk0 * MG_BW_RECORD_BYTES;
						for (uint32_t sl = 0; sl < nrec; ++sl)
							mg_bulk_g2s(dst + MG_TH_SZ_VALS + sl * MG_BW_STAGE_BYTES, src + (size_t)sl * MG_BW_RECORD_BYTES, MG_BW_STAGE_BYTES, &s_bar[ibuf]);
					}
				}
				const uint32_t g0 = rr.sb >> 3, g1 = reg ? (rr.eb - 1u) >> 3 : 0u;
				if (reg && g0 != g1) {
					pp.a = __ldg(reinterpret_cast<const int4 *>(ch->p8 + g0 + 1));
					pp.b = __ldg(reinterpret_cast<const int4 *>(ch->p8 + g1));
					const uint32_t gs = g0 + 1, len = g1 - gs;
					if (MM && len > 0) {
						const int j = 31 - __clz((int)len);
						if (WMAX) {
							pp.mx0 = __ldg(ch->smax[j] + gs);
							pp.mx1 = __ldg(ch->smax[j] + g1 - (1u << j));
						}
						if (WMIN) {
							pp.mn0 = __ldg(ch->smin[j] + gs);
							pp.mn1 = __ldg(ch->smin[j] + g1 - (1u << j));
						}
					}
				}
			}
			// the rows of the tile after that
			mg_pipe_load(q, (itile + G) * MG_TL_THREADS + t, raw);
		}
		ibuf = ibuf + 1 == NBUF ? 0 : ibuf + 1;
		if (it < 0)
			continue;
		// ---- B. tile `tile`: out of buffer buf ----
		const int cb = buf;
		buf = buf + 1 == NBUF ? 0 : buf + 1;
		const int64_t i = tile * MG_TL_THREADS + t;
		const bool live = i < q.count;
		if (!g.fast) {
			if (live)
				mg_tile_slow_row<SUM, MM, Q>(q, i);
			__syncthreads(); // B2
			continue;
		}
		if (warp == 0)
			mg_mbar_wait(&s_bar[cb], (phase >> cb) & 1u);
		phase ^= 1u << cb;
		__syncthreads(); // B2
		if (!live)
			continue;
		const float *s_vals = reinterpret_cast<const float *>(smem + (size_t)cb * STAGE);
		const uint8_t *s_lv = smem + (size_t)cb * STAGE + MG_TH_SZ_VALS;
		const DenseChrom *ch = q.table + g.cmin;
		const uint32_t sb = r.sb, eb = r.eb, nb = eb - sb, b0 = g.b0;
		const bool ok = r.ok, reg = r.reg;
		const uint32_t om = q.out_mask;
		double avg = mg_nan(), sum = mg_nan(), size = mg_nan(), exists = mg_nan(), mn = mg_nan(), mx = mg_nan();
		uint32_t n = 0;
		if (ok) {
			size = 0;
			exists = 0;
		}
		if (reg) {
			const uint32_t g0 = sb >> 3, g1 = (eb - 1u) >> 3;
			double S = 0;
			int nh = 0, nt = 0;
			float pmx = fmaxf(p.mx0, p.mx1), pmn = fminf(p.mn0, p.mn1);
			const float *gh = s_vals + (8u * g0 - b0);
			if (g0 == g1) {
				mg_part8_smem<SUM, MM>(gh, (int)(sb & 7u), (int)(eb - 8u * g0), S, nh, pmx, pmn);
				n = (uint32_t)nh;
			} else {
				double St = 0;
				mg_part8_smem<SUM, MM>(gh, (int)(sb & 7u), 8, S, nh, pmx, pmn);
				mg_part8_smem<SUM, MM>(s_vals + (8u * g1 - b0), 0, (int)(eb - 8u * g1), St, nt, pmx, pmn);
				S = (S + (__hiloint2double(p.b.y, p.b.x) - __hiloint2double(p.a.y, p.a.x))) + St;
				n = (uint32_t)nh + ((uint32_t)p.b.z - (uint32_t)p.a.z) + (uint32_t)nt;
				if (SUM && g0 + 1 < g1 && n > 0 && fabs(S) < ch->abs_sum * 1.4901161193847656e-08 /* 2^-26 */) {
					S = 0;
					for (uint32_t b = sb - b0; b < eb - b0; ++b) {
						const float v = s_vals[b];
						if (v == v)
							S += (double)v;
					}
				}
			}
			if (SUM) {
				size = (double)n;
				exists = n > 0 ? 1. : 0.;
				if (n > 0) {
					avg = mg_f2d(S / (double)n);
					sum = mg_f2d(S);
				}
			}
			if (MM && n > 0) {
				mx = (double)pmx;
				mn = (double)pmn;
			}
		}
		if (SUM) {
			if (om & (1u << MG_AVG)) mg_st_stream(q.out[MG_AVG] + i, avg);
			if (om & (1u << MG_NEAREST)) mg_st_stream(q.out[MG_NEAREST] + i, avg);
			if (om & (1u << MG_SUM)) mg_st_stream(q.out[MG_SUM] + i, sum);
			if (om & (1u << MG_SIZE)) mg_st_stream(q.out[MG_SIZE] + i, size);
			if (om & (1u << MG_EXISTS)) mg_st_stream(q.out[MG_EXISTS] + i, exists);
		}
		if (WMIN) mg_st_stream(q.out[MG_MIN] + i, mn);
		if (WMAX) mg_st_stream(q.out[MG_MAX] + i, mx);
		if (Q) {
			const uint32_t blk = sb / MG_BW_STRIDE, l = sb - blk * MG_BW_STRIDE;
			const uint8_t *lv = s_lv + (size_t)(blk - g.blk0) * MG_BW_STAGE_BYTES;
			const uint16_t *pos = reinterpret_cast<const uint16_t *>(ch->bw + (size_t)blk * MG_BW_RECORD_BYTES + MG_BW_POS_OFFSET);
			const float *vblk = s_vals + ((int)(blk * MG_BW_STRIDE) - (int)b0);
			for (int k = 0; k < q.nq; ++k) {
				double qv = mg_nan();
				if (n > 0) {
					const double index = __dmul_rn((double)(n - 1), q.q_pct[k]);
					const double fl = floor(index);
					const bool want2 = index != fl;
					float x1, x2;
					mg_bw_select<false>(lv, pos, vblk, l, l + nb, (uint32_t)fl, want2, x1, x2);
					qv = mg_interp(x1, x2, __dsub_rn(index, fl));
				}
				mg_st_stream(q.q_out[k] + i, qv);
			}
		}
	}
}
