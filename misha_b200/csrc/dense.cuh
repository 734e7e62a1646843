// dense.cuh -- dense fixed-bin tracks: staging (clean + prefix build) and per-row window functions.
//
// HBM layout per chromosome:
//   vals[nbins (+pad)]  float32, +-inf already turned into NaN (read_bins_bulk, src/GenomeTrackFixedBin.cpp:1066-1069)
//   p8[ngroups + 1]     16-byte entries {double sum, int64 count} of the non-NaN vals[0..8g): one entry per group of 8 bins
//                       (2 bytes per bin: rows that are 15 bins apart stream this array instead of gathering 64-byte
//                       DRAM granules out of a 16-bytes-per-bin one)
// avg / sum / size / exists of a window = partial head group + (p8 difference over the full groups) + partial tail group;
// the two partial groups are the same 32-byte sectors the min / max pyramid reads at level 0 (mg_win_basic).
#pragma once
#include "common.cuh"
#include "rows.cuh"

struct __align__(16) Pref {
	double sum;
	long long cnt;
};

#define MG_ST_LEVELS 12 // sparse-table levels: windows of up to 2^12 groups = 32768 bins without touching the pyramid
#define MG_PYR_MAX 14 // fan-out 8 levels above the raw bins: enough for 2^45 bins

struct DenseChrom {
	const float *vals;
	const Pref *p8;  // per group of 8 bins: totals over vals[0..8g), g = 0..ceil(nbins / 8)
	int64_t nbins;
	double abs_sum; // sum |v| over the chromosome: scale of the prefix rounding error (cancellation guard)
	const double *pref2; // pref2[i] = sum of squares of the non-NaN vals[0..i) (stddev)
	double sq_sum;       // sum v^2 over the chromosome
	// index structures (dense_index.cuh)
	const float *pmax[MG_PYR_MAX];
	const float *pmin[MG_PYR_MAX];
	// sparse tables over the groups of 8: smax[j][g] = max of groups [g, g + 2^j) (j = 0 is pmax[0]): the extremum of any run of
	// fewer than 2^MG_ST_LEVELS full groups is two lookups
	const float *smax[MG_ST_LEVELS];
	const float *smin[MG_ST_LEVELS];
	int64_t st_pitch; // smax[j] = smax[0] + j * st_pitch for every level that exists (one slab per extremum)
	int64_t ngroups;
	const uint2 *wm;          // wavelet matrix over RANK keys, wm_levels x wm_blocks units {ones before, 32 bits}
	const uint32_t *wm_zeros; // zeros per level
	int64_t wm_blocks;
	int64_t wm_m;             // number of non-NaN bins
	int wm_levels;            // ceil(log2(wm_m)): rank keys are balanced, every level halves a range
	const float *wm_sorted;   // the non-NaN values in ascending order: value of rank r
	const uint32_t *wm_ckpt[7]; // rank keys in the order ENTERING levels 4, 8, ..., 28: a walk whose range shrank to <= 8 finishes here
	const uint8_t *bw;        // block-local wavelet matrices (dense_blockwm.cuh), one record per half-overlapping block
	// lse pyramid (dense_index.cuh): plse[j][i] = log-sum-exp (double) of the non-NaN bins [8^(j+1) i, 8^(j+1) (i + 1)), -inf when none
	const double *plse[MG_PYR_MAX];
};

// Index structures beyond the core {vals, p8} are built on first use, one GROUP at a time for every staged chromosome of the
// track, and may be dropped again under the context's HBM budget (mg_set_option "hbm_budget_mb"): an avg-only virtual track holds
// 6 bytes per bin, not 63.
enum DenseGroupId {
	MG_G_SD = 0, // pref2: per-bin prefix of squares (stddev)                                        8 B/bin
	MG_G_EXT,    // min / max pyramids + sparse tables over the groups of 8 (min, max, min.pos, max.pos) ~12 B/bin
	MG_G_LSE,    // lse pyramid                                                                      ~1.2 B/bin
	MG_G_BW,     // block-local wavelet matrices (quantiles of windows <= MG_BW_S bins)              ~7.2 B/bin
	MG_G_WM,     // chromosome-wide wavelet matrix (6) + 5 rank checkpoints (20) + sorted values (4): longer windows ~30 B/bin
	MG_G_COUNT
};
struct DenseGroup {
	std::vector<void *> allocs;
	int64_t bytes = 0;
	bool built = false;
	int64_t last_use = 0;
};

struct mg_dense {
	uint32_t bin_size = 0;
	int nchroms = 0;
	std::vector<DenseChrom> h_table;
	DenseChrom *d_table = nullptr;
	std::vector<void *> allocs; // the core: table, vals, p8
	int64_t bytes = 0;          // everything held, groups included
	DenseGroup group[MG_G_COUNT];
};

#define MG_MAX_Q 8
#define MG_INL_CHROMS 32

struct DenseQuery {
	RowSrc rows;
	int64_t first, count;
	int64_t sshift, eshift;
	const DenseChrom *table;
	const int64_t *chrom_sizes;
	uint32_t bin_size;
	double *out[MG_NUM_FUNCS]; // one column per function (null = not requested); MG_QUANTILE unused here
	uint64_t bin_magic; // ceil(2^64 / bin_size): floor(x / bin_size) == umul64hi(x, bin_magic) for x < 2^32 (0 when bin_size == 1)
	uint32_t bin_magic32, bin_shift32; // floor(x / bin_size) == __umulhi(x, bin_magic32) >> bin_shift32 for x < 2^31 (bin_size > 1)
	uint32_t pos_rel; // bit f set: position function f is reported relative to the window start
	uint32_t out_mask; // bit f set: out[f] is requested
	uint32_t generic_mask; // bit f set: function f follows the reference's generic read_interval path (MG_FUNC_GENERIC)
	int nq;
	int prefetch;
	int use_bw; // quantiles of windows <= MG_BW_S bins from the block-local matrices
	int stage;  // k_dense_rowquery stages the block records its rows need in shared memory (0: walk them through L1; ablation)
	double q_pct[MG_MAX_Q];
	double *q_out[MG_MAX_Q];
	// The chromosome table in the kernel parameters (genomes of <= MG_INL_CHROMS chromosomes shorter than 2^31 bp): sizes, bin counts and the
	// array bases a tile's thread block needs BEFORE it can issue its bulk copies come from the constant bank instead of two dependent global
	// round trips (chrom_sizes[chrom], then table[chrom]) in every thread block's life. inl_n = 0: not filled, the kernels load the tables.
	int inl_n;
	uint32_t inl_csize[MG_INL_CHROMS], inl_nbins[MG_INL_CHROMS];
	const float *inl_vals[MG_INL_CHROMS];
	const Pref *inl_p8[MG_INL_CHROMS];
	const uint8_t *inl_bw[MG_INL_CHROMS];
	const float *inl_smax[MG_INL_CHROMS], *inl_smin[MG_INL_CHROMS]; // sparse-table slabs (null until the min / max group exists)
	uint32_t inl_stpitch[MG_INL_CHROMS];
};
static_assert(sizeof(DenseQuery) <= 4096, "kernel parameters");

// ------------------------------------------------------------------------------------------------------------
// Staging. Tile = 256 threads x 8 bins.
#define MG_ST_THREADS 256
#define MG_ST_PER_THREAD 8
#define MG_ST_TILE (MG_ST_THREADS * MG_ST_PER_THREAD)

template <typename T>
__device__ __forceinline__ T mg_block_excl_scan(T v, T *smem /* >= 32 entries */, T &block_total)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
	T incl = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		T o = __shfl_up_sync(0xffffffffu, incl, d);
		if (lane >= d)
			incl += o;
	}
	if (lane == 31)
		smem[warp] = incl;
	__syncthreads();
	if (warp == 0) {
		T w = lane < nwarps ? smem[lane] : T(0);
		T wi = w;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			T o = __shfl_up_sync(0xffffffffu, wi, d);
			if (lane >= d)
				wi += o;
		}
		smem[lane] = wi - w; // exclusive warp offsets
		if (lane == 31)
			smem[32] = wi; // grand total (nwarps <= 32, upper lanes carry zeros)
	}
	__syncthreads();
	T res = smem[warp] + (incl - v);
	block_total = smem[32];
	__syncthreads();
	return res;
}

// pass 1: +-inf -> NaN in place, per-tile {sum, count, abs sum}
__global__ void __launch_bounds__(MG_ST_THREADS) k_dense_clean_reduce(float *vals, int64_t nbins, double *tile_sum, long long *tile_cnt,
																	   double *tile_abs, double *tile_sq)
{
	__shared__ double sd[33];
	__shared__ long long sc[33];
	const int64_t base = (int64_t)blockIdx.x * MG_ST_TILE + (int64_t)threadIdx.x * MG_ST_PER_THREAD;
	double s = 0, a = 0, q = 0;
	long long c = 0;
#pragma unroll
	for (int j = 0; j < MG_ST_PER_THREAD; ++j) {
		int64_t i = base + j;
		if (i < nbins) {
			float v = vals[i];
			if (isinf(v)) {
				v = mg_nanf();
				vals[i] = v;
			}
			if (!isnan(v)) {
				s += (double)v;
				a += fabs((double)v);
				q = fma((double)v, (double)v, q);
				++c;
			}
		}
	}
	double ts, ta, tq;
	long long tc;
	mg_block_excl_scan<double>(s, sd, ts);
	mg_block_excl_scan<double>(a, sd, ta);
	mg_block_excl_scan<double>(q, sd, tq);
	mg_block_excl_scan<long long>(c, sc, tc);
	if (threadIdx.x == 0) {
		tile_sum[blockIdx.x] = ts;
		tile_cnt[blockIdx.x] = tc;
		tile_abs[blockIdx.x] = ta;
		tile_sq[blockIdx.x] = tq;
	}
}

// pass 2 (one block): exclusive scan of the tile totals in place; total abs sum -> *abs_total
__global__ void __launch_bounds__(1024) k_dense_tile_scan(double *tile_sum, long long *tile_cnt, const double *tile_abs, double *tile_sq,
														  int64_t ntiles, double *abs_total /* [0] = sum |v|, [1] = sum v^2 */)
{
	__shared__ double sd[33];
	__shared__ long long sc[33];
	const int64_t chunk = (ntiles + blockDim.x - 1) / blockDim.x;
	const int64_t lo = (int64_t)threadIdx.x * chunk, hi = min(lo + chunk, ntiles);
	double s = 0, a = 0, q = 0;
	long long c = 0;
	for (int64_t i = lo; i < hi; ++i) {
		s += tile_sum[i];
		c += tile_cnt[i];
		a += tile_abs[i];
		q += tile_sq[i];
	}
	double ts, ta, tq;
	long long tc;
	double s0 = mg_block_excl_scan<double>(s, sd, ts);
	mg_block_excl_scan<double>(a, sd, ta);
	double q0 = mg_block_excl_scan<double>(q, sd, tq);
	long long c0 = mg_block_excl_scan<long long>(c, sc, tc);
	for (int64_t i = lo; i < hi; ++i) {
		double vs = tile_sum[i], vq = tile_sq[i];
		long long vc = tile_cnt[i];
		tile_sum[i] = s0;
		tile_cnt[i] = c0;
		tile_sq[i] = q0;
		s0 += vs;
		c0 += vc;
		q0 += vq;
	}
	if (threadIdx.x == 0) {
		abs_total[0] = ta;
		abs_total[1] = tq;
	}
}

// pass 3: per-bin prefix entries. pref[i] = totals over vals[0..i); pref[nbins] is the chromosome total.
__global__ void __launch_bounds__(MG_ST_THREADS) k_dense_prefix_write(const float *vals, int64_t nbins, const double *tile_sum,
																	   const long long *tile_cnt, const double *tile_sq, Pref *pref, double *pref2)
{
	__shared__ double sd[33];
	__shared__ long long sc[33];
	const int64_t base = (int64_t)blockIdx.x * MG_ST_TILE + (int64_t)threadIdx.x * MG_ST_PER_THREAD;
	float v[MG_ST_PER_THREAD];
	double s = 0, q = 0;
	long long c = 0;
#pragma unroll
	for (int j = 0; j < MG_ST_PER_THREAD; ++j) {
		int64_t i = base + j;
		v[j] = i < nbins ? vals[i] : mg_nanf();
		if (!isnan(v[j])) {
			s += (double)v[j];
			q = fma((double)v[j], (double)v[j], q);
			++c;
		}
	}
	double ts, tq;
	long long tc;
	double s0 = tile_sum[blockIdx.x] + mg_block_excl_scan<double>(s, sd, ts);
	double q0 = tile_sq[blockIdx.x] + mg_block_excl_scan<double>(q, sd, tq);
	long long c0 = tile_cnt[blockIdx.x] + mg_block_excl_scan<long long>(c, sc, tc);
#pragma unroll
	for (int j = 0; j < MG_ST_PER_THREAD; ++j) {
		int64_t i = base + j;
		if (i <= nbins) {
			Pref p;
			p.sum = s0;
			p.cnt = c0;
			if (pref)
				pref[i] = p;
			if (pref2)
				pref2[i] = q0;
		}
		if (!isnan(v[j])) {
			s0 += (double)v[j];
			q0 = fma((double)v[j], (double)v[j], q0);
			++c0;
		}
	}
}

// ------------------------------------------------------------------------------------------------------------
// Window geometry shared by the row kernels. false => every function is NaN for this row.
struct DenseWin {
	const DenseChrom *ch;
	int64_t sbin, ebin;
	int64_t ws;    // the window's start coordinate
	bool one_bin;  // the window lies inside one bin (before the clamp to the track's end): the reference's single-bin assignment
};

__device__ __forceinline__ bool mg_dense_win(const DenseQuery &q, int64_t row, DenseWin &w)
{
	int chrom;
	int64_t s, e, scope;
	mg_get_row(q.rows, row, chrom, s, e, scope);
	int64_t ws, we;
	if (!mg_window(s, e, q.sshift, q.eshift, __ldg(q.chrom_sizes + chrom), ws, we))
		return false;
	w.ch = q.table + chrom;
	w.ws = ws;
	int64_t eb;
	if ((uint64_t)we <= 0xffffffffull - q.bin_size && q.bin_magic) { // 0 <= ws < we: both fit the magic-number division
		// both operands fit 32 bits here: say so, the product is then two 32 x 32 multiplies instead of four
		w.sbin = (int64_t)__umul64hi((uint64_t)(uint32_t)ws, q.bin_magic);                    // src/GenomeTrackFixedBin.cpp:675
		eb = (int64_t)__umul64hi((uint64_t)((uint32_t)we + q.bin_size - 1u), q.bin_magic); // :676
	} else {
		w.sbin = mg_floordiv(ws, q.bin_size);
		eb = mg_ceildiv(we, q.bin_size);
	}
	int64_t nb = w.ch->nbins;
	w.one_bin = eb == w.sbin + 1;
	w.ebin = eb < nb ? eb : nb;                           // read_bins_bulk clamp :1036-1041
	return true;
}

// ------------------------------------------------------------------------------------------------------------
// p8 entries from the per-bin prefix the staging pass produces (which is then released). The count word's top 16 bits say which bins
// hold no value in the entry's own group of 8 (bits 48-55: bin 8 g + j -> bit j) and in the group before it (bits 56-63): a window's two
// partial groups are counted from the two entries it loads anyway, before any bin has arrived.
// (Measured and dropped: entries per 16 bins for k_dense_thin -- half the prefix bytes, 0.155 GB of the C3 step's 1.97, but 16-bin partial
// groups: 0.386 -> 0.409 ms. The step pays for instructions as it pays for bytes.)
#define MG_P8_CNT_BITS 48
__device__ __forceinline__ uint32_t mg_nan_bits8(const float *__restrict__ v)
{
	const float4 a = *reinterpret_cast<const float4 *>(v), b = *reinterpret_cast<const float4 *>(v + 4);
	return (a.x != a.x ? 1u : 0u) | (a.y != a.y ? 2u : 0u) | (a.z != a.z ? 4u : 0u) | (a.w != a.w ? 8u : 0u) | (b.x != b.x ? 16u : 0u) |
		   (b.y != b.y ? 32u : 0u) | (b.z != b.z ? 64u : 0u) | (b.w != b.w ? 128u : 0u);
}
__global__ void __launch_bounds__(256) k_p8_build(const Pref *__restrict__ pref, const float *__restrict__ vals /* padded with a NaN group */, int64_t nbins,
												  Pref *__restrict__ p8, int64_t nentries)
{
	const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (g < nentries) {
		Pref p = pref[min(g * 8, nbins)];
		const unsigned long long cur = mg_nan_bits8(vals + 8 * g), prev = g > 0 ? mg_nan_bits8(vals + 8 * (g - 1)) : 0xffull;
		p.cnt = (long long)((unsigned long long)p.cnt | (cur << MG_P8_CNT_BITS) | (prev << (MG_P8_CNT_BITS + 8)));
		p8[g] = p;
	}
}

#define MG_PREFIX_DIRECT_BINS 4

__device__ __forceinline__ Pref mg_ld_pref(const Pref *p)
{
	const int4 r = __ldg(reinterpret_cast<const int4 *>(p)); // one 16-byte read-only load
	Pref o;
	o.sum = __hiloint2double(r.y, r.x);
	o.cnt = (long long)(((unsigned long long)((unsigned)r.w & 0xffffu) << 32) | (unsigned)r.z);
	return o;
}
// bins [lo, hi) of the aligned group of 8 starting at bin 8 g: one 32-byte sector. Sum in bin order (exact for the few
// floats involved), non-NaN count, extrema.
// x when it is a number and bit J of sel is set, else 0; counts it in n. Inline PTX: left to the compiler, the select moves behind
// the conversion to double and becomes two selects on the halves.
template <uint32_t J>
__device__ __forceinline__ float mg_take(float x, uint32_t sel, int &n)
{
	float c;
	uint32_t b;
	asm("{\n\t.reg .pred p;\n\t.reg .u32 t;\n\tand.b32 t, %3, %4;\n\tsetp.ne.u32 p, t, 0;\n\tsetp.num.and.f32 p, %2, %2, p;\n\tselp.f32 %0, %2, 0f00000000, p;\n\t"
		"selp.u32 %1, 1, 0, p;\n\t}"
		: "=f"(c), "=r"(b)
		: "f"(x), "r"(sel), "n"(1u << J));
	n += (int)b;
	return c;
}

template <bool SUM, bool MM>
__device__ __forceinline__ void mg_part8(const float *__restrict__ vals, int64_t g, int lo, int hi, double &S, int &n, float &mx, float &mn)
{
	const float4 a = __ldg(reinterpret_cast<const float4 *>(vals + 8 * g));
	const float4 b = __ldg(reinterpret_cast<const float4 *>(vals + 8 * g + 4));
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

// Sum / count / extrema of bins [sbin, ebin), sbin < ebin. The full groups [gs, ge) in between are NOT covered by mx / mn
// (the caller continues in the pyramid from level 1 with that range); S and n cover the whole window.
struct WinBasic {
	double S;
	long long n;
	long long before; // non-NaN bins of the chromosome before sbin (the window's start in the compacted order)
	float mx, mn;     // over the partial head / tail groups only
	int64_t gs, ge;   // full groups of the window; gs >= ge when there are none
};

template <bool SUM, bool MM, bool CNT = true>
__device__ __forceinline__ void mg_win_basic(const DenseChrom *ch, int64_t sbin, int64_t ebin, WinBasic &w)
{
	const float *vals = ch->vals;
	double S = 0;
	int n = 0;
	float mx = __int_as_float(0xff800000), mn = __int_as_float(0x7f800000);
	const int64_t g0 = sbin >> 3, g1 = (ebin - 1) >> 3;
	if (g0 == g1) { // inside one group
		mg_part8<SUM, MM>(vals, g0, (int)(sbin - 8 * g0), (int)(ebin - 8 * g0), S, n, mx, mn);
		w.S = S;
		w.n = n;
		w.before = -1; // not known without a p8 entry: mg_win_before()
		w.gs = 1;
		w.ge = 0;
	} else {
		// head group g0 from sbin on, tail group g1 up to ebin, full groups (g0, g1) in between: six independent loads
		// (2 x 2 x 16 bytes of bins + two p8 entries) issued back to back, one memory round trip
		Pref a, b;
		a.sum = b.sum = 0;
		a.cnt = b.cnt = 0;
		if (SUM || CNT) {
			a = mg_ld_pref(ch->p8 + g0 + 1);
			b = mg_ld_pref(ch->p8 + g1);
		}
		int nh = 0, nt = 0;
		double St = 0;
		mg_part8<SUM, MM>(vals, g0, (int)(sbin & 7), 8, S, nh, mx, mn);
		mg_part8<SUM, MM>(vals, g1, 0, (int)(ebin - 8 * g1), St, nt, mx, mn);
		w.S = (S + (b.sum - a.sum)) + St;
		w.n = (long long)nh + (b.cnt - a.cnt) + nt;
		w.before = a.cnt - nh;
		w.gs = g0 + 1;
		w.ge = g1;
	}
	w.mx = mx;
	w.mn = mn;
}

// non-NaN bins before sbin for a window that lies inside one group (the chromosome-wide wavelet matrix needs it)
__device__ __forceinline__ long long mg_win_before(const DenseChrom *ch, int64_t sbin)
{
	const int64_t g0 = sbin >> 3;
	double S = 0;
	int n = 0;
	float mx = 0, mn = 0;
	mg_part8<false, false>(ch->vals, g0, 0, (int)(sbin - 8 * g0), S, n, mx, mn);
	return mg_ld_pref(ch->p8 + g0).cnt + n;
}

// sum of the window in bin order, the reference's arithmetic (src/GenomeTrackFixedBin.cpp:476-591)
__device__ __forceinline__ double mg_win_direct_sum(const DenseChrom *ch, int64_t sbin, int64_t ebin)
{
	double S = 0;
	for (int64_t t = sbin; t < ebin; ++t) {
		const float v = __ldg(ch->vals + t);
		if (!isnan(v))
			S += (double)v;
	}
	return S;
}

// avg / sum / nearest / size / exists of one row. Windows inside two groups are summed in bin order (bit-exact); a
// window whose sum is small against the accumulated rounding scale of the prefix (catastrophic cancellation) is re-summed
// directly.
struct RowSum {
	double avg, sum, size, exists;
};

__device__ __forceinline__ void mg_row_sum(const DenseChrom *ch, const WinBasic &w, int64_t sbin, int64_t ebin, RowSum &r)
{
	double S = w.S;
	// rounding of each prefix entry is O(2^-53 * abs_sum); demand >= ~2^-27 relative accuracy of S
	if (w.gs < w.ge && w.n > 0 && fabs(S) < ch->abs_sum * 1.4901161193847656e-08 /* 2^-26 */)
		S = mg_win_direct_sum(ch, sbin, ebin);
	r.size = (double)w.n;
	r.exists = w.n > 0 ? 1. : 0.;
	r.avg = r.sum = mg_nan();
	if (w.n > 0) {
		r.avg = mg_f2d(S / (double)w.n);
		r.sum = mg_f2d(S);
	}
}

// ------------------------------------------------------------------------------------------------------------
// k_dense_prefix: one thread per row; avg / sum / nearest / size / exists.
__global__ void __launch_bounds__(256) k_dense_prefix(const DenseQuery q)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= q.count)
		return;
	DenseWin w;
	RowSum r;
	r.avg = r.sum = r.size = r.exists = mg_nan();
	if (mg_dense_win(q, q.first + i, w)) {
		r.size = 0;
		r.exists = 0;
		if (w.ebin > w.sbin) {
			WinBasic wb;
			mg_win_basic<true, false>(w.ch, w.sbin, w.ebin, wb);
			mg_row_sum(w.ch, wb, w.sbin, w.ebin, r);
		}
	}
	if (q.out[MG_AVG]) mg_st_stream(q.out[MG_AVG] + i, r.avg);
	if (q.out[MG_NEAREST]) mg_st_stream(q.out[MG_NEAREST] + i, r.avg);
	if (q.out[MG_SUM]) mg_st_stream(q.out[MG_SUM] + i, r.sum);
	if (q.out[MG_SIZE]) mg_st_stream(q.out[MG_SIZE] + i, r.size);
	if (q.out[MG_EXISTS]) mg_st_stream(q.out[MG_EXISTS] + i, r.exists);
}

// ------------------------------------------------------------------------------------------------------------
// k_dense_sweep: one warp per row, any function set. Lanes stride the window (coalesced 128-byte requests);
// count / sum / min / max / Welford (mean, M2) are merged with shuffles. For quantiles the non-NaN values are
// compacted into a per-warp shared-memory buffer and sorted there (bitonic); windows that do not fit are
// selected by an 8-bit radix select that re-reads the window from L2/HBM.
#define MG_SW_WARPS 8
#define MG_SW_CAP 1024 // floats per warp buffer (4 KB); covers +-10 kb at 20 bp

__device__ __forceinline__ uint32_t mg_fkey(float v)
{
	uint32_t b = __float_as_uint(v);
	return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float mg_fkey_inv(uint32_t k)
{
	uint32_t b = k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu);
	return __uint_as_float(b);
}

// bitonic sort of buf[0..P) (P power of two, padded with +inf) by one warp
__device__ __forceinline__ void mg_warp_bitonic(float *buf, int P, int lane)
{
	for (int k = 2; k <= P; k <<= 1) {
		for (int j = k >> 1; j > 0; j >>= 1) {
			for (int t = lane; t < (P >> 1); t += 32) {
				int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
				int ixj = i | j;
				float a = buf[i], b = buf[ixj];
				bool up = (i & k) == 0;
				if ((a > b) == up) {
					buf[i] = b;
					buf[ixj] = a;
				}
			}
			__syncwarp();
		}
	}
}

// k-th smallest (0-based) non-NaN value of vals[sbin..ebin) by MSD radix select, 8 bits per pass; hist = 256 uint32 in smem.
__device__ float mg_warp_radix_select(const float *vals, int64_t sbin, int64_t ebin, long long k, uint32_t *hist, int lane)
{
	uint32_t prefix = 0, mask = 0;
	for (int shift = 24; shift >= 0; shift -= 8) {
		for (int b = lane; b < 256; b += 32)
			hist[b] = 0;
		__syncwarp();
		for (int64_t i = sbin + lane; i < ebin; i += 32) {
			float v = __ldg(vals + i);
			if (!isnan(v)) {
				uint32_t key = mg_fkey(v);
				if ((key & mask) == prefix)
					atomicAdd(&hist[(key >> shift) & 0xff], 1u);
			}
		}
		__syncwarp();
		// find the digit whose cumulative count passes k (serial over 256 bins on all lanes; rare path)
		uint32_t digit = 0;
		long long acc = 0;
		for (int b = 0; b < 256; ++b) {
			uint32_t c = hist[b];
			if (acc + c > k) {
				digit = b;
				break;
			}
			acc += c;
		}
		k -= acc;
		prefix |= digit << shift;
		mask |= 0xffu << shift;
		__syncwarp();
	}
	return mg_fkey_inv(prefix);
}

// src/StreamPercentiler.h:210-217, operands kept un-fused so the double rounding matches the host compiler's
__device__ __forceinline__ double mg_interp(float x1, float x2, double wgt)
{
	return mg_f2d(__dadd_rn(__dmul_rn((double)x1, __dsub_rn(1., wgt)), __dmul_rn((double)x2, wgt)));
}

// Specialised on what the requested function set needs, so that e.g. a max-only query does not pay for the
// double-precision Welford update: SUM (avg/sum/nearest), MM (min/max), SD (stddev), Q (quantiles).
template <bool SUM, bool MM, bool SD, bool Q>
__global__ void __launch_bounds__(MG_SW_WARPS * 32) k_dense_sweep(const DenseQuery q)
{
	__shared__ float sbuf[MG_SW_WARPS][MG_SW_CAP];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	float *buf = sbuf[warp];
	const int64_t nwarps_total = (int64_t)gridDim.x * MG_SW_WARPS;
	for (int64_t i = (int64_t)blockIdx.x * MG_SW_WARPS + warp; i < q.count; i += nwarps_total) {
		DenseWin w;
		const bool ok = mg_dense_win(q, q.first + i, w);
		long long n = 0;
		double S = 0, mean = 0, m2 = 0;
		float fmin = 3.402823466e+38f, fmax = -3.402823466e+38f;
		const bool want_q = Q;
		bool in_smem = true;
		if (ok) {
			const float *vals = w.ch->vals;
			for (int64_t b0 = w.sbin; b0 < w.ebin; b0 += 32) {
				const int64_t b = b0 + lane;
				float v = b < w.ebin ? __ldg(vals + b) : mg_nanf();
				const bool valid = !isnan(v);
				if (valid) {
					++n;
					if (SUM)
						S += (double)v;
					if (MM) {
						fmin = fminf(fmin, v);
						fmax = fmaxf(fmax, v);
					}
					if (SD) {
						const double d = (double)v - mean;
						mean += d / (double)n;
						m2 += d * ((double)v - mean);
					}
				}
			}
		}
		// merge lanes: counts, sums, extrema, Welford pairs (Chan et al.)
#pragma unroll
		for (int m = 16; m > 0; m >>= 1) {
			const long long on = __shfl_xor_sync(0xffffffffu, n, m);
			const long long tn = n + on;
			if (SD) {
				const double omean = __shfl_xor_sync(0xffffffffu, mean, m);
				const double om2 = __shfl_xor_sync(0xffffffffu, m2, m);
				if (tn > 0) {
					// symmetric form so both partners compute identical results
					const double d = omean - mean;
					const double nm = ((double)n * mean + (double)on * omean) / (double)tn;
					m2 = m2 + om2 + d * d * ((double)n * (double)on / (double)tn);
					mean = nm;
				}
			}
			n = tn;
			if (SUM)
				S += __shfl_xor_sync(0xffffffffu, S, m);
			if (MM) {
				fmin = fminf(fmin, __shfl_xor_sync(0xffffffffu, fmin, m));
				fmax = fmaxf(fmax, __shfl_xor_sync(0xffffffffu, fmax, m));
			}
		}
		// quantiles
		double qv[MG_MAX_Q];
#pragma unroll
		for (int k = 0; k < MG_MAX_Q; ++k)
			qv[k] = mg_nan();
		if (want_q && ok && n > 0) {
			in_smem = n <= MG_SW_CAP;
			if (in_smem) {
				// compact non-NaN values into the warp buffer (second pass over the window: L1/L2 hits)
				const float *vals = w.ch->vals;
				int fill = 0;
				for (int64_t b0 = w.sbin; b0 < w.ebin; b0 += 32) {
					const int64_t b = b0 + lane;
					float v = b < w.ebin ? __ldg(vals + b) : mg_nanf();
					const bool valid = !isnan(v);
					const unsigned m = __ballot_sync(0xffffffffu, valid);
					if (valid)
						buf[fill + __popc(m & ((1u << lane) - 1))] = v;
					fill += __popc(m);
				}
				int P = 1;
				while (P < (int)n)
					P <<= 1;
				for (int t = (int)n + lane; t < P; t += 32)
					buf[t] = __int_as_float(0x7f800000);
				__syncwarp();
				mg_warp_bitonic(buf, P, lane);
			}
			for (int k = 0; k < q.nq; ++k) {
				double pct = q.q_pct[k];
				pct = pct < 0. ? 0. : (pct > 1. ? 1. : pct);
				const double index = __dmul_rn((double)(n - 1), pct);
				const long long i1 = (long long)floor(index), i2 = (long long)ceil(index);
				const double wgt = __dsub_rn(index, (double)i1);
				float x1, x2;
				if (in_smem) {
					x1 = buf[i1];
					x2 = buf[i2];
				} else {
					uint32_t *hist = reinterpret_cast<uint32_t *>(buf);
					x1 = mg_warp_radix_select(w.ch->vals, w.sbin, w.ebin, i1, hist, lane);
					x2 = i2 == i1 ? x1 : mg_warp_radix_select(w.ch->vals, w.sbin, w.ebin, i2, hist, lane);
				}
				qv[k] = mg_interp(x1, x2, wgt);
			}
			__syncwarp();
		}
		if (lane == 0) {
			double avg = mg_nan(), sum = mg_nan(), mn = mg_nan(), mx = mg_nan(), sd = mg_nan(), size = mg_nan(), exists = mg_nan();
			if (ok) {
				size = (double)n;
				exists = n > 0 ? 1. : 0.;
				if (n > 0) {
					avg = mg_f2d(S / (double)n);
					sum = mg_f2d(S);
					mn = (double)fmin;
					mx = (double)fmax;
					if (n > 1)
						sd = mg_f2d(sqrt(m2 / (double)(n - 1)));
				}
			}
			if (q.out[MG_AVG]) mg_st_stream(q.out[MG_AVG] + i, avg);
			if (q.out[MG_NEAREST]) mg_st_stream(q.out[MG_NEAREST] + i, avg);
			if (q.out[MG_SUM]) mg_st_stream(q.out[MG_SUM] + i, sum);
			if (q.out[MG_MIN]) mg_st_stream(q.out[MG_MIN] + i, mn);
			if (q.out[MG_MAX]) mg_st_stream(q.out[MG_MAX] + i, mx);
			if (q.out[MG_STDDEV]) mg_st_stream(q.out[MG_STDDEV] + i, sd);
			if (q.out[MG_SIZE]) mg_st_stream(q.out[MG_SIZE] + i, size);
			if (q.out[MG_EXISTS]) mg_st_stream(q.out[MG_EXISTS] + i, exists);
			for (int k = 0; k < q.nq; ++k)
				mg_st_stream(q.q_out[k] + i, qv[k]);
		}
	}
}

// ------------------------------------------------------------------------------------------------------------
// k_dense_ext: one warp per row; the order-dependent functions of the reference's generic path
// (src/GenomeTrackFixedBin.cpp:881-939): first / last (+ position), min.pos / max.pos, lse.
//   first / last  : ballots from either end of the window stop at the first group of 32 bins that holds a value;
//   min / max pos : every lane keeps (value, bin) of its stride with the reference's strict comparisons, lanes merge on
//                   (value, then smaller bin): the first bin holding the extremum, as :886-905 leave it;
//   lse           : M + log(sum exp(v - M)) in double with M the window's maximum, rounded to float -- the value the
//                   reference's reducer path produces (RunningLogSumExp::value, src/utils/RunningLogSumExp.h:32-45,175-178);
//                   its generic path accumulates pairwise in float (src/GenomeTrack1D.h:20-42) and agrees to ~1e-6.
// Position of bin b: max(b * bin_size, window start) (:883-884), as a double.
// k_dense_firstlast: first / last (+ positions) alone, one THREAD per row: the first value of a window is nearly always in its
// first bins (a warp per row spends a whole warp on one load: 2.3 ms against 0.2 ms for the 10 M rows of C3); a lane inside a long
// NaN run just keeps walking.
__global__ void __launch_bounds__(256) k_dense_firstlast(const DenseQuery q)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= q.count)
		return;
	DenseWin w;
	const bool ok = mg_dense_win(q, q.first + i, w);
	int64_t fb = -1, lb = -1;
	float fv = mg_nanf(), lv = mg_nanf();
	if (ok && w.ebin > w.sbin) {
		const float *vals = w.ch->vals;
		if (w.one_bin) { // assign_single_bin_value: the bin's position whatever its value (see k_dense_ext)
			fv = lv = __ldg(vals + w.sbin);
			fb = lb = w.sbin;
		} else {
			if (q.out[MG_FIRST] || q.out[MG_FIRST_POS])
				for (int64_t b = w.sbin; b < w.ebin; ++b) {
					const float v = __ldg(vals + b);
					if (v == v) {
						fv = v;
						fb = b;
						break;
					}
				}
			if (q.out[MG_LAST] || q.out[MG_LAST_POS])
				for (int64_t b = w.ebin - 1; b >= w.sbin; --b) {
					const float v = __ldg(vals + b);
					if (v == v) {
						lv = v;
						lb = b;
						break;
					}
				}
		}
	}
	const double bs = (double)q.bin_size, ws = ok ? (double)w.ws : 0.;
	if (q.out[MG_FIRST]) mg_st_stream(q.out[MG_FIRST] + i, (double)fv);
	if (q.out[MG_LAST]) mg_st_stream(q.out[MG_LAST] + i, (double)lv);
	if (q.out[MG_FIRST_POS]) {
		double p = mg_nan();
		if (fb >= 0) {
			p = fmax((double)fb * bs, ws);
			if (q.pos_rel & (1u << MG_FIRST_POS))
				p -= ws;
		}
		mg_st_stream(q.out[MG_FIRST_POS] + i, p);
	}
	if (q.out[MG_LAST_POS]) {
		double p = mg_nan();
		if (lb >= 0) {
			p = fmax((double)lb * bs, ws);
			if (q.pos_rel & (1u << MG_LAST_POS))
				p -= ws;
		}
		mg_st_stream(q.out[MG_LAST_POS] + i, p);
	}
}

#define MG_EXT_WARPS 8
template <bool SCAN>
__global__ void __launch_bounds__(MG_EXT_WARPS * 32) k_dense_ext(const DenseQuery q)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int64_t nwarps_total = (int64_t)gridDim.x * MG_EXT_WARPS;
	for (int64_t i = (int64_t)blockIdx.x * MG_EXT_WARPS + warp; i < q.count; i += nwarps_total) {
		DenseWin w;
		const bool ok = mg_dense_win(q, q.first + i, w);
		const float *vals = ok ? w.ch->vals : nullptr;
		const int64_t sbin = ok ? w.sbin : 0, ebin = ok ? w.ebin : 0;
		int64_t fb = -1, lb = -1, mnb = -1, mxb = -1;
		float fv = mg_nanf(), lv = mg_nanf();
		double lse = mg_nan();
		if (q.out[MG_FIRST] || q.out[MG_FIRST_POS]) {
			for (int64_t b0 = sbin; b0 < ebin; b0 += 32) {
				const int64_t b = b0 + lane;
				const float v = b < ebin ? __ldg(vals + b) : mg_nanf();
				const unsigned m = __ballot_sync(0xffffffffu, v == v);
				if (m) {
					const int src = __ffs(m) - 1;
					fv = __shfl_sync(0xffffffffu, v, src);
					fb = b0 + src;
					break;
				}
			}
		}
		if (q.out[MG_LAST] || q.out[MG_LAST_POS]) {
			for (int64_t b1 = ebin; b1 > sbin; b1 -= 32) {
				const int64_t b = b1 - 32 + lane;
				const float v = b >= sbin ? __ldg(vals + b) : mg_nanf();
				const unsigned m = __ballot_sync(0xffffffffu, v == v);
				if (m) {
					const int src = 31 - __clz(m);
					lv = __shfl_sync(0xffffffffu, v, src);
					lb = b1 - 32 + src;
					break;
				}
			}
		}
		if (SCAN) {
			float mnv = __int_as_float(0x7f800000), mxv = __int_as_float(0xff800000);
			for (int64_t b = sbin + lane; b < ebin; b += 32) {
				const float v = __ldg(vals + b);
				if (v == v) {
					if (v < mnv || mnb < 0) { // bins ascend within a lane: strict comparisons keep the first occurrence
						mnv = v;
						mnb = b;
					}
					if (v > mxv || mxb < 0) {
						mxv = v;
						mxb = b;
					}
				}
			}
#pragma unroll
			for (int m = 16; m > 0; m >>= 1) {
				const float ov = __shfl_xor_sync(0xffffffffu, mnv, m), oV = __shfl_xor_sync(0xffffffffu, mxv, m);
				const int64_t ob = __shfl_xor_sync(0xffffffffu, mnb, m), oB = __shfl_xor_sync(0xffffffffu, mxb, m);
				if (ob >= 0 && (mnb < 0 || ov < mnv || (ov == mnv && ob < mnb))) {
					mnv = ov;
					mnb = ob;
				}
				if (oB >= 0 && (mxb < 0 || oV > mxv || (oV == mxv && oB < mxb))) {
					mxv = oV;
					mxb = oB;
				}
			}
			if (q.out[MG_LSE] && mxb >= 0) {
				const double M = (double)mxv;
				double s = 0;
				for (int64_t b = sbin + lane; b < ebin; b += 32) {
					const float v = __ldg(vals + b);
					if (v == v)
						s += exp((double)v - M);
				}
#pragma unroll
				for (int m = 16; m > 0; m >>= 1)
					s += __shfl_xor_sync(0xffffffffu, s, m);
				lse = mg_f2d(M + log(s));
			}
		}
		if (ok && w.one_bin && sbin < ebin) {
			// assign_single_bin_value (src/GenomeTrackFixedBin.cpp:150-180): a window inside one bin reports the bin's position for
			// every position function and the bin's value for lse / first / last, whatever the value (a NaN bin keeps its position)
			fv = lv = __ldg(vals + sbin);
			lse = (double)fv;
			fb = lb = mnb = mxb = sbin;
		}
		if (lane == 0) {
			const double bs = (double)q.bin_size, ws = ok ? (double)w.ws : 0.;
#define MG_EXT_POS(F, BIN)                                                                                              \
	if (q.out[F]) {                                                                                                     \
		double p = mg_nan();                                                                                            \
		if ((BIN) >= 0) {                                                                                               \
			p = fmax((double)(BIN) * bs, ws);                                                                           \
			if (q.pos_rel & (1u << (F)))                                                                                \
				p -= ws;                                                                                                \
		}                                                                                                               \
		mg_st_stream(q.out[F] + i, p);                                                                                  \
	}
			if (q.out[MG_FIRST]) mg_st_stream(q.out[MG_FIRST] + i, (double)fv);
			if (q.out[MG_LAST]) mg_st_stream(q.out[MG_LAST] + i, (double)lv);
			MG_EXT_POS(MG_FIRST_POS, fb)
			MG_EXT_POS(MG_LAST_POS, lb)
			if (SCAN) {
				MG_EXT_POS(MG_MIN_POS, mnb)
				MG_EXT_POS(MG_MAX_POS, mxb)
				if (q.out[MG_LSE]) mg_st_stream(q.out[MG_LSE] + i, lse);
			}
#undef MG_EXT_POS
		}
	}
}
