// sparse.cuh -- sparse tracks (sorted, disjoint records) and interval sets: per-row binary-search merge.
//
// HBM layout per chromosome (SoA): start int64[n], end int64[n], val float32[n] (+-inf already NaN,
// src/GenomeTrackSparse.cpp:205-206). Rows are position-sorted, so neighbouring threads walk the same search
// path and the upper levels of the search stay in L1/L2.
#pragma once
#include "common.cuh"
#include "rows.cuh"
#include "dense.cuh" // MG_MAX_Q, mg_fkey, mg_interp

struct SparseChrom {
	const int64_t *start;
	const int64_t *end;
	const float *val;
	int64_t n;
	// coarse position index built at staging: coarse[c] = first record whose end > (c << coarse_shift), c = 0..ncoarse-1.
	// Two lookups bracket the records any window inside a block's hull can touch (no search).
	const uint32_t *coarse;
	int64_t ncoarse;
	int coarse_shift;
};

struct mg_sparse {
	int nchroms = 0;
	std::vector<SparseChrom> h_table;
	SparseChrom *d_table = nullptr;
	std::vector<void *> allocs;
	int64_t bytes = 0;
};

struct mg_ivset {
	int nchroms = 0;
	std::vector<SparseChrom> h_table; // val unused
	SparseChrom *d_table = nullptr;
	std::vector<void *> allocs;
	int64_t bytes = 0;
};

// Regular-grid thread blocks of a launch over fixed-bin rows, worked out on the host (mg_make_grid_table): per scope interval the
// range of block indices whose rows all lie strictly inside it with unclamped windows, and the window start of the first of
// them. A thread block finds its entry by a short search in the kernel parameters and derives its geometry with one
// multiply-add, instead of the scope search, the 64-bit row arithmetic and the range checks of mg_sf_grid_block.
struct GridTable {
	int n; // entries; 0 = no regular-grid block in this launch; -1 = not worked out (long scopes): the device decides per block
	int len;      // window length
	int64_t step; // rows per block x bin size
	unsigned bfirst[MG_SCOPE_INLINE], blast[MG_SCOPE_INLINE];
	int64_t origin0[MG_SCOPE_INLINE];
	int chrom[MG_SCOPE_INLINE];
};

struct SparseQuery {
	RowSrc rows;
	int64_t first, count;
	int64_t sshift, eshift;
	const SparseChrom *table;
	const int64_t *chrom_sizes;
	double *out[MG_NUM_FUNCS];
	uint32_t pos_rel; // bit f set: position function f is reported relative to the window start
	int nq;
	double q_pct[MG_MAX_Q];
	double *q_out[MG_MAX_Q];
	GridTable grid;
};

// src/GenomeTrack1D.h:20-30: the pairwise float accumulation the sparse reader's lse runs in record order
__device__ __forceinline__ void mg_lse_accumulate(float &l1, float l2)
{
	if (l1 > l2) {
		if (!isinf(l2))
			l1 += logf(1.0f + expf(l2 - l1));
	} else {
		if (isinf(l1))
			l1 = l2;
		else
			l1 = l2 + logf(1.0f + expf(l1 - l2));
	}
}

// distance between query [qs,qe) and record [rs,re): src/GInterval.cpp:33-44 with strand 0
__device__ __forceinline__ int64_t mg_dist(int64_t qs, int64_t qe, int64_t rs, int64_t re)
{
	const int64_t lo = qs > rs ? qs : rs, hi = qe < re ? qe : re;
	if (lo < hi)
		return 0;
	const int64_t l = llabs(rs - qe), r = llabs(re - qs);
	return l <= r ? l : r;
}

// k-th smallest (0-based) non-NaN value of val[r0..r1) by a 4-bit MSD radix select run by ONE thread
__device__ float mg_thread_radix_select(const float *val, int64_t r0, int64_t r1, long long k)
{
	uint32_t prefix = 0, mask = 0;
	for (int shift = 28; shift >= 0; shift -= 4) {
		unsigned long long hist[16];
#pragma unroll
		for (int b = 0; b < 16; ++b)
			hist[b] = 0;
		for (int64_t r = r0; r < r1; ++r) {
			const float v = val[r];
			if (!isnan(v)) {
				const uint32_t key = mg_fkey(v);
				if ((key & mask) == prefix)
					hist[(key >> shift) & 15]++;
			}
		}
		uint32_t digit = 15;
		long long acc = 0;
		for (int b = 0; b < 16; ++b) {
			if (acc + (long long)hist[b] > k) {
				digit = b;
				break;
			}
			acc += (long long)hist[b];
		}
		k -= acc;
		prefix |= digit << shift;
		mask |= 0xfu << shift;
	}
	return mg_fkey_inv(prefix);
}

// ------------------------------------------------------------------------------------------------------------------
// Block-level narrowing. The rows of a block are (nearly always) neighbours on one chromosome, so the records any of
// them can touch form one short contiguous range: two binary searches per BLOCK find it, the range is staged in shared
// memory with coalesced loads, and each thread then searches ~log2(range) shared-memory entries instead of
// ~log2(n) dependent global loads. Exactness: for a row window [ws, we) inside the block hull [WS, WE), the first record
// with end > ws and the last one with start < we both lie in [lo, hi] where lo = first record with end > WS and
// hi = first record with start >= WE; one extra record on each side keeps the nearest-neighbour cases local.
#define MG_SP_THREADS 256
#define MG_SP_ROWS 8   // rows per thread: one block narrows for 2048 rows
#define MG_SP_CAP 1024 // records staged per block (20 KB)

struct SpLocal {
	const int64_t *start;
	const int64_t *end;
	const float *val;
	int64_t n;
};

struct SpSmem {
	int64_t start[MG_SP_CAP];
	int64_t end[MG_SP_CAP];
	float val[MG_SP_CAP];
	int64_t red_min[MG_SP_THREADS / 32], red_max[MG_SP_THREADS / 32];
	int red_chrom_lo[MG_SP_THREADS / 32], red_chrom_hi[MG_SP_THREADS / 32];
	int64_t hull_min, hull_max;
	int64_t s_lo[2], s_hi[2]; // the two cooperative searches
	int s_first[2];
	int64_t lo, n;
	int chrom;
	int mode; // 0 = per-thread global search over the whole chromosome, 1 = staged in smem, 2 = narrowed global range
};

// Two 128-ary searches run side by side by the 256 threads of a block: half 0 finds the first record with end > WS,
// half 1 the first record with start >= WE (both predicates are monotone over the sorted, disjoint records).
// Each round is ONE global load per thread, so a block resolves n = 10^6 records in 3 rounds.
__device__ __forceinline__ void mg_sp_block_search(SpSmem &sm, const SparseChrom &ch)
{
	const int half = threadIdx.x >> 7, t = threadIdx.x & 127;
	if (t == 0) {
		sm.s_lo[half] = 0;
		sm.s_hi[half] = ch.n;
	}
	__syncthreads();
	for (;;) {
		const int64_t lo = sm.s_lo[half], hi = sm.s_hi[half];
		const int64_t span0 = sm.s_hi[0] - sm.s_lo[0], span1 = sm.s_hi[1] - sm.s_lo[1];
		const bool last_round = span0 <= 128 && span1 <= 128;
		const int64_t span = hi - lo;
		const int64_t step = span <= 128 ? 1 : (span + 127) / 128;
		const int64_t p = lo + (int64_t)(t + 1) * step - 1; // probe; the answer is the first index whose predicate holds
		bool pred = true; // a probe at or past hi is past the answer by definition (the answer is <= hi)
		if (p < hi)
			pred = half == 0 ? (__ldg(ch.end + p) > sm.hull_min) : (__ldg(ch.start + p) >= sm.hull_max);
		if (t == 0)
			sm.s_first[half] = 128;
		__syncthreads();
		if (pred)
			atomicMin(&sm.s_first[half], t);
		__syncthreads();
		const int T = sm.s_first[half];
		__syncthreads();
		if (t == 0) {
			// probes t < T are false, probe T (if any) is true: answer in (p_{T-1}, p_T]
			const int64_t nlo = T > 0 ? min(hi, lo + (int64_t)T * step) : lo;
			const int64_t nhi = T < 128 ? min(hi, lo + (int64_t)(T + 1) * step - 1) : hi;
			sm.s_lo[half] = step == 1 ? nhi : nlo; // with step == 1 the first true probe IS the answer
			sm.s_hi[half] = nhi;
		}
		__syncthreads();
		if (last_round)
			break;
	}
}

// Computes the block's record view. `ok/chrom/ws/we` describe up to MG_SP_ROWS rows of this thread.
// (mn, mx, clo, chi) = this thread's hull over its valid rows (INT64_MAX / INT64_MIN / INT_MAX / INT_MIN when none)
__device__ __forceinline__ SpLocal mg_sp_narrow(SpSmem &sm, const SparseChrom *table, int64_t mn, int64_t mx, int clo, int chi, bool want_val,
												int &mode_out)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
	for (int k = 16; k > 0; k >>= 1) {
		mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, k));
		mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, k));
		clo = min(clo, __shfl_xor_sync(0xffffffffu, clo, k));
		chi = max(chi, __shfl_xor_sync(0xffffffffu, chi, k));
	}
	if (lane == 0) {
		sm.red_min[warp] = mn;
		sm.red_max[warp] = mx;
		sm.red_chrom_lo[warp] = clo;
		sm.red_chrom_hi[warp] = chi;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		for (int w = 1; w < MG_SP_THREADS / 32; ++w) {
			mn = min(mn, sm.red_min[w]);
			mx = max(mx, sm.red_max[w]);
			clo = min(clo, sm.red_chrom_lo[w]);
			chi = max(chi, sm.red_chrom_hi[w]);
		}
		sm.hull_min = mn;
		sm.hull_max = mx;
		sm.chrom = clo == INT_MAX ? 0 : clo;
		sm.mode = (clo == chi && clo != INT_MAX) ? 1 : 0; // one chromosome for every valid row of the block?
	}
	__syncthreads();
	if (sm.mode == 0) {
		mode_out = 0;
		return SpLocal{nullptr, nullptr, nullptr, 0};
	}
	const SparseChrom ch = table[sm.chrom];
	mg_sp_block_search(sm, ch);
	if (threadIdx.x == 0) {
		const int64_t lo = sm.s_lo[0], l2 = sm.s_lo[1];
		const int64_t a = lo > 0 ? lo - 1 : 0, b = l2 < ch.n ? l2 + 1 : ch.n; // one neighbour on each side
		sm.lo = a;
		sm.n = b > a ? b - a : 0;
		sm.mode = sm.n <= MG_SP_CAP ? 1 : 2;
	}
	__syncthreads();
	mode_out = sm.mode;
	if (sm.mode == 2) // range too long for shared memory: search the narrowed global range
		return SpLocal{ch.start + sm.lo, ch.end + sm.lo, ch.val ? ch.val + sm.lo : nullptr, sm.n};
	for (int64_t t = threadIdx.x; t < sm.n; t += MG_SP_THREADS) { // stage [lo, lo + n) (coalesced)
		sm.start[t] = __ldg(ch.start + sm.lo + t);
		sm.end[t] = __ldg(ch.end + sm.lo + t);
		if (want_val)
			sm.val[t] = __ldg(ch.val + sm.lo + t);
	}
	__syncthreads();
	return SpLocal{sm.start, sm.end, sm.val, sm.n};
}

// Per-row arithmetic. read_interval: src/GenomeTrackSparse.cpp:237-340; calc_vals: src/GenomeTrackSparse.h:143-259.
// Sums and Welford updates run in record order, exactly the reference's order => avg / sum / stddev are bit-exact.
template <bool FULL> // FULL = stddev and quantiles too (the merge kernels' fallback never needs them)
__device__ __forceinline__ void mg_sparse_row(const SparseQuery &q, const SpLocal &ch, bool ok, int64_t ws, int64_t we, int64_t i)
{
	double avg = mg_nan(), sum = mg_nan(), mn = mg_nan(), mx = mg_nan(), sd = mg_nan(), nearest = mg_nan(), size = mg_nan(),
		   exists = mg_nan();
	double qv[MG_MAX_Q];
#pragma unroll
	for (int k = 0; k < MG_MAX_Q; ++k)
		qv[k] = mg_nan();
	// the order-dependent functions (src/GenomeTrackSparse.h:166-224): positions are record starts
	double lse = mg_nan(), v_first = mg_nan(), v_last = mg_nan(), first_pos = mg_nan(), last_pos = mg_nan(), min_pos = mg_nan(), max_pos = mg_nan();
	if (ok) {
		size = 0;
		exists = 0;
		if (ch.n > 0) {
			if (ch.start[0] >= we)
				nearest = (double)ch.val[0]; // before the first record (:281-284)
			else if (ch.end[ch.n - 1] <= ws)
				nearest = (double)ch.val[ch.n - 1]; // after the last record (:286-289)
			else {
				// first record whose end > ws (ends are sorted because records are disjoint)
				int64_t lo = 0, hi = ch.n;
				while (lo < hi) {
					const int64_t mid = lo + ((hi - lo) >> 1);
					if (ch.end[mid] > ws)
						hi = mid;
					else
						lo = mid + 1;
				}
				const int64_t first = lo;
				if (ch.start[first] < we) {
					long long n = 0;
					double S = 0, mean = 0, m2 = 0;
					float fmin = 3.402823466e+38f, fmax = -3.402823466e+38f;
					float flse = __int_as_float(0xff800000);
					int64_t r = first;
					for (; r < ch.n && ch.start[r] < we; ++r) {
						const float v = ch.val[r];
						if (isnan(v))
							continue;
						S += (double)v;
						if (FULL) {
							const double pos = (double)ch.start[r];
							if (v < fmin || n == 0)
								min_pos = pos; // later ties start further right: the first occurrence stays (:174-180)
							if (v > fmax)
								max_pos = pos;
							if (n == 0) {
								v_first = (double)v;
								first_pos = pos;
							}
							v_last = (double)v;
							last_pos = pos;
							if (q.out[MG_LSE])
								mg_lse_accumulate(flse, v);
						}
						if (v < fmin) fmin = v;
						if (v > fmax) fmax = v;
						++n;
						if (FULL) {
							const double d = (double)v - mean;
							mean += d / (double)n;
							m2 += d * ((double)v - mean);
						}
					}
					if (FULL && n > 0)
						lse = (double)flse;
					size = (double)n;
					if (n > 0) {
						exists = 1;
						avg = nearest = mg_f2d(S / (double)n);
						sum = mg_f2d(S);
						mn = (double)fmin;
						mx = (double)fmax;
						if (FULL && n > 1)
							sd = mg_f2d(sqrt(m2 / (double)(n - 1)));
						for (int k = 0; FULL && k < q.nq; ++k) {
							double pct = q.q_pct[k];
							pct = pct < 0. ? 0. : (pct > 1. ? 1. : pct);
							const double index = __dmul_rn((double)(n - 1), pct);
							const long long i1 = (long long)floor(index), i2 = (long long)ceil(index);
							const float x1 = mg_thread_radix_select(ch.val, first, r, i1);
							const float x2 = i2 == i1 ? x1 : mg_thread_radix_select(ch.val, first, r, i2);
							qv[k] = mg_interp(x1, x2, __dsub_rn(index, (double)i1));
						}
					}
				} else {
					// no overlap: closer neighbour, ties to the earlier record (:336-338)
					const int64_t prv = first - 1;
					const int64_t dp = mg_dist(ws, we, ch.start[prv], ch.end[prv]);
					const int64_t dn = mg_dist(ws, we, ch.start[first], ch.end[first]);
					nearest = (double)ch.val[dp <= dn ? prv : first];
				}
			}
		}
	}
	if (q.out[MG_AVG]) mg_st_stream(q.out[MG_AVG] + i, avg);
	if (q.out[MG_NEAREST]) mg_st_stream(q.out[MG_NEAREST] + i, nearest);
	if (q.out[MG_SUM]) mg_st_stream(q.out[MG_SUM] + i, sum);
	if (q.out[MG_MIN]) mg_st_stream(q.out[MG_MIN] + i, mn);
	if (q.out[MG_MAX]) mg_st_stream(q.out[MG_MAX] + i, mx);
	if (FULL && q.out[MG_STDDEV]) mg_st_stream(q.out[MG_STDDEV] + i, sd);
	if (q.out[MG_SIZE]) mg_st_stream(q.out[MG_SIZE] + i, size);
	if (q.out[MG_EXISTS]) mg_st_stream(q.out[MG_EXISTS] + i, exists);
	for (int k = 0; FULL && k < q.nq; ++k)
		mg_st_stream(q.q_out[k] + i, qv[k]);
	if (FULL) {
		const double base = (double)ws;
		if (q.out[MG_LSE]) mg_st_stream(q.out[MG_LSE] + i, lse);
		if (q.out[MG_FIRST]) mg_st_stream(q.out[MG_FIRST] + i, v_first);
		if (q.out[MG_LAST]) mg_st_stream(q.out[MG_LAST] + i, v_last);
		if (q.out[MG_FIRST_POS]) mg_st_stream(q.out[MG_FIRST_POS] + i, (q.pos_rel >> MG_FIRST_POS) & 1 ? first_pos - base : first_pos);
		if (q.out[MG_LAST_POS]) mg_st_stream(q.out[MG_LAST_POS] + i, (q.pos_rel >> MG_LAST_POS) & 1 ? last_pos - base : last_pos);
		if (q.out[MG_MIN_POS]) mg_st_stream(q.out[MG_MIN_POS] + i, (q.pos_rel >> MG_MIN_POS) & 1 ? min_pos - base : min_pos);
		if (q.out[MG_MAX_POS]) mg_st_stream(q.out[MG_MAX_POS] + i, (q.pos_rel >> MG_MAX_POS) & 1 ? max_pos - base : max_pos);
	}
}

template <typename Q>
__device__ __forceinline__ bool mg_sp_row_window(const Q &q, int64_t i, RowCache &rc, int &chrom, int64_t &ws, int64_t &we)
{
	int64_t s, e, scope;
	mg_get_row_cached(q.rows, q.first + i, rc, chrom, s, e, scope);
	return mg_window(s, e, q.sshift, q.eshift, __ldg(q.chrom_sizes + chrom), ws, we);
}

template <typename Q>
__device__ __forceinline__ void mg_sp_thread_hull(const Q &q, int64_t base, int64_t &mn, int64_t &mx, int &clo, int &chi)
{
	mn = INT64_MAX;
	mx = INT64_MIN;
	clo = INT_MAX;
	chi = INT_MIN;
	RowCache rc;
#pragma unroll
	for (int j = 0; j < MG_SP_ROWS; ++j) {
		const int64_t i = base + (int64_t)j * MG_SP_THREADS;
		if (i < q.count) {
			int chrom;
			int64_t ws, we;
			if (mg_sp_row_window(q, i, rc, chrom, ws, we)) {
				mn = min(mn, ws);
				mx = max(mx, we);
				clo = min(clo, chrom);
				chi = max(chi, chrom);
			}
		}
	}
}

// A block handles MG_SP_THREADS x MG_SP_ROWS consecutive rows; thread t takes rows base + j * MG_SP_THREADS + t (coalesced).
__global__ void __launch_bounds__(MG_SP_THREADS) k_sparse_window(const SparseQuery q)
{
	__shared__ SpSmem sm;
	const int64_t base = (int64_t)blockIdx.x * (MG_SP_THREADS * MG_SP_ROWS) + threadIdx.x;
	int64_t mn, mx;
	int clo, chi, mode;
	mg_sp_thread_hull(q, base, mn, mx, clo, chi);
	const SpLocal blk = mg_sp_narrow(sm, q.table, mn, mx, clo, chi, true, mode);
	RowCache rc;
	for (int j = 0; j < MG_SP_ROWS; ++j) {
		const int64_t i = base + (int64_t)j * MG_SP_THREADS;
		if (i >= q.count)
			break;
		int chrom;
		int64_t ws, we;
		const bool ok = mg_sp_row_window(q, i, rc, chrom, ws, we);
		SpLocal ch = blk;
		if (mode == 0 && ok) {
			const SparseChrom c = q.table[chrom];
			ch = SpLocal{c.start, c.end, c.val, c.n};
		}
		mg_sparse_row<true>(q, ch, ok, ws, we, i);
	}
}

// coverage = sum of overlaps with the unified source intervals / window length. src/IntervVarProcessor.cpp:242-325.
// floor(x / b) and ceil(x / b) for |x| < 2^30 by a multiplication: x is lifted by K b (a multiple of b above 2^30) so that the
// unsigned magic-number division applies; magic = ceil(2^64 / b). Filled on the host (the divisor is the iterator's bin size).
struct CovDiv {
	unsigned long long magic;
	int b, K;
};
static inline CovDiv mg_covdiv_make(int64_t b)
{
	CovDiv d;
	d.b = (int)b;
	d.magic = b > 1 ? ~0ull / (unsigned long long)b + 1ull : 0ull;
	d.K = (int)(((1ll << 30) + b - 1) / b);
	return d;
}

struct CoverageQuery {
	RowSrc rows;
	int64_t first, count;
	int64_t sshift, eshift;
	const SparseChrom *table;
	const int64_t *chrom_sizes;
	double *out;
	int raster; // regular-grid blocks rasterise their records (k_coverage_merge); 0: per-row cursors (ablation)
	CovDiv div; // division by the fixed-bin iterator's bin size
	GridTable grid;
};

__global__ void __launch_bounds__(MG_SP_THREADS) k_coverage(const CoverageQuery q)
{
	__shared__ SpSmem sm;
	const int64_t base = (int64_t)blockIdx.x * (MG_SP_THREADS * MG_SP_ROWS) + threadIdx.x;
	int64_t mn, mx;
	int clo, chi, mode;
	mg_sp_thread_hull(q, base, mn, mx, clo, chi);
	const SpLocal blk = mg_sp_narrow(sm, q.table, mn, mx, clo, chi, false, mode);
	RowCache rc;
	for (int j = 0; j < MG_SP_ROWS; ++j) {
		const int64_t i = base + (int64_t)j * MG_SP_THREADS;
		if (i >= q.count)
			break;
		int chrom;
		int64_t ws, we;
		const bool ok = mg_sp_row_window(q, i, rc, chrom, ws, we);
		SpLocal ch = blk;
		if (mode == 0 && ok) {
			const SparseChrom c = q.table[chrom];
			ch = SpLocal{c.start, c.end, c.val, c.n};
		}
		double cov = mg_nan();
		if (ok) {
			// first source interval with end > ws (== lower_bound by start, minus one when the previous one still overlaps)
			int64_t lo = 0, hi = ch.n;
			while (lo < hi) {
				const int64_t mid = lo + ((hi - lo) >> 1);
				if (ch.end[mid] > ws)
					hi = mid;
				else
					lo = mid + 1;
			}
			int64_t total = 0;
			for (int64_t r = lo; r < ch.n; ++r) {
				const int64_t rs = ch.start[r];
				if (rs >= we)
					break;
				const int64_t re = ch.end[r];
				total += (we < re ? we : re) - (ws > rs ? ws : rs);
			}
			cov = (double)total / (double)(we - ws);
		}
		mg_st_stream(q.out + i, cov);
	}
}

// ------------------------------------------------------------------------------------------------------------------
// Merge kernels for the common function sets (no stddev / quantile). Per block of 2048 rows:
//   * the records any of its rows can touch are bracketed by two lookups in the coarse position index (no search) and
//     staged in shared memory as 32-bit offsets from the block hull's start (12 bytes per record);
//   * lane l of a warp takes rows w0 + l + 32 k, k = 0..7, so stores stay coalesced and the "first record with
//     end > ws" of a lane only moves forward by a record or two between its consecutive rows (a restart plus a binary
//     search covers unsorted rows);
//   * blocks that span chromosomes, hold more than MG_SF_CAP records or do not fit 32-bit offsets take the per-row
//     search path (mg_sparse_row / the coverage search) over the row's whole chromosome.
// ROWS (template parameter of everything below) = rows per thread: a block takes 256 x ROWS rows. The per-block set-up -- scope
// lookup, coarse-index lookups, staging the records: three dependent memory round trips behind two barriers -- is paid once per
// block, so more rows per block amortise it (C4: 8 -> 16 rows per thread took nearest from 0.44 to 0.39 ms, coverage from 0.41
// to 0.32 ms); the host picks the largest ROWS whose expected record count per block stays well inside MG_SF_CAP.
#ifndef MG_SF_CAP
#define MG_SF_CAP 1024
#endif
#define MG_SF_SENT (1 << 30) // sentinel coordinate; real offsets stay within +-MG_SF_RANGE so distances never overflow
#define MG_SF_RANGE (1 << 28)

// Staged records live at indices 1..n between sentinels: [0] = {-SENT, -SENT}, [n+1] = [n+2] = {+SENT, +SENT}. With them
// "first record with end > a" always exists, the neighbours lo - 1 / lo + 1 can be read unguarded, and a sentinel never
// wins the nearest-neighbour comparison against a real record.
// VAL = false (coverage: interval sets carry no values) drops the value array and its 4 KB of shared memory.
template <bool VAL>
struct SfSmemT {
	static constexpr bool has_val = VAL;
	int start[MG_SF_CAP + 3];
	int end[MG_SF_CAP + 3];
	float val[VAL ? MG_SF_CAP + 3 : 1];
	int64_t red_min[MG_SP_THREADS / 32], red_max[MG_SP_THREADS / 32];
	int red_clo[MG_SP_THREADS / 32], red_chi[MG_SP_THREADS / 32];
	int64_t origin; // staged positions are relative to the block hull's start
	int64_t lo;     // first staged record
	int n;          // staged records
	int chrom;
	int mode;       // 1 = staged, 0 = per-row search over the whole chromosome
	// regular-grid blocks: what warp 0 worked out for everybody (mg_sf_grid_setup)
	int64_t g_origin, g_first;
	int g_chrom, g_len, g_nrows, g_n, g_status;
};
typedef SfSmemT<true> SfSmem;

// rows of lane l: block_base + warp * 32 * ROWS + l + 32 k
template <int ROWS, typename Q>
__device__ __forceinline__ int64_t mg_sf_row(const Q &q, int k)
{
	return (int64_t)blockIdx.x * (MG_SP_THREADS * ROWS) + (int64_t)(threadIdx.x >> 5) * (32 * ROWS) + (threadIdx.x & 31) + 32 * k;
}

// windows of this lane's rows; bit k of the result = row k exists and its window is non-empty
template <int ROWS, typename Q>
__device__ __forceinline__ unsigned mg_sf_windows(const Q &q, int64_t (&ws)[ROWS], int64_t (&we)[ROWS], int (&chrom)[ROWS])
{
	unsigned ok = 0;
	const int64_t i0 = mg_sf_row<ROWS>(q, 0);
	if (q.rows.kind == 0) {
#pragma unroll
		for (int k = 0; k < ROWS; ++k) {
			const int64_t i = i0 + 32 * k;
			chrom[k] = 0;
			ws[k] = we[k] = 0;
			if (i < q.count) {
				const int64_t r = q.first + i;
				chrom[k] = __ldg(q.rows.chrom + r);
				const int64_t s = __ldg(q.rows.start + r), e = __ldg(q.rows.end + r);
				if (mg_window(s, e, q.sshift, q.eshift, __ldg(q.chrom_sizes + chrom[k]), ws[k], we[k]))
					ok |= 1u << k;
			}
		}
		return ok;
	}
	// implicit fixed-bin rows (src/TrackExpressionFixedBinIterator.cpp:52-60)
	const int64_t binsize = q.rows.binsize, step = 32 * binsize;
	int64_t r = q.first + i0, kscope = -1, row1 = 0, sstart = 0, send = 0, csize = 0, coord = 0;
	int c = 0;
	if (i0 + 32 * (ROWS - 1) < q.count) {
		// Regular case: all rows of the lane are interior rows of ONE scope interval and no window is clamped. Then row k's
		// window is the first row's shifted by k * 32 bins: one scope lookup and two additions per row.
		kscope = mg_scope_of_row(q.rows, r);
		const int64_t row0 = __ldg(q.rows.srow0 + kscope);
		row1 = __ldg(q.rows.srow0 + kscope + 1);
		sstart = __ldg(q.rows.sstart + kscope);
		send = __ldg(q.rows.send + kscope);
		c = __ldg(q.rows.schrom + kscope);
		csize = __ldg(q.chrom_sizes + c);
		coord = (__ldg(q.rows.sbin0 + kscope) + (r - row0)) * binsize;
		const int64_t last = coord + (ROWS - 1) * step;
		const int64_t w0 = coord + q.sshift, len = binsize + q.eshift - q.sshift;
		if (r > row0 && r + 32 * (ROWS - 1) < row1 - 1 && w0 >= 0 && last + binsize + q.eshift <= csize && len > 0) {
#pragma unroll
			for (int k = 0; k < ROWS; ++k) {
				chrom[k] = c;
				ws[k] = w0 + k * step;
				we[k] = ws[k] + len;
			}
			return ROWS == 32 ? 0xffffffffu : (1u << (ROWS & 31)) - 1u;
		}
		kscope = -1; // fall through to the general path
	}
#pragma unroll 1
	for (int k = 0; k < ROWS; ++k, r += 32, coord += step) {
		int64_t wsk = 0, wek = 0;
		bool okk = false;
		if (i0 + 32 * k < q.count) {
			if (kscope < 0 || r >= row1) {
				kscope = kscope < 0 ? mg_scope_of_row(q.rows, r) : kscope + 1;
				while (r >= __ldg(q.rows.srow0 + kscope + 1))
					++kscope;
				row1 = __ldg(q.rows.srow0 + kscope + 1);
				sstart = __ldg(q.rows.sstart + kscope);
				send = __ldg(q.rows.send + kscope);
				c = __ldg(q.rows.schrom + kscope);
				csize = __ldg(q.chrom_sizes + c);
				coord = (__ldg(q.rows.sbin0 + kscope) + (r - __ldg(q.rows.srow0 + kscope))) * binsize;
			}
			const int64_t s = coord > sstart ? coord : sstart;
			const int64_t e = coord + binsize < send ? coord + binsize : send;
			okk = mg_window(s, e, q.sshift, q.eshift, csize, wsk, wek);
		}
		// ws[] / we[] must stay in registers: write them with compile-time indices
#pragma unroll
		for (int j = 0; j < ROWS; ++j)
			if (j == k) {
				chrom[j] = c;
				ws[j] = wsk;
				we[j] = wek;
			}
		if (okk)
			ok |= 1u << k;
	}
	return ok;
}

// block hull -> record slice -> shared memory. Every thread passes the hull of its own valid rows.
template <int ROWS, typename SM>
__device__ __forceinline__ void mg_sf_stage(SM &sm, const SparseChrom *table, unsigned okmask, const int64_t (&ws)[ROWS],
											 const int64_t (&we)[ROWS], const int (&chrom)[ROWS], bool want_val)
{
	int64_t mn = INT64_MAX, mx = INT64_MIN;
	int clo = INT_MAX, chi = INT_MIN;
#pragma unroll
	for (int k = 0; k < ROWS; ++k)
		if ((okmask >> k) & 1u) {
			mn = min(mn, ws[k]);
			mx = max(mx, we[k]);
			clo = min(clo, chrom[k]);
			chi = max(chi, chrom[k]);
		}
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	clo = __reduce_min_sync(0xffffffffu, clo);
	chi = __reduce_max_sync(0xffffffffu, chi);
	// genomic coordinates almost always fit 31 bits: then the hull is two 32-bit warp reductions
	const bool small = okmask == 0 || (uint64_t)mx < 0x7fffffffull; // windows are clamped to >= 0, so mn <= mx < 2^31 too
	if (__all_sync(0xffffffffu, small)) {
		const int m1 = __reduce_min_sync(0xffffffffu, okmask ? (int)mn : INT_MAX);
		const int m2 = __reduce_max_sync(0xffffffffu, okmask ? (int)mx : INT_MIN);
		mn = m1 == INT_MAX ? INT64_MAX : (int64_t)m1;
		mx = m2 == INT_MIN ? INT64_MIN : (int64_t)m2;
	} else {
#pragma unroll
		for (int k = 16; k > 0; k >>= 1) {
			mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, k));
			mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, k));
		}
	}
	if (lane == 0) {
		sm.red_min[warp] = mn;
		sm.red_max[warp] = mx;
		sm.red_clo[warp] = clo;
		sm.red_chi[warp] = chi;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		for (int w = 1; w < MG_SP_THREADS / 32; ++w) {
			mn = min(mn, sm.red_min[w]);
			mx = max(mx, sm.red_max[w]);
			clo = min(clo, sm.red_clo[w]);
			chi = max(chi, sm.red_chi[w]);
		}
		int mode = 1, n = 0;
		int64_t first = 0;
		sm.chrom = 0;
		if (clo != INT_MAX) { // the block has at least one valid row
			mode = clo == chi ? 1 : 0;
			sm.chrom = clo;
			if (mode) {
				const SparseChrom ch = table[clo];
				if (ch.n > 0) {
					const int64_t c0 = min(max(mn, (int64_t)0) >> ch.coarse_shift, ch.ncoarse - 1);
					const int64_t c1 = min((max(mx, (int64_t)0) >> ch.coarse_shift) + 1, ch.ncoarse - 1);
					// record coarse[c1] ends past a cell start >= the hull's end, so the record after it starts past the hull
					const int64_t lo = (int64_t)__ldg(ch.coarse + c0), hi = (int64_t)__ldg(ch.coarse + c1) + 1;
					const int64_t a = lo > 0 ? lo - 1 : 0, b = min(hi + 1, ch.n); // one neighbour on each side
					first = a;
					const int64_t cnt = b - a;
					if (cnt > MG_SF_CAP || mx - mn > MG_SF_RANGE || mn - __ldg(ch.start + a) > MG_SF_RANGE || __ldg(ch.end + b - 1) - mn > MG_SF_RANGE)
						mode = 0;
					else
						n = (int)cnt;
				}
			}
		}
		sm.origin = mn;
		sm.lo = first;
		sm.n = n;
		sm.mode = mode;
	}
	__syncthreads();
	if (sm.mode == 1 && sm.n > 0) {
		const SparseChrom ch = table[sm.chrom];
		const int64_t origin = sm.origin, lo = sm.lo;
		const int n = sm.n;
		for (int t = threadIdx.x; t < n + 3; t += MG_SP_THREADS) {
			int st = t == 0 ? -MG_SF_SENT : MG_SF_SENT, en = st;
			float v = mg_nanf();
			if (t >= 1 && t <= n) {
				st = (int)(__ldg(ch.start + lo + t - 1) - origin);
				en = (int)(__ldg(ch.end + lo + t - 1) - origin);
				if (want_val)
					v = __ldg(ch.val + lo + t - 1);
			}
			sm.start[t] = st;
			sm.end[t] = en;
			if (SM::has_val)
				sm.val[t] = v;
		}
	}
	__syncthreads();
}

// first staged index >= lo with end > a (the right sentinel always qualifies): a few linear steps, then a binary search
__device__ __forceinline__ int mg_sf_advance(const int *__restrict__ end, int n, int lo, int a)
{
	int steps = 0;
	while (end[lo] <= a) {
		++lo;
		if (++steps == 4) {
			int hi = n + 1;
			while (lo < hi) {
				const int mid = (lo + hi) >> 1;
				if (end[mid] > a)
					hi = mid;
				else
					lo = mid + 1;
			}
			break;
		}
	}
	return lo;
}

// the per-row search path of a block the merge kernel cannot stage (windows are recomputed: indexing the caller's ws[] / we[]
// by a loop counter would put them in local memory)
template <int ROWS>
__device__ __forceinline__ void mg_sparse_rows_search(const SparseQuery &q)
{
	RowCache rc;
#pragma unroll 1
	for (int k = 0; k < ROWS; ++k) {
		const int64_t i = mg_sf_row<ROWS>(q, k);
		if (i >= q.count)
			break;
		int c;
		int64_t s, e;
		const bool ok = mg_sp_row_window(q, i, rc, c, s, e);
		const SparseChrom ch = q.table[ok ? c : 0];
		mg_sparse_row<false>(q, SpLocal{ch.start, ch.end, ch.val, ch.n}, ok, s, e, i);
	}
}

// ---- regular-grid blocks -------------------------------------------------------------------------------------------------
// A block of implicit fixed-bin rows that lies strictly inside one scope interval with no clamped window is a regular grid:
// row j's window is [origin + j * binsize, + len). Every thread derives that from blockIdx alone -- no per-row coordinate
// arithmetic, no hull reduction, no elected thread: the coarse-index lookups are uniform loads and the block stages its
// records after ONE barrier.
struct SfGrid {
	int chrom;
	int64_t origin; // window start of the block's first row
	int len;        // window length
	int nrows;      // rows of this block
};

template <int ROWS, typename Q>
__device__ __forceinline__ bool mg_sf_grid_block(const Q &q, SfGrid &g)
{
	if (q.rows.kind != 1)
		return false;
	const int64_t i0 = (int64_t)blockIdx.x * (MG_SP_THREADS * ROWS);
	const int64_t nrows = min((int64_t)(MG_SP_THREADS * ROWS), q.count - i0);
	if (q.grid.n >= 0) { // the host's table
		if (q.grid.n == 0)
			return false;
		const unsigned b = blockIdx.x;
		int lo = 0, hi = q.grid.n; // last entry with bfirst <= b
		while (hi - lo > 1) {
			const int mid = (lo + hi) >> 1;
			if (q.grid.bfirst[mid] <= b)
				lo = mid;
			else
				hi = mid;
		}
		if (b < q.grid.bfirst[lo] || b > q.grid.blast[lo])
			return false;
		g.chrom = q.grid.chrom[lo];
		g.origin = q.grid.origin0[lo] + (int64_t)(b - q.grid.bfirst[lo]) * q.grid.step;
		g.len = q.grid.len;
		g.nrows = (int)nrows;
		return true;
	}
	const int64_t r0 = q.first + i0, r1 = r0 + nrows - 1;
	const int64_t k = mg_scope_of_row(q.rows, r0);
	const int64_t row0 = __ldg(q.rows.srow0 + k), row1 = __ldg(q.rows.srow0 + k + 1);
	if (!(r0 > row0 && r1 < row1 - 1)) // first / last row of a scope interval may be cut by its ends
		return false;
	const int64_t binsize = q.rows.binsize;
	g.chrom = __ldg(q.rows.schrom + k);
	const int64_t w0 = (__ldg(q.rows.sbin0 + k) + (r0 - row0)) * binsize + q.sshift, len = binsize + q.eshift - q.sshift;
	const int64_t span = (nrows - 1) * binsize + len;
	if (w0 < 0 || len <= 0 || span > MG_SF_RANGE || w0 + span > __ldg(q.chrom_sizes + g.chrom))
		return false;
	g.origin = w0;
	g.len = (int)len;
	g.nrows = (int)nrows;
	return true;
}

// Regular-grid set-up of a thread block: 0 = not a regular grid (general path), 1 = grid and its records are staged,
// 2 = grid whose records do not fit the stage (count or 32-bit range): per-row search path. Warp 0 alone does the scope lookup,
// the grid geometry and the two coarse-index lookups that bracket the records (uniform work: seven warps skip ~200
// instructions each, they would only wait at the barrier anyway); after one barrier every thread stages its record.
template <int ROWS, typename Q, typename SM>
__device__ __forceinline__ int mg_sf_grid_setup(const Q &q, SM &sm, const SparseChrom *table, bool want_val, SfGrid &g, int &n_out,
												double *qt = nullptr, int qt_size = 0)
{
	if (q.rows.kind != 1)
		return 0;
	if (threadIdx.x < 32) {
		SfGrid gg;
		gg.chrom = 0;
		gg.origin = 0;
		gg.len = gg.nrows = 0;
		int status = 0, n = 0;
		int64_t first = 0;
		if (mg_sf_grid_block<ROWS>(q, gg)) {
			status = 1;
			const SparseChrom ch = table[gg.chrom];
			if (ch.n > 0) {
				const int64_t hull_end = gg.origin + (int64_t)(gg.nrows - 1) * q.rows.binsize + gg.len;
				const int64_t c0 = min(gg.origin >> ch.coarse_shift, ch.ncoarse - 1), c1 = min((hull_end >> ch.coarse_shift) + 1, ch.ncoarse - 1);
				const int64_t lo = (int64_t)__ldg(ch.coarse + c0), hi = (int64_t)__ldg(ch.coarse + c1) + 1;
				first = lo > 0 ? lo - 1 : 0;
				const int64_t cnt = min(hi + 1, ch.n) - first;
				if (cnt > MG_SF_CAP)
					status = 2;
				else
					n = (int)cnt;
			}
		}
		if (threadIdx.x == 0) {
			sm.g_origin = gg.origin;
			sm.g_first = first;
			sm.g_chrom = gg.chrom;
			sm.g_len = gg.len;
			sm.g_nrows = gg.nrows;
			sm.g_n = n;
			sm.g_status = status;
		}
	}
	__syncthreads();
	const int status = sm.g_status;
	if (status != 1)
		return status;
	g.chrom = sm.g_chrom;
	g.origin = sm.g_origin;
	g.len = sm.g_len;
	g.nrows = sm.g_nrows;
	const int n = sm.g_n;
	bool bad = false;
	// coverage: the len + 1 possible quotients total / len, one division per thread instead of one per row (visible after the
	// barrier that ends the staging)
	if (qt && g.len < qt_size && (int)threadIdx.x <= g.len)
		qt[threadIdx.x] = (double)(int)threadIdx.x / (double)g.len;
	if ((int)threadIdx.x < n + 3) {
		const SparseChrom ch = table[g.chrom];
		const int64_t first = sm.g_first;
		for (int t = threadIdx.x; t < n + 3; t += MG_SP_THREADS) {
			int st = t == 0 ? -MG_SF_SENT : MG_SF_SENT, en = st;
			float v = mg_nanf();
			if (t >= 1 && t <= n) {
				const int64_t rs = __ldg(ch.start + first + t - 1) - g.origin, re = __ldg(ch.end + first + t - 1) - g.origin;
				bad |= rs < -MG_SF_RANGE || re > MG_SF_RANGE;
				st = (int)rs;
				en = (int)re;
				if (want_val)
					v = __ldg(ch.val + first + t - 1);
			}
			sm.start[t] = st;
			sm.end[t] = en;
			if (SM::has_val)
				sm.val[t] = v;
		}
	}
	n_out = n;
	return __syncthreads_or(bad) ? 2 : 1;
}

// Statistics of one window [a, b) over the staged records (indices 1..n between sentinels); lo = this lane's cursor.
struct SpVals {
	double avg, sum, mn, mx, nearest, size, exists;
};

template <bool SUM, bool MM>
__device__ __forceinline__ void mg_sf_eval(const SfSmem &sm, int n, int &lo, int a, int b, SpVals &o)
{
	o.avg = o.sum = o.mn = o.mx = o.nearest = mg_nan();
	o.size = 0;
	o.exists = 0;
	if (n <= 0)
		return;
	lo = mg_sf_advance(sm.end, n, lo, a);
	const int s0 = sm.start[lo];
	if (!(sm.start[lo + 1] < b)) {
		// at most one record overlaps the window (src/GenomeTrackSparse.cpp:281-338): its value is every statistic;
		// otherwise the closer neighbour (ties to the earlier record) is `nearest` and nothing else is defined
		const bool ov = s0 < b;
		const int dp = a - sm.end[lo - 1], dn = s0 - b; // end[lo-1] <= a < b <= start[lo] when nothing overlaps
		const float v = sm.val[(ov || dp > dn) ? lo : lo - 1];
		o.nearest = (double)v;
		if (ov && !isnan(v)) {
			o.avg = o.sum = o.mn = o.mx = (double)v;
			o.size = 1;
			o.exists = 1;
		}
		return;
	}
	int cnt = 0;
	double S = 0;
	float fmin = 3.402823466e+38f, fmax = -3.402823466e+38f;
	for (int r = lo; sm.start[r] < b; ++r) { // record order = the reference's order
		const float v = sm.val[r];
		if (isnan(v))
			continue;
		if (SUM)
			S += (double)v;
		if (MM) {
			fmin = fminf(fmin, v);
			fmax = fmaxf(fmax, v);
		}
		++cnt;
	}
	o.size = (double)cnt;
	if (cnt > 0) {
		o.exists = 1;
		if (SUM) {
			o.avg = o.nearest = mg_f2d(cnt == 1 ? S : S / (double)cnt);
			o.sum = mg_f2d(S);
		}
		if (MM) {
			o.mn = (double)fmin;
			o.mx = (double)fmax;
		}
	}
}

// F = the single requested function (one store per row, everything else compiled out) or -1 = any subset of the basic ones.
template <int F>
__device__ __forceinline__ void mg_sf_store(const SparseQuery &q, int64_t i, const SpVals &o)
{
	if (F >= 0) {
		const double v = F == MG_AVG ? o.avg : F == MG_NEAREST ? o.nearest : F == MG_SUM ? o.sum : F == MG_MIN ? o.mn : F == MG_MAX ? o.mx : F == MG_SIZE ? o.size : o.exists;
		mg_st_stream(q.out[F] + i, v);
	} else {
		if (q.out[MG_AVG]) mg_st_stream(q.out[MG_AVG] + i, o.avg);
		if (q.out[MG_NEAREST]) mg_st_stream(q.out[MG_NEAREST] + i, o.nearest);
		if (q.out[MG_SUM]) mg_st_stream(q.out[MG_SUM] + i, o.sum);
		if (q.out[MG_MIN]) mg_st_stream(q.out[MG_MIN] + i, o.mn);
		if (q.out[MG_MAX]) mg_st_stream(q.out[MG_MAX] + i, o.mx);
		if (q.out[MG_SIZE]) mg_st_stream(q.out[MG_SIZE] + i, o.size);
		if (q.out[MG_EXISTS]) mg_st_stream(q.out[MG_EXISTS] + i, o.exists);
	}
}

#ifndef MG_SF_MINBLOCKS
#define MG_SF_MINBLOCKS 8 // resident blocks per SM: barrier- and latency-bound, measured 0.77 -> 0.46 ms (nearest, C4) from 3 to 8
#endif
template <int F, int ROWS>
__global__ void __launch_bounds__(MG_SP_THREADS, MG_SF_MINBLOCKS) k_sparse_merge(const SparseQuery q)
{
	__shared__ SfSmem sm;
	constexpr bool SUM = F < 0 || F == MG_AVG || F == MG_SUM || F == MG_NEAREST, MM = F < 0 || F == MG_MIN || F == MG_MAX;
	SfGrid g;
	int n;
	const int grid = mg_sf_grid_setup<ROWS>(q, sm, q.table, true, g, n); // block-uniform
	if (grid == 2) {
		mg_sparse_rows_search<ROWS>(q);
		return;
	}
	if (grid == 1) {
		const int j0 = (threadIdx.x >> 5) * (32 * ROWS) + (threadIdx.x & 31), step = 32 * (int)q.rows.binsize;
		const int64_t i0 = (int64_t)blockIdx.x * (MG_SP_THREADS * ROWS) + j0;
		int a = j0 * (int)q.rows.binsize, lo = 1;
#pragma unroll
		for (int k = 0; k < ROWS; ++k, a += step) {
			if (j0 + 32 * k >= g.nrows)
				break;
			SpVals o;
			mg_sf_eval<SUM, MM>(sm, n, lo, a, a + g.len, o);
			mg_sf_store<F>(q, i0 + 32 * k, o);
		}
		return;
	}
	int64_t ws[ROWS], we[ROWS];
	int chrom[ROWS];
	const unsigned okmask = mg_sf_windows<ROWS>(q, ws, we, chrom);
	mg_sf_stage<ROWS>(sm, q.table, okmask, ws, we, chrom, true);
	if (sm.mode == 0) { // rare: per-row search over the row's whole chromosome
		mg_sparse_rows_search<ROWS>(q);
		return;
	}
	const int64_t origin = sm.origin;
	n = sm.n;
	int lo = 1, prev = INT_MIN;
#pragma unroll
	for (int k = 0; k < ROWS; ++k) {
		const int64_t i = mg_sf_row<ROWS>(q, k);
		if (i >= q.count)
			break;
		SpVals o;
		o.avg = o.sum = o.mn = o.mx = o.nearest = o.size = o.exists = mg_nan();
		if ((okmask >> k) & 1u) {
			const int a = (int)(ws[k] - origin), b = (int)(we[k] - origin);
			if (a < prev)
				lo = 1; // unsorted rows: restart
			prev = a;
			mg_sf_eval<SUM, MM>(sm, n, lo, a, b, o);
		}
		mg_sf_store<F>(q, i, o);
	}
}

// coverage of [a, b) by the staged source intervals: 0 / len and len / len need no divide (most rows of a fine iterator lie
// fully outside or inside a source interval)
template <typename SM>
__device__ __forceinline__ double mg_sf_coverage(const SM &sm, int n, int &lo, int a, int b)
{
	int total = 0;
	if (n > 0) {
		lo = mg_sf_advance(sm.end, n, lo, a);
		for (int r = lo; sm.start[r] < b; ++r)
			total += min(b, sm.end[r]) - max(a, sm.start[r]);
	}
	return total == 0 ? 0. : (total == b - a ? 1. : (double)total / (double)(b - a));
}

template <int ROWS>
__device__ __forceinline__ void mg_coverage_rows_search(const CoverageQuery &q)
{
	RowCache rc;
#pragma unroll 1
	for (int k = 0; k < ROWS; ++k) {
		const int64_t i = mg_sf_row<ROWS>(q, k);
		if (i >= q.count)
			break;
		double cov = mg_nan();
		int c;
		int64_t s, e;
		if (mg_sp_row_window(q, i, rc, c, s, e)) {
			const SparseChrom ch = q.table[c];
			int64_t lo = 0, hi = ch.n;
			while (lo < hi) {
				const int64_t mid = lo + ((hi - lo) >> 1);
				if (ch.end[mid] > s)
					hi = mid;
				else
					lo = mid + 1;
			}
			int64_t total = 0;
			for (int64_t r = lo; r < ch.n; ++r) {
				const int64_t rs = ch.start[r];
				if (rs >= e)
					break;
				const int64_t re = ch.end[r];
				total += (e < re ? e : re) - (s > rs ? s : rs);
			}
			cov = (double)total / (double)(e - s);
		}
		mg_st_stream(q.out + i, cov);
	}
}

// ---- coverage of a regular-grid block by rasterising the records ---------------------------------------------------------
// Row j's window is [j b, j b + len) in block coordinates. Instead of every row searching its records, every staged record
// [s, e) marks the rows it touches. The source intervals are disjoint, so a row is either covered completely by ONE record
// (ceil(s / b) <= j <= floor((e - len) / b): its total is len, a plain store) or touched partly by a few (at most
// 2 ceil(len / b) + 2 rows per record: shared-memory atomic adds of the overlap). Each warp rasterises its own 256 rows -- lane r
// takes the r-th record that can touch them -- so after the block barriers of the staging everything is warp-synchronous. A row
// then costs two shared-memory reads and its store, against ~87 instructions per row of the per-row cursor path.
// Same integers as the cursor path (total overlap, window length), so the same doubles come out.
#define MG_COV_RASTER_MAX_BINS 8 // windows of up to this many bins (len <= 8 b): beyond that a record touches too many rows
#define MG_COV_QT 256 // window lengths below this take their quotient total / len from a per-block table
template <int ROWS>
struct CovRaster {
	double qt[MG_COV_QT];                              // qt[t] = (double)t / (double)len
	int acc[MG_SP_THREADS / 32][(32 * ROWS)];      // per warp: the rows' total overlap
};

__device__ __forceinline__ int mg_floordiv_i(int x, const CovDiv &d)
{
	if (d.b == 1)
		return x;
	const unsigned long long y = (unsigned long long)((long long)x + (long long)d.K * d.b); // < 2^32
	return (int)__umul64hi(y, d.magic) - d.K;
}
__device__ __forceinline__ int mg_ceildiv_i(int x, const CovDiv &d) { return -mg_floordiv_i(-x, d); }

template <int ROWS, typename SM>
__device__ __forceinline__ void mg_cov_raster(const SM &sm, CovRaster<ROWS> &cr, int n, const SfGrid &g, const CovDiv &b, double *__restrict__ out)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, len = g.len, bsz = b.b;
	const int row0 = warp * (32 * ROWS), wrows = min((32 * ROWS), g.nrows - row0);
	if (wrows <= 0)
		return;
	int *acc = cr.acc[warp];
#pragma unroll
	for (int k = 0; k < ROWS; ++k)
		acc[lane + 32 * k] = 0;
	__syncwarp();
	if (n > 0) {
		const int A = row0 * bsz, B = (row0 + wrows - 1) * bsz + len; // hull of the warp's windows
		int lo = 1, hi = n + 1;                                        // first staged record with end > A (the right sentinel qualifies)
		while (lo < hi) {
			const int mid = (lo + hi) >> 1;
			if (sm.end[mid] > A)
				hi = mid;
			else
				lo = mid + 1;
		}
		for (int r = lo + lane; r <= n && sm.start[r] < B; r += 32) {
			const int s = sm.start[r] - A, e = sm.end[r] - A; // in the warp's coordinates: row j's window starts at j b
			if (e <= s)
				continue;
			const int jlo = max(0, mg_floordiv_i(s - len, b) + 1); // j b + len > s
			const int jhi = min(wrows - 1, mg_ceildiv_i(e, b) - 1); // j b < e
			if (jlo > jhi)
				continue;
			const int flo = max(jlo, mg_ceildiv_i(s, b));        // j b >= s
			const int fhi = min(jhi, mg_floordiv_i(e - len, b)); // j b + len <= e
			int p1 = jhi, p2 = jhi + 1; // partly covered rows: [jlo, p1] and [p2, jhi]
			if (flo <= fhi) {
				for (int j = flo; j <= fhi; ++j) // no other record touches these rows
					acc[j] = len;
				p1 = flo - 1;
				p2 = fhi + 1;
			}
			for (int j = jlo; j <= p1; ++j)
				atomicAdd(&acc[j], min(j * bsz + len, e) - max(j * bsz, s));
			for (int j = p2; j <= jhi; ++j)
				atomicAdd(&acc[j], min(j * bsz + len, e) - max(j * bsz, s));
		}
	}
	__syncwarp();
	out += row0;
#pragma unroll
	for (int k = 0; k < ROWS; ++k) {
		const int j = lane + 32 * k;
		if (j < wrows) {
			const int total = acc[j];
			mg_st_stream(out + j, len < MG_COV_QT ? cr.qt[total] : (total == 0 ? 0. : (total == len ? 1. : (double)total / (double)len)));
		}
	}
}

template <int ROWS>
__global__ void __launch_bounds__(MG_SP_THREADS, MG_SF_MINBLOCKS) k_coverage_merge(const CoverageQuery q)
{
	__shared__ SfSmemT<false> sm;
	__shared__ CovRaster<ROWS> cr;
	SfGrid g;
	int n;
	const int grid = mg_sf_grid_setup<ROWS>(q, sm, q.table, false, g, n, cr.qt, MG_COV_QT); // block-uniform
	if (grid == 2) {
		mg_coverage_rows_search<ROWS>(q);
		return;
	}
	if (grid == 1) {
		if (q.raster && (int64_t)g.len <= MG_COV_RASTER_MAX_BINS * q.rows.binsize && (int64_t)(MG_SP_THREADS * ROWS) * q.rows.binsize + g.len < MG_SF_RANGE) {
			mg_cov_raster<ROWS>(sm, cr, n, g, q.div, q.out + (int64_t)blockIdx.x * (MG_SP_THREADS * ROWS));
			return;
		}
		const int j0 = (threadIdx.x >> 5) * (32 * ROWS) + (threadIdx.x & 31), step = 32 * (int)q.rows.binsize;
		const int64_t i0 = (int64_t)blockIdx.x * (MG_SP_THREADS * ROWS) + j0;
		int a = j0 * (int)q.rows.binsize, lo = 1;
#pragma unroll
		for (int k = 0; k < ROWS; ++k, a += step) {
			if (j0 + 32 * k >= g.nrows)
				break;
			mg_st_stream(q.out + i0 + 32 * k, mg_sf_coverage(sm, n, lo, a, a + g.len));
		}
		return;
	}
	int64_t ws[ROWS], we[ROWS];
	int chrom[ROWS];
	const unsigned okmask = mg_sf_windows<ROWS>(q, ws, we, chrom);
	mg_sf_stage<ROWS>(sm, q.table, okmask, ws, we, chrom, false);
	if (sm.mode == 0) {
		mg_coverage_rows_search<ROWS>(q);
		return;
	}
	const int64_t origin = sm.origin;
	n = sm.n;
	int lo = 1, prev = INT_MIN;
#pragma unroll
	for (int k = 0; k < ROWS; ++k) {
		const int64_t i = mg_sf_row<ROWS>(q, k);
		if (i >= q.count)
			break;
		double cov = mg_nan();
		if ((okmask >> k) & 1u) {
			const int a = (int)(ws[k] - origin), b = (int)(we[k] - origin);
			if (a < prev)
				lo = 1;
			prev = a;
			cov = mg_sf_coverage(sm, n, lo, a, b);
		}
		mg_st_stream(q.out + i, cov);
	}
}
