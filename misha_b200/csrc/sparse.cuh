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

struct SparseQuery {
	RowSrc rows;
	int64_t first, count;
	int64_t sshift, eshift;
	const SparseChrom *table;
	const int64_t *chrom_sizes;
	double *out[MG_NUM_FUNCS];
	int nq;
	double q_pct[MG_MAX_Q];
	double *q_out[MG_MAX_Q];
};

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
			const float v = __ldg(val + r);
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

// One thread per row. read_interval: src/GenomeTrackSparse.cpp:237-340; calc_vals: src/GenomeTrackSparse.h:143-259.
// Sums and Welford updates run in record order, exactly the reference's order => avg / sum / stddev are bit-exact.
__global__ void __launch_bounds__(256) k_sparse_window(const SparseQuery q)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= q.count)
		return;
	int chrom;
	int64_t s, e, scope;
	mg_get_row(q.rows, q.first + i, chrom, s, e, scope);
	int64_t ws, we;
	double avg = mg_nan(), sum = mg_nan(), mn = mg_nan(), mx = mg_nan(), sd = mg_nan(), nearest = mg_nan(), size = mg_nan(),
		   exists = mg_nan();
	double qv[MG_MAX_Q];
#pragma unroll
	for (int k = 0; k < MG_MAX_Q; ++k)
		qv[k] = mg_nan();
	if (mg_window(s, e, q.sshift, q.eshift, __ldg(q.chrom_sizes + chrom), ws, we)) {
		const SparseChrom ch = q.table[chrom];
		size = 0;
		exists = 0;
		if (ch.n > 0) {
			if (__ldg(ch.start) >= we)
				nearest = (double)__ldg(ch.val); // before the first record (:281-284)
			else if (__ldg(ch.end + ch.n - 1) <= ws)
				nearest = (double)__ldg(ch.val + ch.n - 1); // after the last record (:286-289)
			else {
				// first record whose end > ws (ends are sorted because records are disjoint)
				int64_t lo = 0, hi = ch.n;
				while (lo < hi) {
					const int64_t mid = lo + ((hi - lo) >> 1);
					if (__ldg(ch.end + mid) > ws)
						hi = mid;
					else
						lo = mid + 1;
				}
				const int64_t first = lo;
				if (__ldg(ch.start + first) < we) {
					long long n = 0;
					double S = 0, mean = 0, m2 = 0;
					float fmin = 3.402823466e+38f, fmax = -3.402823466e+38f;
					int64_t r = first;
					for (; r < ch.n && __ldg(ch.start + r) < we; ++r) {
						const float v = __ldg(ch.val + r);
						if (isnan(v))
							continue;
						S += (double)v;
						if (v < fmin) fmin = v;
						if (v > fmax) fmax = v;
						++n;
						const double d = (double)v - mean;
						mean += d / (double)n;
						m2 += d * ((double)v - mean);
					}
					size = (double)n;
					if (n > 0) {
						exists = 1;
						avg = nearest = mg_f2d(S / (double)n);
						sum = mg_f2d(S);
						mn = (double)fmin;
						mx = (double)fmax;
						if (n > 1)
							sd = mg_f2d(sqrt(m2 / (double)(n - 1)));
						for (int k = 0; k < q.nq; ++k) {
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
					const int64_t dp = mg_dist(ws, we, __ldg(ch.start + prv), __ldg(ch.end + prv));
					const int64_t dn = mg_dist(ws, we, __ldg(ch.start + first), __ldg(ch.end + first));
					nearest = (double)__ldg(ch.val + (dp <= dn ? prv : first));
				}
			}
		}
	}
	if (q.out[MG_AVG]) mg_st_stream(q.out[MG_AVG] + i, avg);
	if (q.out[MG_NEAREST]) mg_st_stream(q.out[MG_NEAREST] + i, nearest);
	if (q.out[MG_SUM]) mg_st_stream(q.out[MG_SUM] + i, sum);
	if (q.out[MG_MIN]) mg_st_stream(q.out[MG_MIN] + i, mn);
	if (q.out[MG_MAX]) mg_st_stream(q.out[MG_MAX] + i, mx);
	if (q.out[MG_STDDEV]) mg_st_stream(q.out[MG_STDDEV] + i, sd);
	if (q.out[MG_SIZE]) mg_st_stream(q.out[MG_SIZE] + i, size);
	if (q.out[MG_EXISTS]) mg_st_stream(q.out[MG_EXISTS] + i, exists);
	for (int k = 0; k < q.nq; ++k)
		mg_st_stream(q.q_out[k] + i, qv[k]);
}

// coverage = sum of overlaps with the unified source intervals / window length. src/IntervVarProcessor.cpp:242-325.
struct CoverageQuery {
	RowSrc rows;
	int64_t first, count;
	int64_t sshift, eshift;
	const SparseChrom *table;
	const int64_t *chrom_sizes;
	double *out;
};

__global__ void __launch_bounds__(256) k_coverage(const CoverageQuery q)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= q.count)
		return;
	int chrom;
	int64_t s, e, scope;
	mg_get_row(q.rows, q.first + i, chrom, s, e, scope);
	int64_t ws, we;
	double cov = mg_nan();
	if (mg_window(s, e, q.sshift, q.eshift, __ldg(q.chrom_sizes + chrom), ws, we)) {
		const SparseChrom ch = q.table[chrom];
		// first source interval with end > ws (== lower_bound by start, minus one when the previous one still overlaps)
		int64_t lo = 0, hi = ch.n;
		while (lo < hi) {
			const int64_t mid = lo + ((hi - lo) >> 1);
			if (__ldg(ch.end + mid) > ws)
				hi = mid;
			else
				lo = mid + 1;
		}
		int64_t total = 0;
		for (int64_t r = lo; r < ch.n; ++r) {
			const int64_t rs = __ldg(ch.start + r);
			if (rs >= we)
				break;
			const int64_t re = __ldg(ch.end + r);
			total += (we < re ? we : re) - (ws > rs ? ws : rs);
		}
		cov = (double)total / (double)(we - ws);
	}
	mg_st_stream(q.out + i, cov);
}
