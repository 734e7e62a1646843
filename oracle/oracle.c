/* oracle/oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-threaded CPU restatement of misha's track-extraction / aggregation hot path
 * (tanaylab/misha v5.11.20; citations are paths under the reference checkout). It exists to CHECK the CUDA
 * path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product (misha_b200/) never calls into this file and has no CPU fallback.
 *
 * Parity pinning: every function below is checked (tests/test_oracle_*.py) against
 *   (a) the reference's own compiled objects (oracle/_ref/libmisha_ref.so, built by oracle/Makefile from the
 *       unmodified sources) on the bundled example DB and on seeded random inputs, and
 *   (b) the known answers the reference's tests hold for this path (tests/testthat/test-gdist.R:69-72,
 *       test-vtrack-coverage.R:5-69, test-iterator-overlaps.R:444) and the fixtures in tests/golden/.
 *
 * Stateless restatement: the reference keeps per-track cursors and sliding-window accumulators
 * (Kahan running sum src/GenomeTrackFixedBin.cpp:295-365, sequential hint src/GenomeTrackSparse.cpp:293-303)
 * that change low-order bits / rare corner cases depending on the previous row. This file restates the
 * "fresh window" semantics; DESIGN.md lists the history-dependent corners.
 */
#include <ctype.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_NAN (nan(""))

/* ------------------------------------------------------------------------------------------------------ */
/* Window transform: src/TrackExpressionVars.h:395-402 (Iterator_modifier1D::transform).                    */
/* Returns 0 when the shifted window is out of range (start' >= end').                                      */
static int orc_window(int64_t start, int64_t end, int64_t sshift, int64_t eshift, int64_t chromsize, int64_t *ws, int64_t *we)
{
	int64_t s = start + sshift, e = end + eshift;
	if (s < 0)
		s = 0;
	if (e > chromsize)
		e = chromsize;
	*ws = s;
	*we = e;
	return s < e;
}

static int cmp_float(const void *a, const void *b)
{
	float x = *(const float *)a, y = *(const float *)b;
	return (x > y) - (x < y);
}

/* src/StreamPercentiler.h:191-217 (exact regime: stream size <= reservoir). vals is sorted in place. */
float orc_percentile_sorted(const float *sorted, int64_t n, double percentile)
{
	if (n <= 0)
		return (float)ORC_NAN;
	if (percentile < 0.)
		percentile = 0.;
	if (percentile > 1.)
		percentile = 1.;
	double index = (double)(n - 1) * percentile;
	uint64_t i1 = (uint64_t)floor(index), i2 = (uint64_t)ceil(index);
	double w = index - (double)i1;
	return (float)((double)sorted[i1] * (1 - w) + (double)sorted[i2] * w);
}

float orc_percentile(float *vals, int64_t n, double percentile)
{
	qsort(vals, (size_t)n, sizeof(float), cmp_float);
	return orc_percentile_sorted(vals, n, percentile);
}

static int cmp_double(const void *a, const void *b)
{
	double x = *(const double *)a, y = *(const double *)b;
	return (x > y) - (x < y);
}

/* gquantiles, exact regime (every value fits the reservoir). src/GenomeTrackQuantiles.cpp:167-173: each row's value is  */
/* narrowed to float, NaN rows are dropped; :183-232 / src/StreamPercentiler.h:210-217: the quantile of the N kept      */
/* values is samples[floor(idx)] * (1 - w) + samples[ceil(idx)] * w in double with idx = (N - 1) * clamp(pct, 0, 1);   */
/* NaN when nothing was kept (:113-116). Returns N.                                                                    */
int64_t orc_gquantiles(const double *vals, int64_t n, int nq, const double *pct, double *out)
{
	double *s = (double *)malloc((size_t)(n > 0 ? n : 1) * sizeof(double));
	int64_t N = 0;
	for (int64_t i = 0; i < n; ++i) {
		float v = (float)vals[i];
		if (!isnan(v))
			s[N++] = (double)v;
	}
	qsort(s, (size_t)N, sizeof(double), cmp_double);
	for (int k = 0; k < nq; ++k) {
		if (!N) {
			out[k] = ORC_NAN;
			continue;
		}
		double p = pct[k] < 0. ? 0. : (pct[k] > 1. ? 1. : pct[k]);
		double idx = (double)(N - 1) * p;
		uint64_t i1 = (uint64_t)floor(idx), i2 = (uint64_t)ceil(idx);
		double w = idx - (double)i1;
		out[k] = s[i1] * (1.0 - w) + s[i2] * w;
	}
	free(s);
	return N;
}

/* ------------------------------------------------------------------------------------------------------ */
/* Dense (fixed-bin) track window functions.                                                               */
/* src/GenomeTrackFixedBin.cpp:607-1016 (generic path), :460-591 (single-function path), :1028-1075         */
/* (read_bins_bulk: clamp to num_samples, +-inf -> NaN). Results are float members widened to double        */
/* (src/GenomeTrack1D.h:58-103, src/TrackVarProcessor.cpp:113-173).                                         */
/* Any out_* may be NULL. bins = raw file payload (may contain +-inf).                                      */
void orc_dense_window(const float *bins, int64_t nbins, uint32_t bin_size, int64_t chromsize, int64_t n, const int64_t *start,
					  const int64_t *end, int64_t sshift, int64_t eshift, double percentile, double *out_avg, double *out_sum,
					  double *out_min, double *out_max, double *out_stddev, double *out_quantile, double *out_size)
{
	float *scratch = NULL;
	int64_t scratch_cap = 0;
	for (int64_t i = 0; i < n; ++i) {
		double avg = ORC_NAN, sum = ORC_NAN, mn = ORC_NAN, mx = ORC_NAN, sd = ORC_NAN, q = ORC_NAN, size = 0;
		int64_t ws, we;
		int in_range = orc_window(start[i], end[i], sshift, eshift, chromsize, &ws, &we);
		if (!in_range)
			size = ORC_NAN; /* src/TrackVarProcessor.cpp:55-58: every function is NaN */
		if (in_range) {
			int64_t sbin = (int64_t)(ws / bin_size);                    /* :675 */
			int64_t ebin = (int64_t)ceil(we / (double)bin_size);        /* :676 */
			if (ebin > nbins)
				ebin = nbins; /* read_bins_bulk clamp :1036-1041 */
			int64_t nb = ebin - sbin;
			if (nb > scratch_cap) {
				scratch_cap = nb * 2;
				scratch = (float *)realloc(scratch, (size_t)scratch_cap * sizeof(float));
			}
			uint64_t num_vs = 0;
			double sum_accum = 0, w_mean = 0, w_m2 = 0;
			float fmin = 3.402823466e+38F, fmax = -3.402823466e+38F;
			for (int64_t b = sbin; b < ebin; ++b) {
				float v = bins[b];
				if (isinf(v) || isnan(v))
					continue;
				sum_accum += v;              /* :882 */
				if (v < fmin)
					fmin = v;                /* :886 */
				if (v > fmax)
					fmax = v;                /* :905 */
				++num_vs;
				double delta = v - w_mean;   /* Welford :912-917 */
				w_mean += delta / (double)num_vs;
				double delta2 = v - w_mean;
				w_m2 += delta * delta2;
				scratch[num_vs - 1] = v;     /* quantile feed :922-923 */
			}
			size = (double)num_vs;
			if (num_vs > 0) {
				avg = (double)(float)(sum_accum / num_vs); /* :970 */
				sum = (double)(float)sum_accum;            /* :968 */
				mn = fmin;
				mx = fmax;
				if (num_vs > 1)
					sd = (double)(float)sqrt(w_m2 / (double)(num_vs - 1)); /* :986 */
				if (out_quantile)
					q = (double)orc_percentile(scratch, (int64_t)num_vs, percentile);
			}
		}
		if (out_avg) out_avg[i] = avg;
		if (out_sum) out_sum[i] = sum;
		if (out_min) out_min[i] = mn;
		if (out_max) out_max[i] = mx;
		if (out_stddev) out_stddev[i] = sd;
		if (out_quantile) out_quantile[i] = q;
		if (out_size) out_size[i] = size;
	}
	free(scratch);
}

/* ------------------------------------------------------------------------------------------------------ */
/* Sparse track window functions. Records (rs,re,rv) sorted by start, non-overlapping                      */
/* (src/GenomeTrackSparse.cpp:194-222; +-inf -> NaN at load :205-206).                                     */
/* read_interval: src/GenomeTrackSparse.cpp:237-340; calc_vals: src/GenomeTrackSparse.h:143-259;           */
/* overlap is half-open (src/Segment.h:44); dist2interv src/GInterval.cpp:33-44.                            */
static int64_t orc_dist(int64_t qs, int64_t qe, int64_t rs, int64_t re)
{
	int64_t lo = qs > rs ? qs : rs, hi = qe < re ? qe : re;
	if (lo < hi)
		return 0;
	int64_t left = llabs(rs - qe), right = llabs(re - qs);
	return left <= right ? left : right;
}

void orc_sparse_window(const int64_t *rs, const int64_t *re, const float *rv, int64_t nrec, int64_t chromsize, int64_t n,
					   const int64_t *start, const int64_t *end, int64_t sshift, int64_t eshift, double percentile, double *out_avg,
					   double *out_sum, double *out_min, double *out_max, double *out_stddev, double *out_nearest,
					   double *out_quantile, double *out_size)
{
	float *scratch = NULL;
	int64_t scratch_cap = 0;
	for (int64_t i = 0; i < n; ++i) {
		double avg = ORC_NAN, sum = ORC_NAN, mn = ORC_NAN, mx = ORC_NAN, sd = ORC_NAN, nearest = ORC_NAN, q = ORC_NAN, size = 0;
		int64_t ws, we;
		if (!orc_window(start[i], end[i], sshift, eshift, chromsize, &ws, &we))
			size = ORC_NAN;
		else if (nrec > 0) {
			if (rs[0] >= we)
				nearest = isinf(rv[0]) ? ORC_NAN : rv[0];                 /* :281-284 */
			else if (re[nrec - 1] <= ws)
				nearest = isinf(rv[nrec - 1]) ? ORC_NAN : rv[nrec - 1];   /* :286-289 */
			else {
				/* first record with end > ws (records are disjoint and sorted, so ends are sorted too) */
				int64_t lo = 0, hi = nrec;
				while (lo < hi) {
					int64_t mid = lo + (hi - lo) / 2;
					if (re[mid] > ws)
						hi = mid;
					else
						lo = mid + 1;
				}
				int64_t first = lo; /* < nrec because re[nrec-1] > ws */
				if (rs[first] < we) {
					/* overlap: aggregate over the run of overlapping records (calc_vals) */
					uint64_t num_vs = 0;
					double sum_accum = 0, w_mean = 0, w_m2 = 0;
					float fmin = 3.402823466e+38F, fmax = -3.402823466e+38F;
					for (int64_t r = first; r < nrec && rs[r] < we; ++r) {
						float v = rv[r];
						if (isnan(v) || isinf(v))
							continue;
						if (num_vs + 1 > (uint64_t)scratch_cap) {
							scratch_cap = scratch_cap ? scratch_cap * 2 : 256;
							scratch = (float *)realloc(scratch, (size_t)scratch_cap * sizeof(float));
						}
						sum_accum += v;
						if (v < fmin) fmin = v;
						if (v > fmax) fmax = v;
						++num_vs;
						double delta = v - w_mean;
						w_mean += delta / (double)num_vs;
						double delta2 = v - w_mean;
						w_m2 += delta * delta2;
						scratch[num_vs - 1] = v;
					}
					size = (double)num_vs;
					if (num_vs > 0) {
						avg = nearest = (double)(float)(sum_accum / num_vs);
						sum = (double)(float)sum_accum;
						mn = fmin;
						mx = fmax;
						if (num_vs > 1)
							sd = (double)(float)sqrt(w_m2 / (double)(num_vs - 1));
						if (out_quantile)
							q = (double)orc_percentile(scratch, (int64_t)num_vs, percentile);
					}
				} else {
					/* no overlap: the closer of the two neighbours, ties -> earlier record (:336-338) */
					int64_t nxt = first, prv = first - 1; /* prv >= 0 because rs[0] < we and no overlap */
					float v;
					if (orc_dist(ws, we, rs[prv], re[prv]) <= orc_dist(ws, we, rs[nxt], re[nxt]))
						v = rv[prv];
					else
						v = rv[nxt];
					nearest = isinf(v) ? ORC_NAN : v;
				}
			}
		}
		if (out_avg) out_avg[i] = avg;
		if (out_sum) out_sum[i] = sum;
		if (out_min) out_min[i] = mn;
		if (out_max) out_max[i] = mx;
		if (out_stddev) out_stddev[i] = sd;
		if (out_nearest) out_nearest[i] = nearest;
		if (out_quantile) out_quantile[i] = q;
		if (out_size) out_size[i] = size;
	}
	free(scratch);
}

/* ------------------------------------------------------------------------------------------------------ */
/* The order-dependent functions: lse, first / last (+ position), min.pos / max.pos.                       */
/* Output columns of both functions below (ORC_EXT_COLS doubles per row):                                  */
/*   0 lse  1 first  2 last  3 first.pos  4 last.pos  5 min.pos  6 max.pos   (positions absolute)          */
#define ORC_EXT_COLS 7

/* src/GenomeTrack1D.h:20-30 (the float overload the track readers use) */
static void orc_lse_accumulate(float *l1, float l2)
{
	if (*l1 > l2) {
		if (!isinf(l2))
			*l1 += logf(1.0f + expf(l2 - *l1));
	} else {
		if (isinf(*l1))
			*l1 = l2;
		else
			*l1 = l2 + logf(1.0f + expf(*l1 - l2));
	}
}

/* Dense, generic path src/GenomeTrackFixedBin.cpp:881-939: position of a bin = max(bin start, window start) (:883-884);
 * min.pos keeps the first bin of the minimum (:886-895), max.pos the first bin of the maximum (:897-903).
 * lse: lse_mode 0 = the reducer path's value, (float)(M + log(sum exp(v - M))) as RunningLogSumExp::push / value build it
 * for a fresh window (src/utils/RunningLogSumExp.h:82-96,175-178; src/GenomeTrackFixedBin.cpp:379-392) -- what a vtrack with
 * func = "lse" alone on its track runs; lse_mode 1 = the generic path's pairwise float accumulation (:918-919). */
void orc_dense_window_ext(const float *bins, int64_t nbins, uint32_t bin_size, int64_t chromsize, int64_t n, const int64_t *start,
						  const int64_t *end, int64_t sshift, int64_t eshift, int lse_mode, double *out /* n x ORC_EXT_COLS */)
{
	for (int64_t i = 0; i < n; ++i) {
		double *o = out + i * ORC_EXT_COLS;
		for (int k = 0; k < ORC_EXT_COLS; ++k)
			o[k] = ORC_NAN;
		int64_t ws, we;
		if (!orc_window(start[i], end[i], sshift, eshift, chromsize, &ws, &we))
			continue;
		int64_t sbin = (int64_t)(ws / bin_size);
		int64_t ebin = (int64_t)ceil(we / (double)bin_size);
		if (ebin == sbin + 1) {
			/* a window inside one bin: assign_single_bin_value (:150-180, called at :703-708) sets every position to the
			 * bin's overlap start and lse / first / last to the bin's value WHATEVER that value is (a NaN bin still reports
			 * its position); past the end of the track everything stays NaN (assign_single_bin_missing) */
			if (sbin < nbins) {
				float v = bins[sbin];
				double bin_start = (double)(sbin * (int64_t)bin_size);
				double pos = bin_start > (double)ws ? bin_start : (double)ws;
				o[0] = o[1] = o[2] = isinf(v) ? ORC_NAN : (double)v;
				o[3] = o[4] = o[5] = o[6] = pos;
			}
			continue;
		}
		if (ebin > nbins)
			ebin = nbins;
		uint64_t num_vs = 0;
		float fmin = 3.402823466e+38F, fmax = -3.402823466e+38F, flse = -INFINITY;
		double M = -INFINITY, sum_scaled = 0;
		for (int64_t b = sbin; b < ebin; ++b) {
			float v = bins[b];
			if (isinf(v) || isnan(v))
				continue;
			double bin_start = (double)(b * (int64_t)bin_size);
			double pos = bin_start > (double)ws ? bin_start : (double)ws;
			if (v < fmin) {
				fmin = v;
				o[5] = pos;
			} else if (v == fmin && (isnan(o[5]) || pos < o[5]))
				o[5] = pos;
			if (v > fmax) {
				fmax = v;
				o[6] = pos;
			}
			orc_lse_accumulate(&flse, v);
			if (v > M) { /* RunningLogSumExp::push */
				if (isfinite(M))
					sum_scaled *= exp(M - (double)v);
				M = v;
			}
			sum_scaled += exp((double)v - M);
			if (num_vs == 0) {
				o[1] = v;
				o[3] = pos;
			}
			o[2] = v;
			o[4] = pos;
			++num_vs;
		}
		if (num_vs > 0)
			o[0] = lse_mode ? (double)flse : (double)(float)(M + log(sum_scaled));
		else
			o[5] = ORC_NAN;
	}
}

/* Sparse, calc_vals src/GenomeTrackSparse.h:166-224: positions are the records' starts; lse accumulates in float, record order. */
void orc_sparse_window_ext(const int64_t *rs, const int64_t *re, const float *rv, int64_t nrec, int64_t chromsize, int64_t n,
						   const int64_t *start, const int64_t *end, int64_t sshift, int64_t eshift, double *out /* n x ORC_EXT_COLS */)
{
	for (int64_t i = 0; i < n; ++i) {
		double *o = out + i * ORC_EXT_COLS;
		for (int k = 0; k < ORC_EXT_COLS; ++k)
			o[k] = ORC_NAN;
		int64_t ws, we;
		if (!orc_window(start[i], end[i], sshift, eshift, chromsize, &ws, &we) || nrec == 0 || rs[0] >= we || re[nrec - 1] <= ws)
			continue;
		int64_t lo = 0, hi = nrec;
		while (lo < hi) {
			int64_t mid = lo + (hi - lo) / 2;
			if (re[mid] > ws)
				hi = mid;
			else
				lo = mid + 1;
		}
		uint64_t num_vs = 0;
		float fmin = 3.402823466e+38F, fmax = -3.402823466e+38F, flse = -INFINITY;
		for (int64_t r = lo; r < nrec && rs[r] < we; ++r) {
			float v = rv[r];
			if (isnan(v) || isinf(v))
				continue;
			double pos = (double)rs[r];
			if (v < fmin) {
				fmin = v;
				o[5] = pos;
			} else if (v == fmin && (isnan(o[5]) || pos < o[5]))
				o[5] = pos;
			if (v > fmax) {
				fmax = v;
				o[6] = pos;
			}
			orc_lse_accumulate(&flse, v);
			if (num_vs == 0) {
				o[1] = v;
				o[3] = pos;
			}
			o[2] = v;
			o[4] = pos;
			++num_vs;
		}
		if (num_vs > 0)
			o[0] = (double)flse;
		else
			o[5] = ORC_NAN;
	}
}

/* ------------------------------------------------------------------------------------------------------ */
/* Intervals: sort by (chromid,start) and unify overlaps. src/GIntervals.cpp:59-74.                        */
typedef struct { int32_t chrom; int64_t start, end, idx; } orc_iv;

static int cmp_iv(const void *a, const void *b)
{
	const orc_iv *x = (const orc_iv *)a, *y = (const orc_iv *)b;
	if (x->chrom != y->chrom)
		return x->chrom < y->chrom ? -1 : 1;
	if (x->start != y->start)
		return x->start < y->start ? -1 : 1;
	return (x->idx > y->idx) - (x->idx < y->idx); /* stable */
}

/* In place; returns the new count. */
int64_t orc_unify(int64_t n, int32_t *chrom, int64_t *start, int64_t *end, int unify_touching)
{
	if (n <= 0)
		return 0;
	orc_iv *v = (orc_iv *)malloc((size_t)n * sizeof(orc_iv));
	for (int64_t i = 0; i < n; ++i) {
		v[i].chrom = chrom[i]; v[i].start = start[i]; v[i].end = end[i]; v[i].idx = i;
	}
	qsort(v, (size_t)n, sizeof(orc_iv), cmp_iv);
	int64_t cur = 0;
	for (int64_t i = 1; i < n; ++i) {
		if (v[cur].chrom != v[i].chrom || v[cur].end < v[i].start || (!unify_touching && v[cur].end == v[i].start))
			v[++cur] = v[i];
		else if (v[cur].end < v[i].end)
			v[cur].end = v[i].end;
	}
	for (int64_t i = 0; i <= cur; ++i) {
		chrom[i] = v[i].chrom; start[i] = v[i].start; end[i] = v[i].end;
	}
	free(v);
	return cur + 1;
}

/* Sort permutation of a scope by (chromid, start), stable. perm[k] = original row of the k-th sorted interval. */
void orc_sort_perm(int64_t n, const int32_t *chrom, const int64_t *start, int64_t *perm)
{
	orc_iv *v = (orc_iv *)malloc((size_t)(n ? n : 1) * sizeof(orc_iv));
	for (int64_t i = 0; i < n; ++i) {
		v[i].chrom = chrom[i]; v[i].start = start[i]; v[i].end = 0; v[i].idx = i;
	}
	qsort(v, (size_t)n, sizeof(orc_iv), cmp_iv);
	for (int64_t i = 0; i < n; ++i)
		perm[i] = v[i].idx;
	free(v);
}

/* ------------------------------------------------------------------------------------------------------ */
/* Coverage vtrack (func = "coverage" on an intervals source). src/IntervVarProcessor.cpp:242-325.         */
/* Source intervals (one chromosome): sorted + overlap-unified (src/TrackExpressionVars.cpp:1299-1301).     */
void orc_coverage(const int64_t *cs, const int64_t *ce, int64_t nsrc, int64_t chromsize, int64_t n, const int64_t *start,
				  const int64_t *end, int64_t sshift, int64_t eshift, double *out)
{
	for (int64_t i = 0; i < n; ++i) {
		int64_t ws, we;
		if (!orc_window(start[i], end[i], sshift, eshift, chromsize, &ws, &we)) {
			out[i] = ORC_NAN;
			continue;
		}
		/* lower_bound by start, then also look at the previous interval (:280-292) */
		int64_t lo = 0, hi = nsrc;
		while (lo < hi) {
			int64_t mid = lo + (hi - lo) / 2;
			if (cs[mid] < ws)
				lo = mid + 1;
			else
				hi = mid;
		}
		int64_t total = 0;
		if (lo > 0 && ce[lo - 1] > ws) {
			int64_t os = ws > cs[lo - 1] ? ws : cs[lo - 1], oe = we < ce[lo - 1] ? we : ce[lo - 1];
			total += oe - os;
		}
		for (int64_t k = lo; k < nsrc && cs[k] < we; ++k) {
			if (ce[k] > ws) {
				int64_t os = ws > cs[k] ? ws : cs[k], oe = we < ce[k] ? we : ce[k];
				total += oe - os;
			}
		}
		out[i] = (double)total / (double)(we - ws); /* :324 */
	}
}

/* ------------------------------------------------------------------------------------------------------ */
/* gsummary accumulator: src/GenomeTrackSummary.cpp:22-69 and the output vector :148-155.                  */
/* out = {Total intervals, NaN intervals, Min, Max, Sum, Mean, Std dev}.                                   */
void orc_summary(const double *v, int64_t n, double out[7])
{
	double num_bins = 0, nn = 0, total = 0, minval = 1.7976931348623157e308, maxval = -1.7976931348623157e308, ssq = 0;
	for (int64_t i = 0; i < n; ++i) {
		++num_bins;
		if (!isnan(v[i])) {
			nn++;
			total += v[i];
			if (v[i] < minval) minval = v[i];
			if (v[i] > maxval) maxval = v[i];
			ssq += v[i] * v[i];
		}
	}
	out[0] = num_bins;
	out[1] = num_bins - nn;
	out[2] = nn ? minval : ORC_NAN;
	out[3] = nn ? maxval : ORC_NAN;
	out[4] = nn ? total : ORC_NAN;
	out[5] = nn ? total / nn : ORC_NAN;
	if (nn > 1) {
		double mean = total / nn;
		double var = ssq / (nn - 1) - (mean * mean) * (nn / (nn - 1));
		out[6] = var > 0 ? sqrt(var) : 0;
	} else
		out[6] = ORC_NAN;
}

/* ------------------------------------------------------------------------------------------------------ */
/* gdist binning: BinFinder::init src/BinFinder.cpp:10-43, val2bin src/BinFinder.h:55-108 (right=true is    */
/* what gdist uses), BinsManager::vals2idx src/BinsManager.h:53-72.                                        */
static double orc_uniform_binsize(const double *breaks, int nbreaks)
{
	double binsize = breaks[1] - breaks[0];
	for (int i = 1; i < nbreaks; ++i)
		if ((float)(breaks[i] - breaks[i - 1]) != (float)binsize)
			binsize = 0;
	if (!isfinite(binsize))
		binsize = 0;
	return binsize;
}

int orc_val2bin(const double *breaks, int nbreaks, int include_lowest, double val)
{
	int numbins = nbreaks - 1;
	if (include_lowest && val == breaks[0])
		return 0;
	if (isnan(val) || val <= breaks[0] || val > breaks[nbreaks - 1])
		return -1;
	double binsize = orc_uniform_binsize(breaks, nbreaks);
	if (binsize) {
		int b = (int)ceil((val - breaks[0]) / binsize) - 1;
		return b < numbins - 1 ? b : numbins - 1;
	}
	unsigned s = 0, e = (unsigned)numbins;
	while (e - s > 1) {
		unsigned m = (s + e) / 2;
		if (val <= breaks[m])
			e = m;
		else
			s = m;
	}
	return (int)s;
}

/* N-D histogram over rows; vals is row-major n x ndim; breaks concatenated; counts has prod(nbins) cells,   */
/* first dimension fastest (mult_0 = 1), src/BinsManager.cpp ctor + src/GenomeTrackDistribution.cpp:121-131. */
void orc_hist(const double *vals, int64_t n, int ndim, const double *breaks, const int *nbreaks, int include_lowest, uint64_t *counts)
{
	int64_t mult[16];
	const double *bptr[16];
	int64_t total = 1;
	const double *p = breaks;
	for (int d = 0; d < ndim; ++d) {
		mult[d] = total;
		total *= nbreaks[d] - 1;
		bptr[d] = p;
		p += nbreaks[d];
	}
	memset(counts, 0, (size_t)total * sizeof(uint64_t));
	for (int64_t i = 0; i < n; ++i) {
		int64_t idx = 0;
		int ok = 1;
		for (int d = 0; d < ndim && ok; ++d) {
			double v = vals[i * ndim + d];
			int b = isnan(v) ? -1 : orc_val2bin(bptr[d], nbreaks[d], include_lowest, v);
			if (b < 0)
				ok = 0;
			else
				idx += b * mult[d];
		}
		if (ok)
			counts[idx]++;
	}
}

/* ------------------------------------------------------------------------------------------------------ */
/* Iterators. Scope must already be sorted by (chromid,start) (src/GenomeTrackExtract.cpp:557-562).         */
/* Fixed-bin: src/TrackExpressionFixedBinIterator.cpp:29-62. Returns #rows; writes at most cap.             */
int64_t orc_fixedbin_rows(int64_t binsize, int64_t nscope, const int32_t *schrom, const int64_t *sstart, const int64_t *send,
						  int64_t cap, int32_t *ochrom, int64_t *ostart, int64_t *oend, int64_t *oscope)
{
	int64_t r = 0;
	for (int64_t k = 0; k < nscope; ++k) {
		int64_t cur = (int64_t)(sstart[k] / (double)binsize);
		int64_t endb = (int64_t)ceil(send[k] / (double)binsize);
		for (; cur < endb; ++cur, ++r) {
			if (r < cap) {
				int64_t coord = cur * binsize;
				ochrom[r] = schrom[k];
				ostart[r] = coord > sstart[k] ? coord : sstart[k];
				oend[r] = coord + binsize < send[k] ? coord + binsize : send[k];
				oscope[r] = k;
			}
		}
	}
	return r;
}

/* Intervals iterator: src/TrackExpressionIntervals1DIterator.cpp:21-85 with helpers .h:25-63. Iterator      */
/* intervals are sorted and unified (touching kept apart, src/TrackExpressionScanner.cpp:948) by the caller. */
int64_t orc_intervals_rows(int64_t niter, const int32_t *ichrom, const int64_t *istart, const int64_t *iend, int64_t nscope,
						   const int32_t *schrom, const int64_t *sstart, const int64_t *send, int64_t cap, int32_t *ochrom,
						   int64_t *ostart, int64_t *oend, int64_t *oscope)
{
	int64_t r = 0;
	for (int64_t k = 0; k < nscope; ++k) {
		/* first iterator interval on this chromosome with end > scope.start */
		int64_t lo = 0, hi = niter;
		while (lo < hi) {
			int64_t mid = lo + (hi - lo) / 2;
			if (ichrom[mid] < schrom[k] || (ichrom[mid] == schrom[k] && iend[mid] <= sstart[k]))
				lo = mid + 1;
			else
				hi = mid;
		}
		for (int64_t j = lo; j < niter && ichrom[j] == schrom[k] && istart[j] < send[k]; ++j) {
			int64_t os = sstart[k] > istart[j] ? sstart[k] : istart[j];
			int64_t oe = send[k] < iend[j] ? send[k] : iend[j];
			if (os < oe) {
				if (r < cap) {
					ochrom[r] = schrom[k]; ostart[r] = os; oend[r] = oe; oscope[r] = k;
				}
				++r;
			}
		}
	}
	return r;
}

/* gscreen run-length merge of consecutive TRUE rows. src/GenomeTrackScreener.cpp:195-220, restated with     */
/* its exact control flow (including the stale interval.end comparison after a FALSE row).                   */
int64_t orc_screen_merge(int64_t n, const int32_t *chrom, const int64_t *start, const int64_t *end, const int8_t *flag,
						 int32_t *ochrom, int64_t *ostart, int64_t *oend)
{
	int64_t m = 0;
	int32_t cchrom = -1;
	int64_t cstart = -1, cend = -1;
	for (int64_t i = 0; i < n; ++i) {
		if (flag[i] == 1) {
			if (cend != start[i] || cchrom != chrom[i]) {
				if (cstart != -1) {
					ochrom[m] = cchrom; ostart[m] = cstart; oend[m] = cend; ++m;
				}
				cchrom = chrom[i]; cstart = start[i]; cend = end[i];
			} else
				cend = end[i];
		} else if (cstart != -1) {
			ochrom[m] = cchrom; ostart[m] = cstart; oend[m] = cend; ++m;
			cstart = -1;
		}
	}
	if (cstart != -1) {
		ochrom[m] = cchrom; ostart[m] = cstart; oend[m] = cend; ++m;
	}
	return m;
}

/* ------------------------------------------------------------------------------------------------------ */
/* PWM vtracks (non-spatial). See SURVEY.md 7.0(10).                                                        */
/* PSSM preparation: src/PWMScorer.cpp:1452-1459 (double -> float probs), src/DnaPSSM.cpp:77-106            */
/* (normalize: float sum, float divide, logf), :703-719 (add_dirichlet_prior), src/PwmCoreParams.cpp:20-24. */
void orc_pssm_logp(const double *pssm, int L, double prior, float *logp /* L x 4 */)
{
	for (int i = 0; i < L; ++i) {
		float p[4];
		for (int c = 0; c < 4; ++c)
			p[c] = (float)pssm[i * 4 + c];
		for (int pass = 0; pass < 2; ++pass) {
			float sum = p[0] + p[1] + p[2] + p[3];
			for (int c = 0; c < 4; ++c) {
				p[c] /= sum;
				logp[i * 4 + c] = p[c] != 0 ? logf(p[c]) : -INFINITY;
			}
			if (pass == 1 || prior <= 0)
				break;
			for (int c = 0; c < 4; ++c)
				p[c] = p[c] + (float)prior;
		}
	}
}

/* src/util.h:16-28, all float (std::exp/std::log float overloads). */
static void orc_log_sum_log(float *l1, float l2)
{
	if (*l1 > l2) {
		if (!isinf(l2))
			*l1 += logf(1 + expf(l2 - *l1));
	} else {
		if (isinf(*l1))
			*l1 = l2;
		else
			*l1 = l2 + logf(1 + expf(*l1 - l2));
	}
}

static const int8_t ORC_BASE[256] = {
	['A'] = 1, ['a'] = 1, ['C'] = 2, ['c'] = 2, ['G'] = 3, ['g'] = 3, ['T'] = 4, ['t'] = 4 }; /* code+1; 0 = not ACGT */

/* forward strand log-likelihood at seq[0..L): src/DnaPSSM.cpp:249-264 (terms added for p = 0..L-1);        */
/* when rev_order != 0 the same terms are added for p = L-1..0 (the strand = -1 call chain).      */
static float orc_like_fwd(const float *logp, int L, const char *s, int rev_order)
{
	float lp = 0;
	for (int k = 0; k < L; ++k) {
		int p = rev_order ? L - 1 - k : k;
		int code = ORC_BASE[(unsigned char)s[p]] - 1;
		if (code < 0)
			return -INFINITY;
		lp += logp[p * 4 + code];
	}
	return lp;
}

/* reverse-complement log-likelihood at seq[0..L): src/DnaPSSM.cpp:268-283 (p = L-1..0 over seq 0..L-1).     */
static float orc_like_rc(const float *logp, int L, const char *s, int rev_order)
{
	float lp = 0;
	for (int k = 0; k < L; ++k) {
		int j = rev_order ? L - 1 - k : k; /* sequence offset visited k-th */
		int p = L - 1 - j;
		int code = ORC_BASE[(unsigned char)s[j]] - 1;
		if (code < 0)
			return -INFINITY;
		lp += logp[p * 4 + (3 - code)];
	}
	return lp;
}

/* pwm.max.pos scores through DnaPSSM::max_like_match (src/DnaPSSM.cpp:285-365), whose per-base rule differs from     */
/* calc_like's: N / n / * add the row's average log-probability (DnaPSSM.h:133-135), any other non-ACGT character    */
/* adds nothing. fwd: rows p = 0..L-1 over s[0..L); rc: rows L-1..0 with complemented codes. rev_order as above.     */
static float orc_row_avg(const float *lp) { return (lp[0] + lp[1] + lp[2] + lp[3]) / 4; }
static int orc_neutral(char c) { return c == 'N' || c == 'n' || c == '*'; }

static float orc_match_fwd(const float *logp, int L, const char *s, int rev_order)
{
	float lp = 0;
	for (int k = 0; k < L; ++k) {
		int p = rev_order ? L - 1 - k : k;
		int code = ORC_BASE[(unsigned char)s[p]] - 1;
		if (orc_neutral(s[p]))
			lp += orc_row_avg(logp + p * 4);
		else if (code >= 0)
			lp += logp[p * 4 + code];
	}
	return lp;
}

static float orc_match_rc(const float *logp, int L, const char *s, int rev_order)
{
	float lp = 0;
	for (int k = 0; k < L; ++k) {
		int j = rev_order ? L - 1 - k : k;
		int p = L - 1 - j;
		int code = ORC_BASE[(unsigned char)s[j]] - 1;
		if (orc_neutral(s[j]))
			lp += orc_row_avg(logp + p * 4);
		else if (code >= 0)
			lp += logp[p * 4 + (3 - code)];
	}
	return lp;
}

/* One chromosome. seq = ASCII bases of the whole chromosome (len = chromsize).                              */
/* mode: 0 = pwm (log-sum-exp over anchors, src/utils/RunningLogSumExp.h:32-45,175-178), 1 = pwm.max,         */
/* 2 = pwm.max.pos, 3 = pwm.count (src/PWMScorer.cpp:660-690). strand: 1 / -1 (used when !bidirect and for the add order).     */
/* score_interval: src/PWMScorer.cpp:742-861; expansion src/GenomeSeqScorer.cpp:16-38.                        */
/* Spatial weights (PWMParams spat_factor / spat_bin, src/PWMScorer.cpp:112-165, src/DnaPSSM.cpp:933-1075): the anchor at TARGET   */
/* index t (the reverse complement's index for strand -1) gets log(max(1e-30f, spat[min(t / spat_bin, nspat - 1)])) added to each  */
/* strand's log-likelihood before the strands are united (pwm, pwm.max: integrate_energy_logspat / integrate_energy_max_logspat)   */
/* or to the united value before the threshold (pwm.count: count_motif_hits_with_spatial, :219-249). nspat = 0: no weights.        */
/* pwm.max.pos under spatial weights is not restated (the reference's fresh-window and sliding versions disagree on strand -1).    */
void orc_pwm_spat(const char *seq, int64_t chromsize, const float *logp, int L, int bidirect, int extend, int mode, int strand,
				  float score_thresh, const float *spat, int nspat, int spat_bin, int64_t n, const int64_t *start, const int64_t *end,
				  int64_t sshift, int64_t eshift, double *out);
void orc_pwm(const char *seq, int64_t chromsize, const float *logp, int L, int bidirect, int extend, int mode, int strand,
			 float score_thresh, int64_t n, const int64_t *start, const int64_t *end, int64_t sshift, int64_t eshift, double *out)
{
	orc_pwm_spat(seq, chromsize, logp, L, bidirect, extend, mode, strand, score_thresh, NULL, 0, 1, n, start, end, sshift, eshift, out);
}
void orc_pwm_spat(const char *seq, int64_t chromsize, const float *logp, int L, int bidirect, int extend, int mode, int strand,
				  float score_thresh, const float *spat, int nspat, int spat_bin, int64_t n, const int64_t *start, const int64_t *end,
				  int64_t sshift, int64_t eshift, double *out)
{
	float *slog = NULL;
	if (nspat > 0) {
		slog = (float *)malloc((size_t)nspat * sizeof(float));
		for (int k = 0; k < nspat; ++k)
			slog[k] = logf(spat[k] > 1e-30f ? spat[k] : 1e-30f);
		if (spat_bin < 1)
			spat_bin = 1;
	}
	float *vals = NULL;
	int64_t cap = 0;
	for (int64_t i = 0; i < n; ++i) {
		int64_t ws, we;
		out[i] = ORC_NAN;
		if (!orc_window(start[i], end[i], sshift, eshift, chromsize, &ws, &we))
			continue;
		int64_t xe = we;
		if (extend && L > 1) {
			xe = we + (L - 1);
			if (xe > chromsize)
				xe = chromsize;
		}
		int64_t tlen = xe - ws;
		if (!extend && tlen < L)
			continue;            /* :750-752 NaN */
		if (tlen < L) {          /* extend=TRUE clipped by the chromosome end: score_without_spatial over zero anchors, */
			out[i] = mode == 3 ? 0. : -INFINITY; /* src/PWMScorer.cpp:313-338 -> -inf (pwm, pwm.max) / 0 hits (pwm.count) */
			if (mode == 2) /* max_like_match returns begin with best_dir unset: index 0; the bidirect sign is undefined there */
				out[i] = bidirect ? ORC_NAN : (strand == -1 ? (double)((float)tlen - 1.0f - (float)L + 1) : 1.0);
			continue;
		}
		int64_t i_max = tlen - L;
		int64_t want = we - ws - 1;
		if (want < i_max)
			i_max = want;
		if (i_max > 1000000) /* DnaPSSM::m_max_range default (src/DnaPSSM.h:186-187; :774 and DnaPSSM.cpp:293) */
			i_max = 1000000;
		int64_t na = i_max + 1;
		/* strand -1 scores the reverse complement of [ws, xe): target index t <-> genomic anchor (tlen - L) - t, */
		/* so the first na target anchors are the last na genomic ones */
		int64_t a_base = strand == -1 ? (tlen - L + 1) - na : 0;
		if (mode == 2 && nspat > 0)
			continue; /* not restated: NaN */
		if (mode == 2) {
			/* pwm.max.pos: src/PWMScorer.cpp:325-337 -> max_like_match(combine_strands = false) over the target, */
			/* then compute_position_result (:168-182), all in float */
			int rev = strand == -1, best_dir = 1;
			int64_t best_t = -1;
			float best = -INFINITY;
			for (int64_t t = 0; t < na; ++t) {
				const char *s = seq + ws + (rev ? (tlen - L) - t : t);
				float l = rev ? orc_match_rc(logp, L, s, 1) : orc_match_fwd(logp, L, s, 0);
				float total = l;
				int dir = 1;
				if (bidirect) {
					float rl = rev ? orc_match_fwd(logp, L, s, 1) : orc_match_rc(logp, L, s, 0);
					if (rl > l) {
						total = rl;
						dir = -1;
					}
				}
				if (total > best) {
					best = total;
					best_t = t;
					best_dir = dir;
				}
			}
			if (best_t >= 0) {
				float pos = (float)best_t + 1.0f;
				if (rev)
					pos = (float)tlen - pos - (float)L + 1;
				if (bidirect)
					pos = pos * best_dir;
				out[i] = pos;
			}
			continue;
		}
		if (na > cap) {
			cap = na * 2;
			vals = (float *)realloc(vals, (size_t)cap * sizeof(float));
		}
		int rev = strand == -1;
		for (int64_t a = 0; a < na; ++a) {
			const char *s = seq + ws + a_base + a;
			float v, sl = 0.f;
			if (nspat > 0) { /* target index of this anchor, its spatial bin */
				int64_t t = rev ? na - 1 - a : a, b = t / spat_bin;
				sl = slog[b < nspat ? b : nspat - 1];
			}
			if (bidirect) {
				float f = orc_like_fwd(logp, L, s, rev), r = orc_like_rc(logp, L, s, rev);
				if (nspat > 0 && mode != 3) { /* each strand weighted, then united */
					f += sl;
					r += sl;
				}
				v = f;
				orc_log_sum_log(&v, r);
				if (nspat > 0 && mode == 3) /* pwm.count: the union, then the weight */
					v += sl;
			} else {
				v = strand == -1 ? orc_like_rc(logp, L, s, rev) : orc_like_fwd(logp, L, s, 0);
				if (nspat > 0)
					v += sl;
			}
			/* pwm.count under spatial weights: every scorer answers its first interval from spat_seed, whose per-bin hit counts stop at  */
			/* the end of the LAST spatial bin (j_end = min(W, (b + 1) B) for b < bins, src/PWMScorer.cpp:1098-1111) -- anchors past    */
			/* nspat * spat_bin are not counted, although pwm / pwm.max extend the last bin over them (bin_total_recompute, :934-937)   */
			if (nspat > 0 && mode == 3 && (rev ? na - 1 - a : a) >= (int64_t)nspat * spat_bin)
				v = -INFINITY;
			vals[a] = v;
		}
		if (mode == 0) {
			double M = -INFINITY;
			for (int64_t a = 0; a < na; ++a)
				if (vals[a] > M)
					M = vals[a];
			if (!isfinite(M))
				out[i] = -INFINITY;
			else {
				double ss = 0;
				if (rev) /* strand -1 visits anchors from the far end */
					for (int64_t a = na - 1; a >= 0; --a)
						ss += exp((double)vals[a] - M);
				else
					for (int64_t a = 0; a < na; ++a)
						ss += exp((double)vals[a] - M);
				out[i] = (double)(float)(M + log(ss));
			}
		} else if (mode == 1) {
			float m = -INFINITY;
			for (int64_t a = 0; a < na; ++a)
				if (vals[a] > m)
					m = vals[a];
			out[i] = m;
		} else if (mode == 3) {
			int64_t c = 0;
			for (int64_t a = 0; a < na; ++a)
				if (vals[a] >= score_thresh)
					++c;
			out[i] = (double)(float)c;
		}
	}
	free(vals);
	free(slog);
}

/* ------------------------------------------------------------------------------------------------------ */
/* kmer.count / kmer.frac: KmerCounter::score_interval + count_in_interval (src/KmerCounter.cpp:38-176),    */
/* interval expansion src/GenomeSeqScorer.cpp:16-38, reverse complement of the fetched target               */
/* src/GenomeSeqFetch.cpp:147-160. seq = the chromosome's bytes as stored (mixed case). mode 0 = count, 1 = */
/* fraction; strand 0 = both, 1, -1. The reference is followed step by step: fetch, reverse-complement,     */
/* upper-case, memcmp at every start inside the original interval.                                          */
static char orc_complement(char c)
{ /* basepair2complementary, src/GenomeUtils.h */
	switch (c) {
	case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a';
	case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
	default: return c;
	}
}

void orc_kmer(const char *seq, int64_t chromsize, const char *kmer_in, int mode, int extend, int strand, int64_t n, const int64_t *start,
			  const int64_t *end, int64_t sshift, int64_t eshift, double *out)
{
	const int64_t k = (int64_t)strlen(kmer_in);
	char *kmer = (char *)malloc((size_t)k + 1);
	for (int64_t j = 0; j <= k; ++j)
		kmer[j] = (char)toupper((unsigned char)kmer_in[j]);
	char *target = NULL;
	int64_t cap = 0;
	for (int64_t i = 0; i < n; ++i) {
		int64_t ws, we;
		out[i] = ORC_NAN;
		if (!orc_window(start[i], end[i], sshift, eshift, chromsize, &ws, &we))
			continue;
		int64_t xe = we;
		if (extend && k > 1) {
			xe = we + (k - 1);
			if (xe > chromsize)
				xe = chromsize;
		}
		const int64_t tlen = xe - ws;
		if (tlen > cap) {
			cap = tlen;
			target = (char *)realloc(target, (size_t)cap);
		}
		uint64_t total = 0, possible = 0;
		for (int dir = 1; dir >= -1; dir -= 2) {
			if (!(strand == 0 || strand == dir))
				continue;
			uint64_t count = 0;
			possible = 0;
			if (tlen >= k) {
				for (int64_t t = 0; t < tlen; ++t) {
					char c = dir == 1 ? seq[ws + t] : orc_complement(seq[xe - 1 - t]);
					target[t] = (char)toupper((unsigned char)c);
				}
				const int64_t ostart = 0; /* fetch start == original start: only the end is extended */
				int64_t oend = tlen - (xe - we);
				if (oend > tlen)
					oend = tlen;
				possible = oend > ostart ? (uint64_t)(oend - ostart) : 0;
				for (int64_t pos = ostart; pos < oend && pos <= tlen - k; ++pos)
					if (memcmp(target + pos, kmer, (size_t)k) == 0)
						++count;
			}
			total += count;
		}
		if (mode == 1) {
			uint64_t valid = possible > (uint64_t)(k - 1) ? possible - (uint64_t)(k - 1) : 0;
			if (strand == 0)
				valid *= 2;
			out[i] = valid > 0 ? (double)((float)total / (float)valid) : 0.0;
		} else
			out[i] = (double)(float)total;
	}
	free(target);
	free(kmer);
}

/* ------------------------------------------------------------------------------------------------------ */
/* masked.count / masked.frac: MaskedBpCounter::score_interval (src/MaskedBpCounter.cpp:19-64): lower-case  */
/* (soft-masked) bases of the window, no extension, no strand; the fraction is float / size_t -> float.     */
void orc_masked(const char *seq, int64_t chromsize, int mode, int64_t n, const int64_t *start, const int64_t *end, int64_t sshift,
				int64_t eshift, double *out)
{
	for (int64_t i = 0; i < n; ++i) {
		int64_t ws, we;
		out[i] = ORC_NAN;
		if (!orc_window(start[i], end[i], sshift, eshift, chromsize, &ws, &we))
			continue;
		uint64_t masked = 0;
		for (int64_t t = ws; t < we; ++t)
			masked += islower((unsigned char)seq[t]) != 0;
		const uint64_t len = (uint64_t)(we - ws);
		if (len == 0)
			out[i] = 0.;
		else
			out[i] = mode == 1 ? (double)((float)masked / (float)len) : (double)(float)masked;
	}
}
