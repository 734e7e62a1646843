// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into, or called from, the product path).
//
// A thin extern "C" harness over the reference's own, UNMODIFIED hot-path objects, compiled in place from
// /root/reference/src by oracle/Makefile into oracle/_ref/libmisha_ref.so (git-ignored; it travels to the GPU
// box as a prebuilt binary). It drives the L5 "data objects" exactly the way
// TrackExpressionVars::set_vars does (src/TrackExpressionVars.cpp:2080-2150): one track object per
// (track, chromosome), read_interval() once per row in row order (so the reference's history-dependent
// sliding-window paths are exercised), then the last_*() accessors.
//
// Objects driven here:
//   GenomeTrackFixedBin::read_interval        src/GenomeTrackFixedBin.cpp:607-1016
//   GenomeTrackSparse::read_interval          src/GenomeTrackSparse.cpp:237-340
//   StreamPercentiler<float>                  src/StreamPercentiler.h:170-255
//   BinFinder::val2bin                        src/BinFinder.h:55-108
//   PWMScorer::score_interval                 src/PWMScorer.cpp:742-861
//   TrackExpressionFixedBinIterator           src/TrackExpressionFixedBinIterator.cpp:29-62
//   TrackExpressionIntervals1DIterator        src/TrackExpressionIntervals1DIterator.cpp:21-85
//   GIntervals::sort / unify_overlaps         src/GIntervals.cpp:59-74
#include <memory>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <stdexcept>

#include "GenomeTrackFixedBin.h"
#include "GenomeTrackSparse.h"
#include "GenomeTrackArrays.h"
#include "StreamPercentiler.h"
#include "BinFinder.h"
#include "GenomeChromKey.h"
#include "GIntervals.h"
#include "PWMScorer.h"
#include "DnaPSSM.h"
#include "KmerCounter.h"
#include "MaskedBpCounter.h"
#include "TrackExpressionFixedBinIterator.h"
#include "TrackExpressionIntervals1DIterator.h"

static thread_local std::string g_err;

extern "C" const char *ref_last_error() { return g_err.c_str(); }

#define REF_TRY try {
#define REF_CATCH                                                                                                       \
	}                                                                                                                   \
	catch (TGLException & e) {                                                                                          \
		g_err = e.msg();                                                                                                \
		return -1;                                                                                                      \
	}                                                                                                                   \
	catch (std::exception & e) {                                                                                        \
		g_err = e.what();                                                                                               \
		return -1;                                                                                                      \
	}

// Output column order of ref_dense_eval / ref_sparse_eval.
enum { COL_AVG, COL_MIN, COL_MAX, COL_NEAREST, COL_STDDEV, COL_SUM, COL_QUANTILE, NUM_COLS };

static void register_funcs(GenomeTrack1D &t, uint32_t func_mask, int use_quantile, uint64_t reservoir)
{
	for (int f = 0; f < GenomeTrack1D::NUM_FUNCS; ++f)
		if (func_mask & (1u << f))
			t.register_function((GenomeTrack1D::Functions)f);
	if (use_quantile)
		t.register_quantile(reservoir, 100000, 100000); // gquantile.edge.data.size default, R/zzz.R:50
}

static void collect(GenomeTrack1D &t, uint32_t func_mask, int use_quantile, double percentile, float *row)
{
	const float nan = std::numeric_limits<float>::quiet_NaN();
	row[COL_AVG] = (func_mask & (1u << GenomeTrack1D::AVG)) ? t.last_avg() : nan;
	row[COL_MIN] = (func_mask & (1u << GenomeTrack1D::MIN)) ? t.last_min() : nan;
	row[COL_MAX] = (func_mask & (1u << GenomeTrack1D::MAX)) ? t.last_max() : nan;
	row[COL_NEAREST] = (func_mask & (1u << GenomeTrack1D::NEAREST)) ? t.last_nearest() : nan;
	row[COL_STDDEV] = (func_mask & (1u << GenomeTrack1D::STDDEV)) ? t.last_stddev() : nan;
	row[COL_SUM] = (func_mask & (1u << GenomeTrack1D::SUM)) ? t.last_sum() : nan;
	row[COL_QUANTILE] = use_quantile ? t.last_quantile(percentile) : nan;
}

// Evaluate `n` windows [start,end) on one chromosome file of a dense track. func_mask uses the reference's
// GenomeTrack1D::Functions bit positions (AVG=0, MIN=1, MAX=2, NEAREST=3, STDDEV=4, SUM=5).
extern "C" int ref_dense_eval(const char *chrom_file, int chromid, uint32_t func_mask, int use_quantile, double percentile,
							  int64_t n, const int64_t *start, const int64_t *end, float *out /* n x NUM_COLS */)
{
	REF_TRY
	GenomeTrackFixedBin t;
	t.init_read(chrom_file, chromid);
	register_funcs(t, func_mask, use_quantile, 1000000);
	for (int64_t i = 0; i < n; ++i) {
		GInterval iv(chromid, start[i], end[i], 0);
		t.read_interval(iv);
		collect(t, func_mask, use_quantile, percentile, out + i * NUM_COLS);
	}
	return 0;
	REF_CATCH
}

// The order-dependent functions through the reference's own readers. out: n x 7 doubles in the order of oracle.c's
// orc_*_window_ext: lse, first, last, first.pos, last.pos, min.pos, max.pos (absolute). func_mask picks what is registered:
// the reducer path (LSE alone) and the generic path (LSE next to other functions) compute lse differently.
template <class T>
static void eval_ext(T &t, int chromid, uint32_t func_mask, int64_t n, const int64_t *start, const int64_t *end, double *out)
{
	const double nan = std::numeric_limits<double>::quiet_NaN();
	for (int64_t i = 0; i < n; ++i) {
		GInterval iv(chromid, start[i], end[i], 0);
		t.read_interval(iv);
		double *o = out + i * 7;
		o[0] = (func_mask & (1u << GenomeTrack1D::LSE)) ? t.last_lse() : nan;
		o[1] = (func_mask & (1u << GenomeTrack1D::FIRST)) ? t.last_first() : nan;
		o[2] = (func_mask & (1u << GenomeTrack1D::LAST)) ? t.last_last() : nan;
		o[3] = (func_mask & (1u << GenomeTrack1D::FIRST_POS)) ? t.last_first_pos() : nan;
		o[4] = (func_mask & (1u << GenomeTrack1D::LAST_POS)) ? t.last_last_pos() : nan;
		o[5] = (func_mask & (1u << GenomeTrack1D::MIN_POS)) ? t.last_min_pos() : nan;
		o[6] = (func_mask & (1u << GenomeTrack1D::MAX_POS)) ? t.last_max_pos() : nan;
	}
}

extern "C" int ref_dense_eval_ext(const char *chrom_file, int chromid, uint32_t func_mask, int64_t n, const int64_t *start, const int64_t *end,
								  double *out)
{
	REF_TRY
	GenomeTrackFixedBin t;
	t.init_read(chrom_file, chromid);
	register_funcs(t, func_mask, 0, 0);
	eval_ext(t, chromid, func_mask, n, start, end, out);
	return 0;
	REF_CATCH
}

// size / exists / lse of a dense track as the reference's reader reports them under a given set of registered functions: the
// reducer path (only sum / lse / exists / size registered) and the generic path (anything else next to them) disagree on windows
// inside one NaN bin (src/GenomeTrackFixedBin.cpp:150-180 vs :231-240). out: n x 3 doubles {size, exists, lse}.
extern "C" int ref_dense_eval_counts(const char *chrom_file, int chromid, uint32_t func_mask, int64_t n, const int64_t *start, const int64_t *end,
									 double *out)
{
	REF_TRY
	GenomeTrackFixedBin t;
	t.init_read(chrom_file, chromid);
	register_funcs(t, func_mask, 0, 0);
	const double nan = std::numeric_limits<double>::quiet_NaN();
	for (int64_t i = 0; i < n; ++i) {
		GInterval iv(chromid, start[i], end[i], 0);
		t.read_interval(iv);
		out[3 * i + 0] = (func_mask & (1u << GenomeTrack1D::SIZE)) ? t.last_size() : nan;
		out[3 * i + 1] = (func_mask & (1u << GenomeTrack1D::EXISTS)) ? t.last_exists() : nan;
		out[3 * i + 2] = (func_mask & (1u << GenomeTrack1D::LSE)) ? t.last_lse() : nan;
	}
	return 0;
	REF_CATCH
}

extern "C" int ref_sparse_eval_ext(const char *chrom_file, int chromid, uint32_t func_mask, int64_t n, const int64_t *start, const int64_t *end,
								   double *out)
{
	REF_TRY
	GenomeTrackSparse t;
	t.init_read(chrom_file, chromid);
	register_funcs(t, func_mask, 0, 0);
	eval_ext(t, chromid, func_mask, n, start, end, out);
	return 0;
	REF_CATCH
}

// Array tracks (src/GenomeTrackArrays.{h,cpp}): written by the reference's own writer, read back through its own reader with a slice and a
// slice function set (gvtrack.array.slice). vals / idx = the records' (value, column) pairs back to back, rec_first[nrec + 1] their offsets.
extern "C" int ref_array_write(const char *chrom_file, int chromid, int64_t nrec, const int64_t *start, const int64_t *end, const int64_t *rec_first,
							   const float *vals, const uint32_t *idx)
{
	REF_TRY
	GenomeTrackArrays t;
	t.init_write(chrom_file, chromid);
	for (int64_t r = 0; r < nrec; ++r) {
		GenomeTrackArrays::ArrayVals av;
		for (int64_t k = rec_first[r]; k < rec_first[r + 1]; ++k)
			av.push_back(GenomeTrackArrays::ArrayVal(vals[k], idx[k]));
		t.write_next_interval(GInterval(chromid, start[r], end[r], 0), av.begin(), av.end());
	}
	return 0;
	REF_CATCH
}

// slice_func: GenomeTrackArrays::SliceFunctions (0 avg, 1 min, 2 max, 3 stddev, 4 sum, 5 quantile); out as ref_dense_eval (n x NUM_COLS)
extern "C" int ref_array_eval(const char *chrom_file, int chromid, int nslice, const uint32_t *slice, int slice_func, double slice_percentile,
							  uint32_t func_mask, int use_quantile, double percentile, int64_t n, const int64_t *start, const int64_t *end, float *out)
{
	REF_TRY
	GenomeTrackArrays t;
	t.init_read(chrom_file, chromid);
	std::vector<unsigned> sl(slice, slice + nslice);
	if (slice_func == GenomeTrackArrays::S_QUANTILE)
		t.set_slice_quantile(slice_percentile, 1000000, 100000, 100000, sl);
	else
		t.set_slice_function((GenomeTrackArrays::SliceFunctions)slice_func, sl);
	register_funcs(t, func_mask, use_quantile, 1000000);
	for (int64_t i = 0; i < n; ++i) {
		GInterval iv(chromid, start[i], end[i], 0);
		t.read_interval(iv);
		collect(t, func_mask, use_quantile, percentile, out + i * NUM_COLS);
	}
	return 0;
	REF_CATCH
}

extern "C" int ref_dense_info(const char *chrom_file, int chromid, uint32_t *bin_size, int64_t *num_samples)
{
	REF_TRY
	GenomeTrackFixedBin t;
	t.init_read(chrom_file, chromid);
	*bin_size = t.get_bin_size();
	*num_samples = t.get_num_samples();
	return 0;
	REF_CATCH
}

extern "C" int ref_sparse_eval(const char *chrom_file, int chromid, uint32_t func_mask, int use_quantile, double percentile,
							   int64_t n, const int64_t *start, const int64_t *end, float *out /* n x NUM_COLS */)
{
	REF_TRY
	GenomeTrackSparse t;
	t.init_read(chrom_file, chromid);
	register_funcs(t, func_mask, use_quantile, 1000000);
	for (int64_t i = 0; i < n; ++i) {
		GInterval iv(chromid, start[i], end[i], 0);
		t.read_interval(iv);
		collect(t, func_mask, use_quantile, percentile, out + i * NUM_COLS);
	}
	return 0;
	REF_CATCH
}

// gquantiles' own object and loop, src/GenomeTrackQuantiles.cpp:165-173 + calc_medians' exact branch (:100-105):
// StreamPercentiler<double> fed with float-narrowed non-NaN values, one get_percentile per requested quantile.
extern "C" int ref_gquantiles(const double *vals, int64_t n, int nq, const double *pct, double *out)
{
	REF_TRY
	StreamPercentiler<double> sp(n + 16, 1000, 1000);
	for (int64_t i = 0; i < n; ++i) {
		float val = vals[i];
		if (!std::isnan(val))
			sp.add(val, NULL);
	}
	for (int k = 0; k < nq; ++k) {
		bool est;
		out[k] = sp.stream_size() ? sp.get_percentile(pct[k], est) : std::numeric_limits<double>::quiet_NaN();
	}
	return 0;
	REF_CATCH
}

// Exact-regime percentile (n <= reservoir), src/StreamPercentiler.h:191-217.
extern "C" int ref_percentile(const float *vals, int64_t n, double percentile, float *out)
{
	REF_TRY
	StreamPercentiler<float> sp;
	sp.init(n + 16, 1000, 1000);
	for (int64_t i = 0; i < n; ++i)
		sp.add(vals[i], NULL);
	bool est;
	*out = n ? sp.get_percentile(percentile, est) : std::numeric_limits<float>::quiet_NaN();
	return 0;
	REF_CATCH
}

extern "C" int ref_val2bin(const double *breaks, int nbreaks, int include_lowest, int right, const double *vals, int64_t n, int *out)
{
	REF_TRY
	BinFinder bf(breaks, (unsigned)nbreaks, include_lowest != 0, right != 0);
	for (int64_t i = 0; i < n; ++i)
		out[i] = bf.val2bin(vals[i]);
	return 0;
	REF_CATCH
}

static void make_chromkey(GenomeChromKey &ck, int nchroms, const char *const *names, const int64_t *sizes)
{
	for (int i = 0; i < nchroms; ++i)
		ck.add_chrom(names[i], (uint64_t)sizes[i]);
}

// PWM vtrack scoring. pssm = L x 4 row-major probabilities in A,C,G,T order as the user supplies them;
// prepared as PWMParams::parse does (src/TrackExpressionParams.h:96-143): DnaProbVec(pa,pc,pg,pt) per row,
// set_bidirect, add_dirichlet_prior((float)prior). mode: 0=pwm (TOTAL_LIKELIHOOD), 1=pwm.max, 2=pwm.max.pos, 3=pwm.count.
extern "C" int ref_pwm_eval(const char *groot, int nchroms, const char *const *names, const int64_t *sizes, const double *pssm, int L,
							int bidirect, double prior, int extend, int mode, int strand, float score_thresh, int64_t n,
							const int32_t *chromid, const int64_t *start, const int64_t *end, float *out)
{
	REF_TRY
	GenomeChromKey ck;
	make_chromkey(ck, nchroms, names, sizes);
	DnaPSSM p;
	p.resize(L);
	for (int i = 0; i < L; ++i) {
		float pa = pssm[i * 4 + 0], pc = pssm[i * 4 + 1], pg = pssm[i * 4 + 2], pt = pssm[i * 4 + 3];
		p[i] = DnaProbVec(pa, pc, pg, pt);
	}
	p.set_bidirect(bidirect != 0);
	if (prior > 0)
		p.add_dirichlet_prior((float)prior);
	PWMScorer scorer(p, std::string(groot), extend != 0, (PWMScorer::ScoringMode)mode, (char)strand, std::vector<float>(), 1, score_thresh);
	for (int64_t i = 0; i < n; ++i) {
		GInterval iv(chromid[i], start[i], end[i], 0);
		out[i] = scorer.score_interval(iv, ck);
	}
	return 0;
	REF_CATCH
}

// The same with spatial weights (PWMParams spat_factor / spat_bin). fresh != 0: a new PWMScorer per interval, i.e. no sliding-window
// state carried from one interval to the next.
extern "C" int ref_pwm_eval_spat(const char *groot, int nchroms, const char *const *names, const int64_t *sizes, const double *pssm, int L,
								 int bidirect, double prior, int extend, int mode, int strand, float score_thresh, const float *spat, int nspat,
								 int spat_bin, int fresh, int64_t n, const int32_t *chromid, const int64_t *start, const int64_t *end, float *out)
{
	REF_TRY
	GenomeChromKey ck;
	make_chromkey(ck, nchroms, names, sizes);
	DnaPSSM p;
	p.resize(L);
	for (int i = 0; i < L; ++i) {
		float pa = pssm[i * 4 + 0], pc = pssm[i * 4 + 1], pg = pssm[i * 4 + 2], pt = pssm[i * 4 + 3];
		p[i] = DnaProbVec(pa, pc, pg, pt);
	}
	p.set_bidirect(bidirect != 0);
	if (prior > 0)
		p.add_dirichlet_prior((float)prior);
	std::vector<float> sp(spat, spat + nspat);
	std::unique_ptr<PWMScorer> scorer;
	for (int64_t i = 0; i < n; ++i) {
		if (!scorer || fresh)
			scorer.reset(new PWMScorer(p, std::string(groot), extend != 0, (PWMScorer::ScoringMode)mode, (char)strand, sp, spat_bin, score_thresh));
		GInterval iv(chromid[i], start[i], end[i], 0);
		out[i] = scorer->score_interval(iv, ck);
	}
	return 0;
	REF_CATCH
}

// kmer.count / kmer.frac vtracks: KmerCounter::score_interval (src/KmerCounter.cpp:38-101). mode: 0 = count (SUM), 1 = fraction;
// strand: 0 = both, 1, -1.
extern "C" int ref_kmer_eval(const char *groot, int nchroms, const char *const *names, const int64_t *sizes, const char *kmer, int mode,
							 int extend, int strand, int64_t n, const int32_t *chromid, const int64_t *start, const int64_t *end, float *out)
{
	REF_TRY
	GenomeChromKey ck;
	make_chromkey(ck, nchroms, names, sizes);
	KmerCounter counter(std::string(kmer), std::string(groot), (KmerCounter::CountMode)mode, extend != 0, (char)strand);
	for (int64_t i = 0; i < n; ++i) {
		GInterval iv(chromid[i], start[i], end[i], 0);
		out[i] = counter.score_interval(iv, ck);
	}
	return 0;
	REF_CATCH
}

// masked.count / masked.frac vtracks: MaskedBpCounter::score_interval (src/MaskedBpCounter.cpp:19-64). mode: 0 = count, 1 = fraction.
extern "C" int ref_masked_eval(const char *groot, int nchroms, const char *const *names, const int64_t *sizes, int mode, int64_t n,
							   const int32_t *chromid, const int64_t *start, const int64_t *end, float *out)
{
	REF_TRY
	GenomeChromKey ck;
	make_chromkey(ck, nchroms, names, sizes);
	MaskedBpCounter counter(std::string(groot), (MaskedBpCounter::CountMode)mode);
	for (int64_t i = 0; i < n; ++i) {
		GInterval iv(chromid[i], start[i], end[i], 0);
		out[i] = counter.score_interval(iv, ck);
	}
	return 0;
	REF_CATCH
}

// The log-probability table the reference scores with (after normalisation + prior), L x 4 floats.
extern "C" int ref_pssm_logp(const double *pssm, int L, double prior, float *out)
{
	REF_TRY
	DnaPSSM p;
	p.resize(L);
	for (int i = 0; i < L; ++i)
		p[i] = DnaProbVec((float)pssm[i * 4 + 0], (float)pssm[i * 4 + 1], (float)pssm[i * 4 + 2], (float)pssm[i * 4 + 3]);
	if (prior > 0)
		p.add_dirichlet_prior((float)prior);
	for (int i = 0; i < L; ++i)
		for (int c = 0; c < 4; ++c)
			out[i * 4 + c] = p[i].get_log_prob_from_code(c);
	return 0;
	REF_CATCH
}

static void fill_intervals(GIntervals &gi, int64_t n, const int32_t *chromid, const int64_t *start, const int64_t *end)
{
	gi.reserve(n);
	for (int64_t i = 0; i < n; ++i)
		gi.push_back(GInterval(chromid[i], start[i], end[i], 0, (void *)(intptr_t)i));
}

// Rows of the fixed-bin iterator over a scope (scope is sorted by the caller as the entry points do,
// src/GenomeTrackExtract.cpp:557-562). Returns the number of rows; writes at most `cap`.
extern "C" int64_t ref_fixedbin_rows(int64_t binsize, int64_t nscope, const int32_t *chromid, const int64_t *start, const int64_t *end,
									 int64_t cap, int32_t *ochrom, int64_t *ostart, int64_t *oend, int64_t *oscope_idx)
{
	REF_TRY
	GIntervals scope;
	fill_intervals(scope, nscope, chromid, start, end);
	scope.sort();
	TrackExpressionFixedBinIterator it;
	int64_t r = 0;
	for (it.begin(binsize, scope); !it.isend(); it.next()) {
		if (r < cap) {
			ochrom[r] = it.last_interval().chromid;
			ostart[r] = it.last_interval().start;
			oend[r] = it.last_interval().end;
			oscope_idx[r] = (int64_t)(intptr_t)it.last_scope_interval().udata;
		}
		++r;
	}
	return r;
	REF_CATCH
}

// Rows of the intervals iterator: iterator intervals sorted + overlap-unified with touching intervals kept apart
// (src/TrackExpressionScanner.cpp:934-954), intersected with the sorted scope.
extern "C" int64_t ref_intervals_rows(int64_t niter, const int32_t *ichrom, const int64_t *istart, const int64_t *iend, int64_t nscope,
									  const int32_t *schrom, const int64_t *sstart, const int64_t *send, int64_t cap, int32_t *ochrom,
									  int64_t *ostart, int64_t *oend, int64_t *oscope_idx)
{
	REF_TRY
	GIntervals iter, scope;
	fill_intervals(iter, niter, ichrom, istart, iend);
	fill_intervals(scope, nscope, schrom, sstart, send);
	iter.sort();
	iter.unify_overlaps(false);
	scope.sort();
	TrackExpressionIntervals1DIterator it;
	int64_t r = 0;
	for (it.begin(iter, scope); !it.isend(); it.next()) {
		if (r < cap) {
			ochrom[r] = it.last_interval().chromid;
			ostart[r] = it.last_interval().start;
			oend[r] = it.last_interval().end;
			oscope_idx[r] = (int64_t)(intptr_t)it.last_scope_interval().udata;
		}
		++r;
	}
	return r;
	REF_CATCH
}

// sort + unify_overlaps(unify_touching) of an interval set, src/GIntervals.cpp:59-74. Returns new count, in place.
extern "C" int64_t ref_unify(int64_t n, int32_t *chromid, int64_t *start, int64_t *end, int unify_touching)
{
	REF_TRY
	GIntervals gi;
	fill_intervals(gi, n, chromid, start, end);
	gi.sort();
	gi.unify_overlaps(unify_touching != 0);
	for (size_t i = 0; i < gi.size(); ++i) {
		chromid[i] = gi[i].chromid;
		start[i] = gi[i].start;
		end[i] = gi[i].end;
	}
	return (int64_t)gi.size();
	REF_CATCH
}
