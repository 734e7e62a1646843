// GpuTrackExprScanner.h -- the host side of the drop-in, in the reference's own language.
//
// Mirrors the part of TrackExprScanner (src/TrackExpressionScanner.h:21-63, .cpp:215-379) that the .Call entry points of
// the hot path drive: begin() / next() / isend() / last_real() / last_interval1d() / last_scope_idx(). Where the
// reference fills a 1000-row buffer by calling read_interval() per row and per track (eval_next, .cpp:280-379), this
// scanner fills a buffer of up to 2^20 rows with ONE libmishagpu call per (source, iterator-modifier) group -- the
// grouping add_track_n_imdf does (src/TrackExpressionVars.cpp:370-399) -- and the per-row accessors read that buffer.
// GPU-aware callers skip the per-row loop altogether and take whole columns (batch_*).
//
// R is not available in this image, so the only expressions this scanner evaluates itself are bare variable names
// (the m_eval_exprs[i] == R_NilValue case of the reference, src/TrackExpressionScanner.h:143-158); anything else is
// R's job on the returned columns, exactly as before. Virtual tracks are registered with vtrack_*(), the C++ face of
// gvtrack.create + gvtrack.iterator (R/vtrack.R; Iterator_modifier1D src/TrackExpressionVars.h:395-402).
//
// Errors: every failed mg_* call throws MishaGpuException carrying mg_last_error() -- the TGLException that verror()
// raises in the reference (src/rdbutils.cpp:720-733); there is no CPU fallback.
#pragma once
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/mishagpu.h"

namespace mishagpu {

struct MishaGpuException : std::runtime_error {
	using std::runtime_error::runtime_error;
};

// GInterval's fields (src/GInterval.h): chromid, [start, end)
struct Interval {
	int32_t chromid;
	int64_t start, end;
};

struct VTrack {
	enum Source { DENSE, SPARSE, COVERAGE, PWM, KMER };
	Source source = DENSE;
	const mg_dense *dense = nullptr;
	const mg_sparse *sparse = nullptr;
	const mg_ivset *ivset = nullptr;
	// sequence vtracks (pwm, pwm.max, pwm.max.pos, pwm.count): PWMParams, src/TrackExpressionParams.h:74-240
	const mg_seq *seq = nullptr;
	std::vector<float> logp; // L x 4 log-probabilities (mg_pssm_logp)
	int pwm_mode = MG_PWM_TOTAL, bidirect = 1, extend = 1, strand = 1;
	float score_thresh = 0;
	// kmer.count / kmer.frac (KmerCounter, src/KmerCounter.cpp): mg_kmer_mode, the k-mer; extend / strand above (strand 0 = both)
	int kmer_mode = MG_KMER_COUNT;
	std::string kmer;
	int func = MG_AVG;   // enum mg_func
	double param = 0;    // percentile of MG_QUANTILE
	int64_t sshift = 0, eshift = 0;
};

class GpuTrackExprScanner {
public:
	explicit GpuTrackExprScanner(mg_ctx *ctx, int64_t buf_rows = int64_t(1) << 20);
	~GpuTrackExprScanner();
	GpuTrackExprScanner(const GpuTrackExprScanner &) = delete;
	GpuTrackExprScanner &operator=(const GpuTrackExprScanner &) = delete;

	// gvtrack.create(name, src, func, params) followed by gvtrack.iterator(name, sshift=, eshift=)
	void vtrack_dense(const std::string &name, const mg_dense *track, int func, double param = 0, int64_t sshift = 0, int64_t eshift = 0);
	void vtrack_sparse(const std::string &name, const mg_sparse *track, int func, double param = 0, int64_t sshift = 0, int64_t eshift = 0);
	void vtrack_coverage(const std::string &name, const mg_ivset *src, int64_t sshift = 0, int64_t eshift = 0);
	// gvtrack.create(name, NULL, "pwm" | "pwm.max" | "pwm.max.pos" | "pwm.count", list(pssm = , bidirect = , prior = , extend = , strand = ,
	// score.thresh = )): pssm = L x 4 probabilities (A, C, G, T), turned into the reference's log table by mg_pssm_logp
	void vtrack_pwm(const std::string &name, const mg_seq *seq, int mode, const std::vector<double> &pssm, double prior = 0.01, bool bidirect = true,
					bool extend = true, int strand = 1, float score_thresh = 0, int64_t sshift = 0, int64_t eshift = 0);
	// gvtrack.create(name, NULL, "kmer.count" | "kmer.frac", list(kmer = , extend = TRUE, strand = 0)), .vtrack_params_kmer R/vtrack.R:292-329
	void vtrack_kmer(const std::string &name, const mg_seq *seq, int mode, const std::string &kmer, bool extend = true, int strand = 0,
					 int64_t sshift = 0, int64_t eshift = 0);
	void vtrack_rm(const std::string &name);

	// Fixed-bin iterator (iterator = binsize) or intervals iterator over `scope`. The scope is sorted by (chromid, start) as
	// every entry point does before scanning (src/GenomeTrackExtract.cpp:557-562); unify_scope = what gscreen / gsummary / gdist
	// additionally do (unify_overlaps, src/TrackExpressionScanner.cpp:934-954). Returns !isend().
	bool begin(const std::vector<std::string> &track_exprs, const std::vector<Interval> &scope, int64_t iterator_binsize,
			   bool unify_scope = false);
	bool begin(const std::vector<std::string> &track_exprs, const std::vector<Interval> &scope, const std::vector<Interval> &iterator_intervals,
			   bool unify_scope = false);
	bool next();
	bool next_batch(); // skip the rest of the current buffer (whole-column consumers)
	bool isend() const { return m_isend; }

	double last_real(int track_expr_idx) const { return m_cols[track_expr_idx][m_buf_idx]; }
	Interval last_interval1d() const { return Interval{m_chrom[m_buf_idx], m_start[m_buf_idx], m_end[m_buf_idx]}; }
	// index of the scope interval in the ORDER THE CALLER PASSED (intervalID - 1 of gextract, src/IntervalConverter.cpp:357)
	int64_t last_scope_idx() const { return m_scope_idx[m_buf_idx]; }

	// the current buffer as whole columns (valid until the next next() that refills)
	int64_t total_rows() const { return m_nrows; }
	int64_t batch_first() const { return m_buf_first; }
	int64_t batch_size() const { return m_buf_count; }
	const double *batch_real(int track_expr_idx) const { return m_cols[track_expr_idx]; }
	const int32_t *batch_chroms() const { return m_chrom; }
	const int64_t *batch_starts() const { return m_start; }
	const int64_t *batch_ends() const { return m_end; }
	const int64_t *batch_scope_idx() const { return m_scope_idx; }
	const mg_rows *rows() const { return m_rows; }
	const std::vector<std::string> &get_track_exprs() const { return m_track_exprs; }

private:
	void check(int rc) const;
	void prepare_scope(const std::vector<Interval> &scope, bool unify_scope);
	bool start(const std::vector<std::string> &track_exprs);
	void release();
	bool fill(int64_t first);

	mg_ctx *m_ctx;
	int64_t m_buf_rows;
	std::map<std::string, VTrack> m_vtracks;
	std::vector<std::string> m_track_exprs;
	std::vector<const VTrack *> m_vars;
	std::vector<Interval> m_scope;     // sorted (and possibly unified)
	std::vector<int64_t> m_scope_orig; // position of each sorted scope interval in the caller's order
	mg_rows *m_rows = nullptr;
	int64_t m_nrows = 0, m_buf_first = 0, m_buf_count = 0, m_buf_idx = 0;
	bool m_isend = true;
	// device + pinned host buffers, one row slot per buffer row
	std::vector<double *> m_dcols, m_cols;
	int32_t *m_dchrom = nullptr, *m_chrom = nullptr;
	int64_t *m_dstart = nullptr, *m_dend = nullptr, *m_dscope = nullptr, *m_start = nullptr, *m_end = nullptr, *m_scope_idx = nullptr;
};

// ---- the entry points' result assembly over the scanner ------------------------------------------------------------------
// C_gextract (src/GenomeTrackExtract.cpp:518-1126): chrom / start / end, one double column per expression, intervalID.
struct ExtractResult {
	std::vector<int32_t> chrom;
	std::vector<int64_t> start, end;
	std::vector<std::vector<double>> values;
	std::vector<int32_t> interval_id; // 1-based row of the caller's scope
};
// per_row = drive the scanner exactly like the reference's loop (next() / last_real()); otherwise copy whole columns
ExtractResult gextract(GpuTrackExprScanner &scanner, const std::vector<std::string> &track_exprs, const std::vector<Interval> &scope,
					   int64_t iterator_binsize, const std::vector<Interval> *iterator_intervals = nullptr, bool per_row = false);

// gtracksummary (src/GenomeTrackSummary.cpp:119-172) of one dense-track expression: Total intervals, NaN intervals, Min, Max, Sum, Mean, Std dev
void gsummary_dense(mg_ctx *ctx, const mg_dense *track, const std::vector<Interval> &scope, int64_t iterator_binsize, int func, double param,
					int64_t sshift, int64_t eshift, double out[7]);

// ---- several GPUs behind this one thread (mg_group_*, include/mishagpu.h) --------------------------------------------------------
// What gextract_multitask / gtracksummary_multitask / gtrackdist_multitask / gquantiles_multitask do with forked children
// (src/GenomeTrackExtract.cpp:1128-1560, src/GenomeTrackSummary.cpp:174-226, src/GenomeTrackDistribution.cpp:82-150,
// src/GenomeTrackQuantiles.cpp:278-400), for dense-track vtracks: the session owns the group and its staged tracks, the calls take
// the scope the way the .Call entry points get it and return what the parent process assembles.
class GpuGroupSession {
public:
	GpuGroupSession(const std::vector<int> &devices, const std::vector<int64_t> &chrom_sizes);
	~GpuGroupSession();
	GpuGroupSession(const GpuGroupSession &) = delete;
	GpuGroupSession &operator=(const GpuGroupSession &) = delete;
	mg_group *group() const { return m_g; }
	int size() const { return mg_group_size(m_g); }
	// a dense track of the group; every chromosome is staged on its owner
	mg_gdense *dense_create(uint32_t bin_size);
	void dense_stage(mg_gdense *t, int chromid, const float *bins, int64_t nbins);
	// gextract with the scope as the (intervals) iterator: one column per {func, param}, all under the same sshift / eshift; rows come
	// back in the scope's sorted order with intervalID = 1-based row of the caller's scope, like ExtractResult of the scanner
	ExtractResult gextract(const mg_gdense *t, const std::vector<std::pair<int, double>> &funcs, int64_t sshift, int64_t eshift,
						   const std::vector<Interval> &scope);
	// gsummary / gdist / gquantiles of one function over the fixed-bin iterator of the (sorted, unified) scope
	void gsummary(const mg_gdense *t, const std::vector<Interval> &scope, int64_t iterator_binsize, int func, double param, int64_t sshift,
				  int64_t eshift, double out[7]);
	std::vector<uint64_t> gdist(const mg_gdense *t, const std::vector<Interval> &scope, int64_t iterator_binsize, int func, double param,
								int64_t sshift, int64_t eshift, const std::vector<double> &breaks, bool include_lowest);
	std::vector<double> gquantiles(const mg_gdense *t, const std::vector<Interval> &scope, int64_t iterator_binsize, int func, double param,
								   int64_t sshift, int64_t eshift, const std::vector<double> &percentiles);

private:
	void check(int rc) const;
	mg_group *m_g = nullptr;
	std::vector<mg_gdense *> m_tracks;
};

} // namespace mishagpu
