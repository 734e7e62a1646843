/* mishagpu.h -- C ABI of the B200-native track-extraction / aggregation path (libmishagpu.so).
 *
 * This is boundary B2 of SURVEY.md 8(b): what misha's host C++ (the .Call entry points, TrackExprScanner,
 * TrackExpressionVars) calls INSTEAD of its per-row CPU arithmetic. Plain C: opaque handles, pointers and
 * sizes; no SEXP, no C++ types, no torch types. Every function returns 0 on success and a non-zero status on
 * failure (message via mg_last_error); nothing throws or longjmps across the boundary. There is no CPU
 * fallback behind any entry point.
 *
 * Reference interfaces replaced (paths under the misha checkout, v5.11.20):
 *   staging      GenomeTrackFixedBin::init_read        src/GenomeTrackFixedBin.cpp:1129   -> mg_dense_stage
 *                GenomeTrackSparse::read_file_into_mem src/GenomeTrackSparse.cpp:162-235  -> mg_sparse_stage
 *                add_vtrack_var_src_interv             src/TrackExpressionVars.cpp:1292-1301 -> mg_ivset_stage
 *                GenomeSeqFetch::read_interval         src/GenomeSeqFetch.cpp:50-175      -> mg_seq_stage
 *   rows         TrackExpressionFixedBinIterator::next src/TrackExpressionFixedBinIterator.cpp:29-62 -> mg_rows_fixedbin
 *                TrackExpressionIntervals1DIterator    src/TrackExpressionIntervals1DIterator.cpp:21-85 -> mg_rows_intervals
 *   per-row work TrackExpressionVars::set_vars         src/TrackExpressionVars.cpp:2080-2150:
 *                  Iterator_modifier1D::transform      src/TrackExpressionVars.h:395-402  (sshift/eshift clamp)
 *                  GenomeTrackFixedBin::read_interval  src/GenomeTrackFixedBin.cpp:607-1016 -> mg_dense_window
 *                  GenomeTrackSparse::read_interval    src/GenomeTrackSparse.cpp:237-340  -> mg_sparse_window
 *                  IntervVarProcessor::process_coverage src/IntervVarProcessor.cpp:242-325 -> mg_coverage
 *                  PWMScorer::score_interval           src/PWMScorer.cpp:742-861          -> mg_pwm
 *                  KmerCounter::score_interval         src/KmerCounter.cpp:38-176         -> mg_kmer
 *                  MaskedBpCounter::score_interval     src/MaskedBpCounter.cpp:19-64      -> mg_masked
 *                  global.percentile* lookup           src/TrackVarProcessor.cpp:98-110   -> mg_percentile_lookup
 *   consumers    IntervalSummary::update/merge         src/GenomeTrackSummary.cpp:22-69   -> mg_summary*, mg_summary_finalize
 *                BinsManager::vals2idx + counters      src/BinsManager.h:53-72, src/GenomeTrackDistribution.cpp:121-141 -> mg_hist*
 *                C_gscreen run merge                   src/GenomeTrackScreener.cpp:195-220 -> mg_screen_merge
 *                C_gquantiles                          src/GenomeTrackQuantiles.cpp:124-262 -> mg_quantiles
 *                gintervals_quantiles / _summary       src/GenomeTrackQuantiles.cpp:612-860, src/GenomeTrackSummary.cpp:248-431
 *                                                                                          -> mg_quantiles_segmented, mg_summary_segmented
 *                gbins_summary / gbins_quantiles       src/GenomeTrackSummary.cpp:433-506, src/GenomeTrackQuantiles.cpp:1199-1330
 *                                                                                          -> mg_bins_summary, mg_bins_quantiles
 *
 * Conventions
 *   - One mg_ctx per process and GPU; calls on one ctx are serialised by the caller (R is single-threaded).
 *   - "d_" pointers are device memory on the ctx's GPU, "h_" pointers host memory (pinned recommended).
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream). Device entry points only
 *     enqueue work; the caller synchronises (mg_sync) before reading results.
 *   - Value columns are double (R numeric) holding float32-rounded values exactly as the reference's float
 *     members widened to double (src/GenomeTrack1D.h:58-103).
 */
#ifndef MISHAGPU_H_
#define MISHAGPU_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MG_ABI_VERSION 1

typedef struct mg_ctx mg_ctx;
typedef struct mg_dense mg_dense;   /* dense fixed-bin track staged in HBM (all chromosomes) */
typedef struct mg_sparse mg_sparse; /* sparse track staged in HBM */
typedef struct mg_ivset mg_ivset;   /* sorted, overlap-unified interval set (coverage source) */
typedef struct mg_seq mg_seq;       /* 2-bit packed genome sequence + validity mask */
typedef struct mg_rows mg_rows;     /* iterator rows resident on the device (explicit, or implicit fixed-bin) */

/* Virtual-track functions (gvtrack.create `func`, R/vtrack.R; Track_var::Val_func src/TrackExpressionVars.h:118-163). */
enum mg_func {
	MG_AVG = 0,      /* "avg"; a bare dense track name */
	MG_SUM = 1,      /* "sum" */
	MG_MIN = 2,      /* "min" */
	MG_MAX = 3,      /* "max" */
	MG_STDDEV = 4,   /* "stddev" */
	MG_QUANTILE = 5, /* "quantile", param = percentile */
	MG_NEAREST = 6,  /* "nearest" (dense: == avg) */
	MG_SIZE = 7,     /* "size": number of non-NaN values */
	MG_EXISTS = 8,   /* "exists" */
	/* the order-dependent functions of the generic read_interval path (src/GenomeTrackFixedBin.cpp:881-939,
	 * src/GenomeTrackSparse.h:143-259). Positions are absolute coordinates -- dense: max(bin start, window start), sparse: the
	 * record's start -- or, with param != 0, relative to the (shifted, clamped) window start as the "*.pos.relative" vtrack
	 * functions return them (src/TrackVarProcessor.cpp:113-173). "sample" / "sample.pos.*" draw from R's RNG: not offered. */
	MG_LSE = 9,        /* "lse": log(sum(exp(v))) */
	MG_FIRST = 10,     /* "first": first non-NaN value */
	MG_LAST = 11,      /* "last" */
	MG_FIRST_POS = 12, /* "first.pos.abs" / "first.pos.relative" */
	MG_LAST_POS = 13,  /* "last.pos.abs" / "last.pos.relative" */
	MG_MIN_POS = 14,   /* "min.pos.abs" / "min.pos.relative": position of the first bin / record holding the minimum */
	MG_MAX_POS = 15,   /* "max.pos.abs" / "max.pos.relative" */
	MG_NUM_FUNCS = 16
};
/* Path flag, OR-ed into a function id of a dense-track call. The reference's dense reader answers MG_SIZE, MG_EXISTS and MG_LSE   */
/* through one of two code paths, chosen by WHICH OTHER functions share the track object (classify_fast_path_mode,                */
/* src/GenomeTrackFixedBin.cpp:90-147), and the two disagree:                                                                     */
/*   reducer path (only sum / lse / exists / size registered; the default here): a window inside one NaN bin has size = exists = 0 */
/*       (assign_single_value, :231-240); lse = the double max-trick sum (RunningLogSumExp);                                      */
/*   generic path (any other function or a quantile next to them): a window inside ONE bin reports size = exists = 1 whatever the  */
/*       bin holds (assign_single_bin_value, :150-180); lse accumulates pairwise in float in bin order (lse_accumulate,           */
/*       src/GenomeTrack1D.h:18-28).                                                                                              */
/* The host knows which functions it registered on the track object it replaces and says so per function: MG_SIZE | MG_FUNC_GENERIC. */
#define MG_FUNC_GENERIC 0x100

/* PWM vtrack modes (PWMScorer::ScoringMode src/PWMScorer.h:16-22). */
enum mg_masked_mode { MG_MASKED_COUNT = 0 /* "masked.count" */, MG_MASKED_FRAC = 1 /* "masked.frac" */ };
enum mg_kmer_mode { MG_KMER_COUNT = 0 /* "kmer.count" */, MG_KMER_FRAC = 1 /* "kmer.frac" */ };
enum mg_pwm_mode { MG_PWM_TOTAL = 0 /* "pwm" */, MG_PWM_MAX = 1 /* "pwm.max" */, MG_PWM_MAX_POS = 2 /* "pwm.max.pos" */, MG_PWM_COUNT = 3 /* "pwm.count" */ };

/* ---- context ------------------------------------------------------------------------------------------- */
int mg_abi_version(void);
/* chrom_sizes[chromid] as GenomeChromKey (src/GenomeChromKey.h) holds them; fixed for the ctx lifetime. */
int mg_init(int device, int nchroms, const int64_t *chrom_sizes, mg_ctx **out);
void mg_shutdown(mg_ctx *ctx);
const char *mg_last_error(const mg_ctx *ctx); /* ctx may be NULL: error of the last failed mg_init in this thread */
int mg_sync(mg_ctx *ctx, void *stream);
/* Tuning / ablation switches (same results whatever their value, except pwm_strict's last float bits):
 *   "force_sweep" = 1    dense min / max / quantile through the warp-per-row sweep kernel, sparse / coverage through the
 *                        per-row search kernels, instead of the index structures / merge kernels (tests, profiling)
 *   "block_wm" = 0       quantiles from the chromosome-wide wavelet matrix only (default 1: block-local matrices for windows
 *                        of <= 1024 bins)
 *   "stage_slots" = n    block records a row-query thread block stages in shared memory for its quantile walks (default 5,
 *                        0 = every walk reads its record through L1; at most 12)
 *   "cov_raster" = 0     coverage over fixed-bin rows by per-row cursors instead of rasterising the records into the rows
 *   "merge_rows" = 8     sparse / coverage merge kernels always at 8 rows per thread (default 0: 16 for fixed-bin rows whose
 *                        thread blocks are expected to hold few records)
 *   "hist_red" = 0       gdist over a dense track counts with a 16-bit load / add / store instead of one red.shared.add
 *   "prefetch" = 0       no L1 prefetch of block records / window lines in the row-query kernels
 *   "pwm_strict" = 1     PWM log-sum-exp with double-precision exp / log rounded to float (default: float expf / logf)
 *   "pipe_chunk_rows"    rows per slot of the host-buffer pipeline (default 3 * 2^17)
 *   "pipe_slots" = n     device slots (one stream each) the chunks of a host batch rotate through (default 3, 2 .. 8)
 *   "tile_kernel" = n    which dense window kernel runs: 3 (default) k_dense_thin when two or more function groups (sums,
 *                        extrema, quantiles) share a scan or a quantile is asked for, the per-row kernels otherwise; 4 k_dense_thin always; 1 / 2
 *                        k_dense_tile (sums to the reference's bits); 5 / 6 k_dense_pipe; 0 the per-row kernels
 *   "inline_table" = 0   the tile kernels load chromosome sizes / array bases from the tables in HBM instead of their parameters
 *   "hbm_budget_mb" = n  what the dense tracks of this ctx may hold; least-recently-used index groups are dropped beyond it
 *   "zero_copy" = 1      mg_h_* kernels read / write pinned host buffers directly instead of the DMA pipeline (default 0:
 *                        measured slower on B200)
 *   "l2_fetch_granularity" = 32 / 64 / 128   cudaLimitMaxL2FetchGranularity */
int mg_set_option(mg_ctx *ctx, const char *name, int64_t value);
/* number of kernels this ctx has launched since mg_init (bench.py's gpu_launches claim) */
int64_t mg_launch_count(const mg_ctx *ctx);
/* device properties of the ctx's GPU: SM count, bytes of HBM, compute capability major*10+minor */
int mg_device_info(const mg_ctx *ctx, int *sm_count, int64_t *hbm_bytes, int *cc);

/* plumbing for hosts that have no other CUDA binding (the Python mirror uses these; R/C++ hosts may too) */
int mg_dev_alloc(mg_ctx *ctx, int64_t bytes, void **d_ptr);
int mg_dev_free(mg_ctx *ctx, void *d_ptr);
int mg_host_alloc(mg_ctx *ctx, int64_t bytes, void **h_ptr); /* pinned */
int mg_host_free(mg_ctx *ctx, void *h_ptr);
int mg_copy_h2d(mg_ctx *ctx, void *d_dst, const void *h_src, int64_t bytes, void *stream);
int mg_copy_d2h(mg_ctx *ctx, void *h_dst, const void *d_src, int64_t bytes, void *stream);
int mg_dev_memset(mg_ctx *ctx, void *d_ptr, int value, int64_t bytes, void *stream);

/* ---- staging ("stages chromosome track blocks into HBM once") ------------------------------------------ */
/* Dense: h_bins = payload of the per-chromosome file after the 4-byte bin-size header (may hold +-inf, which */
/* become NaN on load as in read_bins_bulk). Builds the chromosome's index structures once: group-of-8 prefix  */
/* sums (avg/sum/size), sparse tables and pyramids (min/max), wavelet matrices (quantile) -- DESIGN.md 2.     */
int mg_dense_create(mg_ctx *ctx, uint32_t bin_size, mg_dense **out);
int mg_dense_stage(mg_ctx *ctx, mg_dense *t, int chromid, const float *h_bins, int64_t nbins);
void mg_dense_free(mg_ctx *ctx, mg_dense *t);
int64_t mg_dense_bytes(const mg_dense *t); /* HBM bytes held */

/* Sparse: records sorted by start, non-overlapping, 0 <= start < end (validated; error otherwise).          */
int mg_sparse_create(mg_ctx *ctx, mg_sparse **out);
int mg_sparse_stage(mg_ctx *ctx, mg_sparse *t, int chromid, const int64_t *h_start, const int64_t *h_end, const float *h_val, int64_t n);
void mg_sparse_free(mg_ctx *ctx, mg_sparse *t);

/* Interval set for `coverage`: per chromosome, sorted by start and overlap-unified by the caller.           */
int mg_ivset_create(mg_ctx *ctx, mg_ivset **out);
int mg_ivset_stage(mg_ctx *ctx, mg_ivset *s, int chromid, const int64_t *h_start, const int64_t *h_end, int64_t n);
void mg_ivset_free(mg_ctx *ctx, mg_ivset *s);

/* Sequence: h_ascii = seq/<chrom>.seq bytes (1 per base, mixed case, N = unknown). Packed on the device.    */
int mg_seq_create(mg_ctx *ctx, mg_seq **out);
int mg_seq_stage(mg_ctx *ctx, mg_seq *s, int chromid, const char *h_ascii, int64_t len);
void mg_seq_free(mg_ctx *ctx, mg_seq *s);

/* ---- iterator rows --------------------------------------------------------------------------------------- */
/* Explicit rows uploaded from the host (already in output order). */
int mg_rows_upload(mg_ctx *ctx, int64_t n, const int32_t *h_chrom, const int64_t *h_start, const int64_t *h_end, void *stream, mg_rows **out);
/* Explicit rows over caller-owned device arrays (not copied, not freed). */
int mg_rows_wrap(mg_ctx *ctx, int64_t n, const int32_t *d_chrom, const int64_t *d_start, const int64_t *d_end, mg_rows **out);
/* Implicit fixed-bin iterator over a scope sorted by (chromid,start): rows are never materialised. */
int mg_rows_fixedbin(mg_ctx *ctx, int64_t binsize, int64_t nscope, const int32_t *h_chrom, const int64_t *h_start, const int64_t *h_end, mg_rows **out);
/* Intervals iterator: iterator intervals (sorted, unified with touching kept apart) x sorted scope, computed on */
/* the device; h_ arrays are host memory. */
int mg_rows_intervals(mg_ctx *ctx, int64_t niter, const int32_t *h_ichrom, const int64_t *h_istart, const int64_t *h_iend, int64_t nscope,
					  const int32_t *h_schrom, const int64_t *h_sstart, const int64_t *h_send, void *stream, mg_rows **out);
int64_t mg_rows_count(const mg_rows *r);
/* Coordinates (and the scope index each row came from, -1 for uploaded rows) of rows [first, first+count) into */
/* device arrays; any output may be NULL. */
int mg_rows_coords(mg_ctx *ctx, const mg_rows *r, int64_t first, int64_t count, int32_t *d_chrom, int64_t *d_start, int64_t *d_end,
				   int64_t *d_scope_idx, void *stream);
void mg_rows_free(mg_ctx *ctx, mg_rows *r);

/* ---- per-row vtrack values (device in, device out) --------------------------------------------------------- */
/* For rows [first, first+count): window = clamp(row +- shift); d_out[k][i] = funcs[k] of row first+i.         */
int mg_dense_window(mg_ctx *ctx, const mg_dense *t, const mg_rows *rows, int64_t first, int64_t count, int64_t sshift, int64_t eshift,
					int nfuncs, const int *funcs, const double *params, double *const *d_out, void *stream);
int mg_sparse_window(mg_ctx *ctx, const mg_sparse *t, const mg_rows *rows, int64_t first, int64_t count, int64_t sshift, int64_t eshift,
					 int nfuncs, const int *funcs, const double *params, double *const *d_out, void *stream);
int mg_coverage(mg_ctx *ctx, const mg_ivset *s, const mg_rows *rows, int64_t first, int64_t count, int64_t sshift, int64_t eshift,
				double *d_out, void *stream);
/* h_logp: L x 4 float log-probabilities (A,C,G,T) after normalisation and prior (mg_pssm_logp). strand: 1 / -1. */
int mg_pwm(mg_ctx *ctx, const mg_seq *s, const float *h_logp, int L, int bidirect, int extend, int mode, int strand, float score_thresh,
		   const mg_rows *rows, int64_t first, int64_t count, int64_t sshift, int64_t eshift, double *d_out, void *stream);
/* The same under spatial weights (gvtrack.create(..., "pwm*", list(spat_factor = , spat_bin = )); PWMParams, src/TrackExpressionParams.h,  */
/* PWMScorer's spat_factor / spat_bin, src/PWMScorer.cpp:112-165): factor k weighs the anchors whose index in the scored target lies in    */
/* [k spat_bin, (k + 1) spat_bin), the last factor everything beyond; log(max(1e-30f, factor)) is added to each strand's log-likelihood   */
/* (pwm, pwm.max) or to the strands' union before the threshold (pwm.count, which -- like the reference's scorer -- does not count anchors */
/* past the last bin). Fresh-window semantics. MG_PWM_MAX_POS is refused with weights. nspat = 0: mg_pwm.                                 */
int mg_pwm_spatial(mg_ctx *ctx, const mg_seq *s, const float *h_logp, int L, int bidirect, int extend, int mode, int strand, float score_thresh,
				   const float *h_spat_factor, int nspat, int spat_bin, const mg_rows *rows, int64_t first, int64_t count, int64_t sshift,
				   int64_t eshift, double *d_out, void *stream);
/* kmer.count / kmer.frac vtracks (KmerCounter::score_interval, src/KmerCounter.cpp:38-176; SequenceVarProcessor calls it per  */
/* row like score_interval of the PWM scorer): occurrences of `kmer` (case folded; A, C, G, T only, at most 32 bases) starting   */
/* inside the window, on the forward strand (strand 1), as its reverse complement (strand -1) or both (strand 0); `extend`       */
/* lets a match run up to k - 1 bases past the window's end. MG_KMER_FRAC divides by the (window length - k + 1) starts, twice     */
/* that for both strands, in float as the reference does. Bit-exact. d_out[i] for rows [first, first + count).                    */
int mg_kmer(mg_ctx *ctx, const mg_seq *s, const char *kmer, int mode, int extend, int strand, const mg_rows *rows, int64_t first, int64_t count,
			int64_t sshift, int64_t eshift, double *d_out, void *stream);
/* masked.count / masked.frac vtracks (MaskedBpCounter::score_interval, src/MaskedBpCounter.cpp:19-64): lower-case (soft-masked) */
/* bases of the window -- a popcount over the lower-case bit staged per base by mg_seq_stage; MG_MASKED_FRAC divides by the      */
/* window length in float as the reference does. Bit-exact.                                                                     */
int mg_masked(mg_ctx *ctx, const mg_seq *s, int mode, const mg_rows *rows, int64_t first, int64_t count, int64_t sshift, int64_t eshift,
			  double *d_out, void *stream);
/* PSSM rows (L x 4 probabilities, A,C,G,T) -> the float log table the reference scores with                    */
/* (src/PWMScorer.cpp:1452-1459, src/DnaPSSM.cpp:77-106,703-719). Host arithmetic, float, bit-exact.            */
int mg_pssm_logp(const double *h_pssm, int L, double prior, float *h_logp);

/* ---- consumers: gsummary / gdist / gscreen ------------------------------------------------------------------ */
/* Partial summary = {num_bins, num_non_nan, sum, min, max, sum_sq} (IntervalSummary's six doubles), ACCUMULATED  */
/* into d_partial (initialise with mg_summary_init). Sums are order-dependent to ~1 ulp; counts/min/max exact.   */
int mg_summary_init(mg_ctx *ctx, double *d_partial, void *stream);
int mg_summary(mg_ctx *ctx, const double *d_vals, int64_t n, double *d_partial, void *stream);
/* fused: summary of one dense-track function over rows, values never leave the chip */
int mg_dense_summary(mg_ctx *ctx, const mg_dense *t, const mg_rows *rows, int64_t first, int64_t count, int64_t sshift, int64_t eshift,
					 int func, double param, double *d_partial, void *stream);
/* host arithmetic of gtracksummary's result vector (src/GenomeTrackSummary.cpp:148-155): 6 partials -> 7 outputs */
void mg_summary_finalize(const double partial[6], double out[7]);
/* the same for n records (gintervals.summary / gbins.summary result rows): partials[6 n] -> out[7 n] */
void mg_summary_finalize_rows(const double *partials, int64_t n, double *out);

/* N-D histogram (gdist). h_breaks = concatenated break vectors, nbreaks[d] each; d_counts has prod(nbreaks[d]-1)  */
/* uint64 cells, first dimension fastest; counts are ACCUMULATED (zero them first). Bit-exact bins.               */
int mg_hist(mg_ctx *ctx, int ndim, const double *const *d_cols, int64_t n, const double *h_breaks, const int *nbreaks, int include_lowest,
			uint64_t *d_counts, void *stream);
int mg_dense_hist(mg_ctx *ctx, const mg_dense *t, const mg_rows *rows, int64_t first, int64_t count, int64_t sshift, int64_t eshift,
				  int func, double param, const double *h_breaks, int nbreaks, int include_lowest, uint64_t *d_counts, void *stream);

/* gquantiles (C_gquantiles, src/GenomeTrackQuantiles.cpp:124-262) over a device column of row values: every value is   */
/* narrowed to float, NaN rows are dropped, and quantile q of the N kept values is s[floor(i)] (1 - w) + s[ceil(i)] w in   */
/* double, i = (N - 1) clamp(q, 0, 1) -- the reference's exact regime. Selection is exact for any N: where the reference   */
/* switches to R-RNG reservoir sampling beyond gmax.data.size values, this call still returns the exact quantiles.        */
/* h_out[nq] (host), *h_nkept = N (may be NULL). Synchronises the stream. At most 64 quantiles per call.                  */
int mg_quantiles(mg_ctx *ctx, const double *d_vals, int64_t n, int nq, const double *h_percentiles, double *h_out, int64_t *h_nkept,
				 void *stream);

/* gintervals.quantiles (gintervals_quantiles, src/GenomeTrackQuantiles.cpp:612-860, 1-D): the reference resets one          */
/* StreamPercentiler<double> at every scope interval (:726-741) and reads the percentiles with calc_medians (:95-120). Here    */
/* the value column of ALL rows stays in HBM and segment s = rows [h_seg_first[s], h_seg_first[s+1]) is one scope interval's  */
/* run of rows (nseg + 1 ascending host offsets). Segments up to 16384 rows are sorted in shared memory, one warp or one     */
/* thread block each; longer ones are radix-selected by one thread block each: at most three launches. Same arithmetic as  */
/* mg_quantiles; h_out[k * nseg + s] = quantile k of segment s (NaN for a segment without values), h_nkept[s] (may be NULL)  */
/* its number of non-NaN values. Synchronises the stream. At most 64 quantiles per call.                                      */
int mg_quantiles_segmented(mg_ctx *ctx, const double *d_vals, int64_t nseg, const int64_t *h_seg_first, int nq, const double *h_percentiles,
						   double *h_out, int64_t *h_nkept, void *stream);

/* gintervals.summary (gintervals_summary, src/GenomeTrackSummary.cpp:248-431): one IntervalSummary per scope interval. */
/* Segments as in mg_quantiles_segmented; h_partials[6 * s ..] = segment s' six doubles (mg_summary_finalize turns them   */
/* into the result row). One launch, a warp per segment; segments beyond 2^18 rows use mg_summary's kernels. Row values   */
/* are narrowed to float first (`float val = scanner.last_real(0)`). Synchronises the stream.                            */
int mg_summary_segmented(mg_ctx *ctx, const double *d_vals, int64_t nseg, const int64_t *h_seg_first, double *h_partials, void *stream);

/* gbins.summary (gbins_summary, src/GenomeTrackSummary.cpp:433-506): IntervalSummary of d_vals' rows grouped by the N-D  */
/* bin (BinsManager::vals2idx, src/BinsManager.h:53-72) of their ndim bin columns; breaks as in mg_hist, first dimension */
/* fastest. h_partials[6 * b ..] = bin b's six doubles. Rows outside every bin are skipped. Synchronises the stream.     */
int mg_bins_summary(mg_ctx *ctx, const double *d_vals, int ndim, const double *const *d_bin_cols, int64_t n, const double *h_breaks,
					const int *nbreaks, int include_lowest, double *h_partials, void *stream);
/* gbins.quantiles (gbins_quantiles, src/GenomeTrackQuantiles.cpp:1199-1330): a StreamPercentiler<double> per bin, exact */
/* regime. The non-NaN values are scattered next to their bin's others on the device and the bins go through             */
/* mg_quantiles_segmented. h_out[k * total_bins + b], NaN for an empty bin; h_nkept[b] may be NULL. Synchronises.        */
int mg_bins_quantiles(mg_ctx *ctx, const double *d_vals, int ndim, const double *const *d_bin_cols, int64_t n, const double *h_breaks,
					  const int *nbreaks, int include_lowest, int nq, const double *h_percentiles, double *h_out, int64_t *h_nkept, void *stream);

/* global.percentile / global.percentile.min / global.percentile.max (src/TrackVarProcessor.cpp:98-110): the column of the   */
/* vtrack's avg / min / max values (mg_*_window) is mapped IN PLACE through the track's precomputed percentile table --     */
/* h_breaks[nbreaks] and h_bins[nbreaks], the `breaks` attribute and the values of <track>/vars/pv.percentiles, which the    */
/* host keeps reading as before (src/TrackExpressionVars.cpp:1788-1823). NaN stays NaN; a value at or below the first break */
/* gets bins[0], one above the last break 1. Enqueues on the stream.                                                        */
int mg_percentile_lookup(mg_ctx *ctx, double *d_col, int64_t n, const double *h_breaks, const double *h_bins, int nbreaks, void *stream);

/* gscreen: merge consecutive TRUE rows (d_flag[i] == 1) of rows [first, first+count) into intervals. Outputs are   */
/* device arrays with room for `count` entries; *h_nout receives the number produced (the call synchronises).     */
int mg_screen_merge(mg_ctx *ctx, const mg_rows *rows, int64_t first, int64_t count, const int8_t *d_flag, int32_t *d_chrom,
					int64_t *d_start, int64_t *d_end, int64_t *h_nout, void *stream);
/* The same with the intervals in HOST arrays of `capacity` entries (pinned memory makes the copies run at the link's speed): the */
/* TRUE runs are counted first, the device side of the result is allocated at its size. *h_nout receives the number of intervals;  */
/* when it exceeds `capacity` nothing is copied and the caller calls again with larger arrays (what C_gscreen's growing vector of   */
/* intervals does, src/GenomeTrackScreener.cpp:195-220). Synchronises the stream.                                                  */
int mg_h_screen_merge(mg_ctx *ctx, const mg_rows *rows, int64_t first, int64_t count, const int8_t *d_flag, int32_t *h_chrom,
					  int64_t *h_start, int64_t *h_end, int64_t capacity, int64_t *h_nout, void *stream);

/* gscreen predicates that are comparisons of value columns against constants, all joined by `&` or all by `|` (e.g.   */
/* "near > 0.5 & cov > 0.25"): d_flag[i] = 1 where R's result is TRUE, else 0. A comparison with NaN is NA in R, and  */
/* NA & x / NA | FALSE are never TRUE, so a NaN operand counts as FALSE; the TRUE set is exactly R's                 */
/* (src/GenomeTrackScreener.cpp:203: only LOGICAL == 1 opens or extends an interval). Saves the device-to-host copy  */
/* of the columns (2.5 GB for two columns at a 20 bp iterator over hg38). The host keeps evaluating every other       */
/* expression on the returned vectors as before.                                                                     */
enum mg_cmp_op { MG_CMP_GT = 0, MG_CMP_GE = 1, MG_CMP_LT = 2, MG_CMP_LE = 3, MG_CMP_EQ = 4, MG_CMP_NE = 5 };
#define MG_PRED_MAX_TERMS 8
int mg_screen_compare(mg_ctx *ctx, int nterms, const double *const *d_cols, const int *ops, const double *thresholds, int combine_or,
					  int64_t n, int8_t *d_flag, void *stream);

/* ---- host-buffer convenience: what a .Call entry point does per vtrack group --------------------------------- */
/* Host arrays in, host columns out; uploads, kernels and downloads are chunked and overlapped on internal        */
/* streams. This is the end-to-end path bench.py times as `e2e`.                                                  */
int mg_h_dense_window(mg_ctx *ctx, const mg_dense *t, int64_t n, const int32_t *h_chrom, const int64_t *h_start, const int64_t *h_end,
					  int64_t sshift, int64_t eshift, int nfuncs, const int *funcs, const double *params, double *const *h_out);
/* The same pipeline for the other sources: arguments as mg_sparse_window / mg_coverage / mg_pwm with the rows and the result  */
/* column(s) in host memory.                                                                                                 */
int mg_h_sparse_window(mg_ctx *ctx, const mg_sparse *t, int64_t n, const int32_t *h_chrom, const int64_t *h_start, const int64_t *h_end,
					   int64_t sshift, int64_t eshift, int nfuncs, const int *funcs, const double *params, double *const *h_out);
int mg_h_coverage(mg_ctx *ctx, const mg_ivset *s, int64_t n, const int32_t *h_chrom, const int64_t *h_start, const int64_t *h_end,
				  int64_t sshift, int64_t eshift, double *h_out);
int mg_h_pwm(mg_ctx *ctx, const mg_seq *s, const float *h_logp, int L, int bidirect, int extend, int mode, int strand, float score_thresh,
			 int64_t n, const int32_t *h_chrom, const int64_t *h_start, const int64_t *h_end, int64_t sshift, int64_t eshift, double *h_out);

/* ---- several GPUs behind one host thread ----------------------------------------------------------------------------------- */
/* Replaces, for a single-threaded R process, what the reference does by forking: prepare4multitasking's split of the scope over     */
/* child processes (src/rdbinterval.cpp:446-564), the parent's concatenation of the children's rows                                  */
/* (src/GenomeTrackExtract.cpp:1461-1509) and its merges of their partial results (src/GenomeTrackSummary.cpp:205-216,              */
/* src/GenomeTrackDistribution.cpp:131-137, src/GenomeTrackQuantiles.cpp:278-400). A group owns one mg_ctx per device; every        */
/* chromosome is staged on ONE member, its owner (whole chromosomes, longest first, to the member holding the fewest bins: work per  */
/* member is proportional to the bins it owns); rows are routed to the owner of their chromosome; every member has its work enqueued */
/* before any is waited for. gsummary / gdist partials meet on member 0 (cudaMemcpyPeerAsync, NVLink where the pair has peer access)  */
/* and are merged there in member order; gquantiles runs ONE exact selection over the members' columns (the per-pass digit         */
/* histograms meet on the host). With one device a group call returns what the mg_ctx call returns, bit for bit.                   */
typedef struct mg_group mg_group;
typedef struct mg_gdense mg_gdense;
int mg_group_init(int ndev, const int *dev_ids, int nchroms, const int64_t *chrom_sizes, mg_group **out);
void mg_group_shutdown(mg_group *g);
const char *mg_group_last_error(const mg_group *g); /* g may be NULL: error of the last failed mg_group_init in this thread */
int mg_group_size(const mg_group *g);
mg_ctx *mg_group_ctx(mg_group *g, int member);
int mg_group_owner(const mg_group *g, int chromid); /* member that stages / serves this chromosome */
int mg_group_set_option(mg_group *g, const char *name, int64_t value); /* mg_set_option on every member */
int mg_group_sync(mg_group *g);
int mg_group_dense_create(mg_group *g, uint32_t bin_size, mg_gdense **out);
int mg_group_dense_stage(mg_group *g, mg_gdense *t, int chromid, const float *h_bins, int64_t nbins); /* on the owner */
void mg_group_dense_free(mg_group *g, mg_gdense *t);
int64_t mg_group_dense_bytes(const mg_gdense *t, int member);
/* gextract (mg_h_dense_window over the group): host rows in any order, host columns out in the rows' order */
int mg_group_h_dense_window(mg_group *g, const mg_gdense *t, int64_t n, const int32_t *h_chrom, const int64_t *h_start, const int64_t *h_end,
							int64_t sshift, int64_t eshift, int nfuncs, const int *funcs, const double *params, double *const *h_out);
/* gsummary / gdist / gquantiles of one dense-track function over the fixed-bin iterator `binsize` of a scope (host intervals):   */
/* h_partial[6] as mg_dense_summary accumulates it (finalise with mg_summary_finalize), h_counts[nbreaks - 1], h_out[nq].          */
int mg_group_dense_summary(mg_group *g, const mg_gdense *t, int64_t binsize, int64_t nscope, const int32_t *h_chrom, const int64_t *h_start,
						   const int64_t *h_end, int64_t sshift, int64_t eshift, int func, double param, double *h_partial);
int mg_group_dense_hist(mg_group *g, const mg_gdense *t, int64_t binsize, int64_t nscope, const int32_t *h_chrom, const int64_t *h_start,
						const int64_t *h_end, int64_t sshift, int64_t eshift, int func, double param, const double *h_breaks, int nbreaks,
						int include_lowest, uint64_t *h_counts);
int mg_group_dense_quantiles(mg_group *g, const mg_gdense *t, int64_t binsize, int64_t nscope, const int32_t *h_chrom, const int64_t *h_start,
							 const int64_t *h_end, int64_t sshift, int64_t eshift, int func, double param, int nq, const double *h_percentiles,
							 double *h_out, int64_t *h_nkept);

#ifdef __cplusplus
}
#endif
#endif /* MISHAGPU_H_ */
