// GpuSession.cpp -- the R-side half of the drop-in: the file a maintainer adds to misha's src/ (INTEGRATION.md sections 2-4).
//
// It is compiled HERE only to prove that it builds: tests/test_abi_and_host.py runs `g++ -fsyntax-only` on it against the
// reference's own headers (/root/reference/src, when present) and the declaration-only R shim of oracle/shim -- no R in this image, so
// nothing below can run here. Entry points keep the reference's .Call signatures (src/misha-init.cpp:83-120) and FALL THROUGH to the
// reference's own *_multitask functions for everything they do not take (2-D scopes, expressions that are not a bare dense track
// under a fixed-bin iterator, bands): the R API stays a drop-in whatever the expression.
//
//   gtracksummary_gpu   <- gtracksummary_multitask   src/GenomeTrackSummary.cpp:174-226
//   gtrackdist_gpu      <- gtrackdist_multitask      src/GenomeTrackDistribution.cpp:82-150
//   gquantiles_gpu      <- gquantiles_multitask      src/GenomeTrackQuantiles.cpp:278-400
//   gextract_gpu        <- gextract_multitask        src/GenomeTrackExtract.cpp:1128-1560 (scope = iterator, bare dense track columns)
//
// One mg_group per R process holds every device of `options(gcuda.devices = ...)` (default: device 0); the R process stays
// single-threaded and never forks on this path -- a CUDA context does not survive fork() (src/rdbutils.cpp:304-377).
#include <cstdint>
#include <limits>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "rdbinterval.h"
#include "rdbutils.h"
#include "BinsManager.h"
#include "GenomeTrack.h"
#include "GenomeTrackFixedBin.h"
#include "GIntervalsFetcher1D.h"
#include "GIntervalsFetcher2D.h"

#include "mishagpu.h"

using namespace std;
using namespace rdb;

// the reference's own entry points: what every call falls back to
extern "C" {
SEXP gtracksummary_multitask(SEXP _expr, SEXP _intervals, SEXP _iterator_policy, SEXP _band, SEXP _envir);
SEXP gtrackdist_multitask(SEXP _intervals, SEXP _track_exprs, SEXP _breaks, SEXP _include_lowest, SEXP _iterator_policy, SEXP _band, SEXP _envir);
SEXP gquantiles_multitask(SEXP _intervals, SEXP _expr, SEXP _percentiles, SEXP _iterator_policy, SEXP _band, SEXP _envir);
SEXP gextract_multitask(SEXP _intervals, SEXP _exprs, SEXP _colnames, SEXP _iterator_policy, SEXP _band, SEXP _file, SEXP _intervals_set_out,
						SEXP _intervals_join, SEXP _envir);
}

namespace {

struct GpuSession { // one per R process, created by the first GPU-routed .Call
	mg_group *group = nullptr;
	unordered_map<string, mg_gdense *> dense; // track name -> staged handle ("stages chromosome track blocks into HBM once")
	~GpuSession()
	{
		for (auto &kv : dense)
			mg_group_dense_free(group, kv.second);
		mg_group_shutdown(group);
	}
};

GpuSession &session()
{
	static GpuSession s;
	return s;
}

void gpu_check(int rc)
{
	if (rc)
		verror("%s", mg_group_last_error(session().group)); // an R error through TGLException, never a CPU answer in its place
}

mg_group *gpu_group(const GenomeChromKey &ck)
{
	GpuSession &s = session();
	if (s.group)
		return s.group;
	vector<int64_t> sizes(ck.get_num_chroms());
	for (size_t i = 0; i < sizes.size(); ++i)
		sizes[i] = (int64_t)ck.get_chrom_size((int)i);
	vector<int> devices;
	SEXP opt = Rf_GetOption1(Rf_install("gcuda.devices"));
	if (Rf_isInteger(opt) || Rf_isReal(opt))
		for (int i = 0; i < Rf_length(opt); ++i)
			devices.push_back(Rf_isInteger(opt) ? INTEGER(opt)[i] : (int)REAL(opt)[i]);
	if (devices.empty())
		devices.push_back(0);
	if (mg_group_init((int)devices.size(), devices.data(), (int)sizes.size(), sizes.data(), &s.group))
		verror("%s", mg_group_last_error(nullptr));
	return s.group;
}

// a dense track of the database, staged chromosome by chromosome on its owners (replaces the per-chromosome open of
// TrackExpressionVars::start_chrom, src/TrackExpressionVars.cpp:1885-1957). nullptr: not a dense track.
mg_gdense *gpu_stage_dense(IntervUtils &iu, const string &track)
{
	GpuSession &s = session();
	auto it = s.dense.find(track);
	if (it != s.dense.end())
		return it->second;
	const GenomeChromKey &ck = iu.get_chromkey();
	const string dir = track2path(iu.get_env(), track);
	if (GenomeTrack::get_type(dir.c_str(), ck) != GenomeTrack::FIXED_BIN)
		return nullptr;
	mg_group *g = gpu_group(ck);
	mg_gdense *h = nullptr;
	vector<float> bins;
	for (int chromid = 0; chromid < (int)ck.get_num_chroms(); ++chromid) {
		GenomeTrackFixedBin t;
		const string file = dir + "/" + GenomeTrack::get_1d_filename(ck, chromid);
		try {
			t.init_read(file.c_str(), chromid); // src/GenomeTrackFixedBin.cpp:1129
		} catch (TGLException &) {
			continue; // a chromosome without data: its rows are NaN on either path
		}
		if (!h)
			gpu_check(mg_group_dense_create(g, t.get_bin_size(), &h));
		t.read_bins_bulk(0, t.get_num_samples(), bins); // :1028; +-inf become NaN on the device as read_bins_bulk's callers expect
		gpu_check(mg_group_dense_stage(g, h, chromid, bins.data(), (int64_t)bins.size()));
	}
	if (h)
		s.dense[track] = h;
	return h;
}

// The shape this file takes over: ONE expression that is a bare dense track name, a numeric (fixed-bin) iterator, no band, a 1-D scope.
// Everything else is the reference's.
struct DenseCall {
	mg_gdense *track = nullptr;
	int64_t binsize = 0;
	vector<int32_t> chrom;
	vector<int64_t> start, end;
};

bool dense_call(SEXP _expr, SEXP _intervals, SEXP _iterator_policy, SEXP _band, IntervUtils &iu, DenseCall &c, bool unify)
{
	if (!Rf_isString(_expr) || Rf_length(_expr) != 1 || !Rf_isNull(_band))
		return false;
	if (!(Rf_isReal(_iterator_policy) || Rf_isInteger(_iterator_policy)) || Rf_length(_iterator_policy) != 1)
		return false;
	c.binsize = Rf_isReal(_iterator_policy) ? (int64_t)REAL(_iterator_policy)[0] : (int64_t)INTEGER(_iterator_policy)[0];
	if (c.binsize <= 0)
		return false;
	GIntervalsFetcher1D *intervals1d = NULL;
	GIntervalsFetcher2D *intervals2d = NULL;
	iu.convert_rintervs(_intervals, &intervals1d, &intervals2d);
	unique_ptr<GIntervalsFetcher1D> g1(intervals1d);
	unique_ptr<GIntervalsFetcher2D> g2(intervals2d);
	if (intervals2d->size())
		return false;
	c.track = gpu_stage_dense(iu, CHAR(STRING_ELT(_expr, 0)));
	if (!c.track)
		return false;
	intervals1d->sort();
	if (unify)
		intervals1d->unify_overlaps();
	for (intervals1d->begin_iter(); !intervals1d->isend(); intervals1d->next()) {
		const GInterval &iv = intervals1d->cur_interval();
		c.chrom.push_back(iv.chromid);
		c.start.push_back(iv.start);
		c.end.push_back(iv.end);
	}
	return true;
}

} // namespace

extern "C" {

SEXP gtracksummary_gpu(SEXP _expr, SEXP _intervals, SEXP _iterator_policy, SEXP _band, SEXP _envir)
{
	try {
		RdbInitializer rdb_init;
		IntervUtils iu(_envir);
		DenseCall c;
		if (!dense_call(_expr, _intervals, _iterator_policy, _band, iu, c, true))
			return gtracksummary_multitask(_expr, _intervals, _iterator_policy, _band, _envir);
		double part[6], out[7];
		gpu_check(mg_group_dense_summary(session().group, c.track, c.binsize, (int64_t)c.chrom.size(), c.chrom.data(), c.start.data(), c.end.data(), 0, 0,
										 MG_AVG, 0., part));
		mg_summary_finalize(part, out); // the arithmetic of src/GenomeTrackSummary.cpp:148-155
		static const char *names[7] = {"Total intervals", "NaN intervals", "Min", "Max", "Sum", "Mean", "Std dev"};
		SEXP answer = rprotect_ptr(RSaneAllocVector(REALSXP, 7));
		SEXP colnames = rprotect_ptr(RSaneAllocVector(STRSXP, 7));
		for (int i = 0; i < 7; ++i) {
			REAL(answer)[i] = out[i];
			SET_STRING_ELT(colnames, i, Rf_mkChar(names[i]));
		}
		Rf_setAttrib(answer, R_NamesSymbol, colnames);
		runprotect(2);
		return answer;
	} catch (TGLException &e) {
		rerror("%s", e.msg());
	} catch (const bad_alloc &e) {
		rerror("Out of memory");
	}
	return R_NilValue;
}

SEXP gtrackdist_gpu(SEXP _intervals, SEXP _track_exprs, SEXP _breaks, SEXP _include_lowest, SEXP _iterator_policy, SEXP _band, SEXP _envir)
{
	try {
		RdbInitializer rdb_init;
		IntervUtils iu(_envir);
		DenseCall c;
		if (!Rf_isString(_track_exprs) || Rf_length(_track_exprs) != 1 || !Rf_isNewList(_breaks) || Rf_length(_breaks) != 1 ||
			!dense_call(_track_exprs, _intervals, _iterator_policy, _band, iu, c, true))
			return gtrackdist_multitask(_intervals, _track_exprs, _breaks, _include_lowest, _iterator_policy, _band, _envir);
		SEXP br = VECTOR_ELT(_breaks, 0);
		if (!Rf_isReal(br) || Rf_length(br) < 2)
			verror("breaks[1] must be a numeric vector of at least two elements");
		const int nbreaks = Rf_length(br);
		const bool include_lowest = Rf_isLogical(_include_lowest) && LOGICAL(_include_lowest)[0] == 1;
		vector<uint64_t> counts((size_t)nbreaks - 1);
		gpu_check(mg_group_dense_hist(session().group, c.track, c.binsize, (int64_t)c.chrom.size(), c.chrom.data(), c.start.data(), c.end.data(), 0, 0, MG_AVG,
									  0., REAL(br), nbreaks, include_lowest ? 1 : 0, counts.data()));
		// distribution[] of src/GenomeTrackDistribution.cpp:121-141 as an R numeric vector; dimnames as BinsManager::set_dims writes them
		BinsManager bins_manager(_breaks, _include_lowest);
		SEXP answer = rprotect_ptr(RSaneAllocVector(REALSXP, (R_xlen_t)counts.size()));
		for (size_t i = 0; i < counts.size(); ++i)
			REAL(answer)[i] = (double)counts[i];
		SEXP dim = rprotect_ptr(RSaneAllocVector(INTSXP, 1));
		SEXP dimnames = rprotect_ptr(RSaneAllocVector(VECSXP, 1));
		bins_manager.set_dims(dim, dimnames);
		Rf_setAttrib(answer, R_DimSymbol, dim);
		Rf_setAttrib(answer, R_DimNamesSymbol, dimnames);
		runprotect(3);
		return answer;
	} catch (TGLException &e) {
		rerror("%s", e.msg());
	} catch (const bad_alloc &e) {
		rerror("Out of memory");
	}
	return R_NilValue;
}

SEXP gquantiles_gpu(SEXP _intervals, SEXP _expr, SEXP _percentiles, SEXP _iterator_policy, SEXP _band, SEXP _envir)
{
	try {
		RdbInitializer rdb_init;
		IntervUtils iu(_envir);
		DenseCall c;
		if (!Rf_isReal(_percentiles) || Rf_length(_percentiles) < 1 || !dense_call(_expr, _intervals, _iterator_policy, _band, iu, c, true))
			return gquantiles_multitask(_intervals, _expr, _percentiles, _iterator_policy, _band, _envir);
		const int nq = Rf_length(_percentiles);
		SEXP answer = rprotect_ptr(RSaneAllocVector(REALSXP, nq));
		int64_t nkept = 0;
		// one exact selection over every member's share of the column: the sampling branch of src/GenomeTrackQuantiles.cpp:233-247 has no
		// counterpart (selection stays exact beyond gmax.data.size values)
		gpu_check(mg_group_dense_quantiles(session().group, c.track, c.binsize, (int64_t)c.chrom.size(), c.chrom.data(), c.start.data(), c.end.data(), 0, 0,
										   MG_AVG, 0., nq, REAL(_percentiles), REAL(answer), &nkept));
		runprotect(1);
		return answer;
	} catch (TGLException &e) {
		rerror("%s", e.msg());
	} catch (const bad_alloc &e) {
		rerror("Out of memory");
	}
	return R_NilValue;
}

// gextract(track, intervals = scope, iterator = scope): the scope's own intervals are the rows (config C3's shape with a bare track;
// virtual tracks add their function / shifts from .misha$GVTRACKS the same way).
SEXP gextract_gpu(SEXP _intervals, SEXP _exprs, SEXP _colnames, SEXP _iterator_policy, SEXP _band, SEXP _file, SEXP _intervals_set_out,
				  SEXP _intervals_join, SEXP _envir)
{
	try {
		RdbInitializer rdb_init;
		IntervUtils iu(_envir);
		// the scope as its own iterator is what a NULL / intervals-valued policy over a 1-D scope means for a bare dense track only when the
		// caller says so; this sketch takes `iterator = intervals` (policy == the scope object) and leaves the rest to the reference
		if (!Rf_isString(_exprs) || Rf_length(_exprs) != 1 || !Rf_isNull(_band) || !Rf_isNull(_file) || !Rf_isNull(_intervals_set_out) ||
			_iterator_policy != _intervals)
			return gextract_multitask(_intervals, _exprs, _colnames, _iterator_policy, _band, _file, _intervals_set_out, _intervals_join, _envir);
		GIntervalsFetcher1D *intervals1d = NULL;
		GIntervalsFetcher2D *intervals2d = NULL;
		iu.convert_rintervs(_intervals, &intervals1d, &intervals2d);
		unique_ptr<GIntervalsFetcher1D> g1(intervals1d);
		unique_ptr<GIntervalsFetcher2D> g2(intervals2d);
		mg_gdense *track = intervals2d->size() ? nullptr : gpu_stage_dense(iu, CHAR(STRING_ELT(_exprs, 0)));
		if (!track)
			return gextract_multitask(_intervals, _exprs, _colnames, _iterator_policy, _band, _file, _intervals_set_out, _intervals_join, _envir);
		intervals1d->sort(); // src/GenomeTrackExtract.cpp:1172-1177
		vector<int32_t> chrom;
		vector<int64_t> start, end;
		vector<int> ids; // intervalID = 1-based row of the caller's scope (udata, src/IntervalConverter.cpp:357)
		for (intervals1d->begin_iter(); !intervals1d->isend(); intervals1d->next()) {
			const GInterval &iv = intervals1d->cur_interval();
			chrom.push_back(iv.chromid);
			start.push_back(iv.start);
			end.push_back(iv.end);
			ids.push_back((int)(intptr_t)iv.udata + 1);
		}
		const R_xlen_t n = (R_xlen_t)chrom.size();
		if (!n)
			return R_NilValue;
		SEXP col = rprotect_ptr(RSaneAllocVector(REALSXP, n));
		const int func = MG_AVG;
		const double param = 0;
		double *outp = REAL(col);
		gpu_check(mg_group_h_dense_window(session().group, track, (int64_t)n, chrom.data(), start.data(), end.data(), 0, 0, 1, &func, &param, &outp));
		// the data frame: chrom / start / end from the sorted scope, the value column, intervalID
		SEXP answer = rprotect_ptr(iu.convert_intervs(intervals1d, GInterval::NUM_COLS + 2, false));
		SEXP rids = rprotect_ptr(RSaneAllocVector(INTSXP, n));
		for (R_xlen_t i = 0; i < n; ++i)
			INTEGER(rids)[i] = ids[(size_t)i];
		SET_VECTOR_ELT(answer, GInterval::NUM_COLS, col);
		SET_VECTOR_ELT(answer, GInterval::NUM_COLS + 1, rids);
		SEXP names = Rf_getAttrib(answer, R_NamesSymbol);
		SET_STRING_ELT(names, GInterval::NUM_COLS, Rf_isNull(_colnames) ? STRING_ELT(_exprs, 0) : STRING_ELT(_colnames, 0));
		SET_STRING_ELT(names, GInterval::NUM_COLS + 1, Rf_mkChar("intervalID"));
		runprotect(3);
		return answer;
	} catch (TGLException &e) {
		rerror("%s", e.msg());
	} catch (const bad_alloc &e) {
		rerror("Out of memory");
	}
	return R_NilValue;
}

} // extern "C"
