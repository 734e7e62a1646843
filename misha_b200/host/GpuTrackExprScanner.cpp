// GpuTrackExprScanner.cpp -- see the header. Host logic only: sorting / unifying the scope, grouping the variables,
// buffer management; every value comes from libmishagpu.so (include/mishagpu.h).
#include "GpuTrackExprScanner.h"

#include <algorithm>
#include <cstring>
#include <numeric>
#include <tuple>

namespace mishagpu {

GpuTrackExprScanner::GpuTrackExprScanner(mg_ctx *ctx, int64_t buf_rows) : m_ctx(ctx), m_buf_rows(buf_rows < 1024 ? 1024 : buf_rows)
{
	if (!ctx)
		throw MishaGpuException("GpuTrackExprScanner: no context (mg_init failed? this library has no CPU fallback)");
}

GpuTrackExprScanner::~GpuTrackExprScanner() { release(); }

void GpuTrackExprScanner::check(int rc) const
{
	if (rc)
		throw MishaGpuException(mg_last_error(m_ctx));
}

void GpuTrackExprScanner::vtrack_dense(const std::string &name, const mg_dense *track, int func, double param, int64_t sshift, int64_t eshift)
{
	VTrack v;
	v.source = VTrack::DENSE;
	v.dense = track;
	v.func = func;
	v.param = param;
	v.sshift = sshift;
	v.eshift = eshift;
	m_vtracks[name] = v;
}

void GpuTrackExprScanner::vtrack_sparse(const std::string &name, const mg_sparse *track, int func, double param, int64_t sshift, int64_t eshift)
{
	VTrack v;
	v.source = VTrack::SPARSE;
	v.sparse = track;
	v.func = func;
	v.param = param;
	v.sshift = sshift;
	v.eshift = eshift;
	m_vtracks[name] = v;
}

void GpuTrackExprScanner::vtrack_coverage(const std::string &name, const mg_ivset *src, int64_t sshift, int64_t eshift)
{
	VTrack v;
	v.source = VTrack::COVERAGE;
	v.ivset = src;
	v.sshift = sshift;
	v.eshift = eshift;
	m_vtracks[name] = v;
}

void GpuTrackExprScanner::vtrack_kmer(const std::string &name, const mg_seq *seq, int mode, const std::string &kmer, bool extend, int strand,
									  int64_t sshift, int64_t eshift)
{
	if (kmer.empty())
		throw MishaGpuException("kmer sequence cannot be empty");
	VTrack v;
	v.source = VTrack::KMER;
	v.seq = seq;
	v.kmer_mode = mode;
	v.kmer = kmer;
	v.extend = extend;
	v.strand = strand;
	v.sshift = sshift;
	v.eshift = eshift;
	m_vtracks[name] = v;
}

void GpuTrackExprScanner::vtrack_pwm(const std::string &name, const mg_seq *seq, int mode, const std::vector<double> &pssm, double prior, bool bidirect,
									 bool extend, int strand, float score_thresh, int64_t sshift, int64_t eshift)
{
	if (pssm.empty() || pssm.size() % 4)
		throw MishaGpuException("Virtual track " + name + ": PWM functions require a matrix with columns A, C, G, T");
	VTrack v;
	v.source = VTrack::PWM;
	v.seq = seq;
	v.pwm_mode = mode;
	v.logp.resize(pssm.size());
	if (mg_pssm_logp(pssm.data(), (int)(pssm.size() / 4), prior, v.logp.data()))
		throw MishaGpuException("Virtual track " + name + ": invalid PSSM");
	v.bidirect = bidirect;
	v.extend = extend;
	v.strand = strand;
	v.score_thresh = score_thresh;
	v.sshift = sshift;
	v.eshift = eshift;
	m_vtracks[name] = v;
}

void GpuTrackExprScanner::vtrack_rm(const std::string &name) { m_vtracks.erase(name); }

void GpuTrackExprScanner::release()
{
	for (double *p : m_dcols)
		mg_dev_free(m_ctx, p);
	for (double *p : m_cols)
		mg_host_free(m_ctx, p);
	m_dcols.clear();
	m_cols.clear();
	void *dev[] = {m_dchrom, m_dstart, m_dend, m_dscope};
	void *host[] = {m_chrom, m_start, m_end, m_scope_idx};
	for (void *p : dev)
		if (p)
			mg_dev_free(m_ctx, p);
	for (void *p : host)
		if (p)
			mg_host_free(m_ctx, p);
	m_dchrom = m_chrom = nullptr;
	m_dstart = m_dend = m_dscope = m_start = m_end = m_scope_idx = nullptr;
	if (m_rows)
		mg_rows_free(m_ctx, m_rows);
	m_rows = nullptr;
	m_isend = true;
}

// sort by (chromid, start), stable; optionally merge overlapping or touching intervals (GIntervals::unify_overlaps(true))
void GpuTrackExprScanner::prepare_scope(const std::vector<Interval> &scope, bool unify_scope)
{
	std::vector<int64_t> perm(scope.size());
	std::iota(perm.begin(), perm.end(), int64_t(0));
	std::stable_sort(perm.begin(), perm.end(), [&](int64_t a, int64_t b) {
		return std::tie(scope[a].chromid, scope[a].start) < std::tie(scope[b].chromid, scope[b].start);
	});
	m_scope.clear();
	m_scope_orig.clear();
	for (int64_t k : perm) {
		const Interval &iv = scope[k];
		if (unify_scope && !m_scope.empty() && m_scope.back().chromid == iv.chromid && m_scope.back().end >= iv.start) {
			if (m_scope.back().end < iv.end)
				m_scope.back().end = iv.end;
			continue;
		}
		m_scope.push_back(iv);
		m_scope_orig.push_back(unify_scope ? (int64_t)m_scope.size() - 1 : k);
	}
}

bool GpuTrackExprScanner::start(const std::vector<std::string> &track_exprs)
{
	m_track_exprs = track_exprs;
	m_vars.clear();
	for (const std::string &e : track_exprs) {
		auto it = m_vtracks.find(e);
		if (it == m_vtracks.end())
			throw MishaGpuException("Track expression \"" + e + "\" is not a virtual track known to this scanner "
									"(expressions other than a bare variable are evaluated by R on the returned columns)");
		m_vars.push_back(&it->second);
	}
	m_nrows = mg_rows_count(m_rows);
	const int64_t cap = std::max<int64_t>(1, std::min(m_buf_rows, m_nrows));
	for (size_t i = 0; i < m_vars.size(); ++i) {
		void *d = nullptr, *h = nullptr;
		check(mg_dev_alloc(m_ctx, cap * 8, &d));
		m_dcols.push_back((double *)d);
		check(mg_host_alloc(m_ctx, cap * 8, &h));
		m_cols.push_back((double *)h);
	}
	void *p = nullptr;
	check(mg_dev_alloc(m_ctx, cap * 4, &p)); m_dchrom = (int32_t *)p;
	check(mg_dev_alloc(m_ctx, cap * 8, &p)); m_dstart = (int64_t *)p;
	check(mg_dev_alloc(m_ctx, cap * 8, &p)); m_dend = (int64_t *)p;
	check(mg_dev_alloc(m_ctx, cap * 8, &p)); m_dscope = (int64_t *)p;
	check(mg_host_alloc(m_ctx, cap * 4, &p)); m_chrom = (int32_t *)p;
	check(mg_host_alloc(m_ctx, cap * 8, &p)); m_start = (int64_t *)p;
	check(mg_host_alloc(m_ctx, cap * 8, &p)); m_end = (int64_t *)p;
	check(mg_host_alloc(m_ctx, cap * 8, &p)); m_scope_idx = (int64_t *)p;
	m_buf_rows = cap;
	m_isend = m_nrows == 0;
	m_buf_first = m_buf_count = m_buf_idx = 0;
	if (!m_isend)
		fill(0);
	return !m_isend;
}

bool GpuTrackExprScanner::begin(const std::vector<std::string> &track_exprs, const std::vector<Interval> &scope, int64_t iterator_binsize,
								bool unify_scope)
{
	release();
	prepare_scope(scope, unify_scope);
	std::vector<int32_t> c(m_scope.size());
	std::vector<int64_t> s(m_scope.size()), e(m_scope.size());
	for (size_t i = 0; i < m_scope.size(); ++i) {
		c[i] = m_scope[i].chromid;
		s[i] = m_scope[i].start;
		e[i] = m_scope[i].end;
	}
	check(mg_rows_fixedbin(m_ctx, iterator_binsize, (int64_t)c.size(), c.data(), s.data(), e.data(), &m_rows));
	return start(track_exprs);
}

bool GpuTrackExprScanner::begin(const std::vector<std::string> &track_exprs, const std::vector<Interval> &scope,
								const std::vector<Interval> &iterator_intervals, bool unify_scope)
{
	release();
	prepare_scope(scope, unify_scope);
	// iterator intervals: sort + unify_overlaps(false): touching intervals stay apart (src/TrackExpressionScanner.cpp:934-954)
	std::vector<Interval> it(iterator_intervals);
	std::stable_sort(it.begin(), it.end(), [](const Interval &a, const Interval &b) { return std::tie(a.chromid, a.start) < std::tie(b.chromid, b.start); });
	std::vector<int32_t> ic, sc(m_scope.size());
	std::vector<int64_t> is, ie, ss(m_scope.size()), se(m_scope.size());
	for (const Interval &iv : it) {
		if (!ic.empty() && ic.back() == iv.chromid && ie.back() > iv.start) {
			if (ie.back() < iv.end)
				ie.back() = iv.end;
			continue;
		}
		ic.push_back(iv.chromid);
		is.push_back(iv.start);
		ie.push_back(iv.end);
	}
	for (size_t i = 0; i < m_scope.size(); ++i) {
		sc[i] = m_scope[i].chromid;
		ss[i] = m_scope[i].start;
		se[i] = m_scope[i].end;
	}
	check(mg_rows_intervals(m_ctx, (int64_t)ic.size(), ic.data(), is.data(), ie.data(), (int64_t)sc.size(), sc.data(), ss.data(), se.data(), nullptr,
							&m_rows));
	return start(track_exprs);
}

// one libmishagpu call per (source, sshift, eshift) group for rows [first, first + count), then one copy per column
bool GpuTrackExprScanner::fill(int64_t first)
{
	const int64_t count = std::min(m_buf_rows, m_nrows - first);
	if (count <= 0) {
		m_isend = true;
		return false;
	}
	std::vector<bool> done(m_vars.size(), false);
	for (size_t i = 0; i < m_vars.size(); ++i) {
		if (done[i])
			continue;
		const VTrack &v = *m_vars[i];
		if (v.source == VTrack::COVERAGE) {
			check(mg_coverage(m_ctx, v.ivset, m_rows, first, count, v.sshift, v.eshift, m_dcols[i], nullptr));
			done[i] = true;
			continue;
		}
		if (v.source == VTrack::PWM) { // PWMScorer::score_interval per row (src/SequenceVarProcessor.cpp:471) -> one mg_pwm per vtrack
			check(mg_pwm(m_ctx, v.seq, v.logp.data(), (int)(v.logp.size() / 4), v.bidirect, v.extend, v.pwm_mode, v.strand, v.score_thresh, m_rows,
						 first, count, v.sshift, v.eshift, m_dcols[i], nullptr));
			done[i] = true;
			continue;
		}
		if (v.source == VTrack::KMER) { // KmerCounter::score_interval per row -> one mg_kmer per vtrack
			check(mg_kmer(m_ctx, v.seq, v.kmer.c_str(), v.kmer_mode, v.extend, v.strand, m_rows, first, count, v.sshift, v.eshift, m_dcols[i], nullptr));
			done[i] = true;
			continue;
		}
		std::vector<int> funcs;
		std::vector<double> params;
		std::vector<double *> cols;
		std::vector<size_t> members;
		for (size_t j = i; j < m_vars.size(); ++j) {
			const VTrack &u = *m_vars[j];
			if (done[j] || u.source != v.source || u.dense != v.dense || u.sparse != v.sparse || u.sshift != v.sshift || u.eshift != v.eshift)
				continue;
			// one column per function and call: the same function twice (or two expressions naming one vtrack) gets its own call
			bool dup = false;
			for (size_t m : members)
				dup = dup || (m_vars[m]->func == u.func && u.func != MG_QUANTILE);
			if (dup)
				continue;
			funcs.push_back(u.func);
			params.push_back(u.param);
			cols.push_back(m_dcols[j]);
			members.push_back(j);
			done[j] = true;
		}
		if (v.source == VTrack::DENSE)
			check(mg_dense_window(m_ctx, v.dense, m_rows, first, count, v.sshift, v.eshift, (int)funcs.size(), funcs.data(), params.data(), cols.data(),
								  nullptr));
		else
			check(mg_sparse_window(m_ctx, v.sparse, m_rows, first, count, v.sshift, v.eshift, (int)funcs.size(), funcs.data(), params.data(),
								   cols.data(), nullptr));
	}
	check(mg_rows_coords(m_ctx, m_rows, first, count, m_dchrom, m_dstart, m_dend, m_dscope, nullptr));
	for (size_t i = 0; i < m_vars.size(); ++i)
		check(mg_copy_d2h(m_ctx, m_cols[i], m_dcols[i], count * 8, nullptr));
	check(mg_copy_d2h(m_ctx, m_chrom, m_dchrom, count * 4, nullptr));
	check(mg_copy_d2h(m_ctx, m_start, m_dstart, count * 8, nullptr));
	check(mg_copy_d2h(m_ctx, m_end, m_dend, count * 8, nullptr));
	check(mg_copy_d2h(m_ctx, m_scope_idx, m_dscope, count * 8, nullptr));
	check(mg_sync(m_ctx, nullptr));
	for (int64_t r = 0; r < count; ++r) // sorted-scope index -> the caller's order
		m_scope_idx[r] = m_scope_idx[r] >= 0 ? m_scope_orig[(size_t)m_scope_idx[r]] : -1;
	m_buf_first = first;
	m_buf_count = count;
	m_buf_idx = 0;
	return true;
}

bool GpuTrackExprScanner::next()
{
	if (m_isend)
		return false;
	if (++m_buf_idx < m_buf_count)
		return true;
	if (!fill(m_buf_first + m_buf_count)) {
		m_isend = true;
		return false;
	}
	return true;
}

bool GpuTrackExprScanner::next_batch()
{
	if (m_isend)
		return false;
	m_buf_idx = m_buf_count - 1;
	return next();
}

ExtractResult gextract(GpuTrackExprScanner &scanner, const std::vector<std::string> &track_exprs, const std::vector<Interval> &scope,
					   int64_t iterator_binsize, const std::vector<Interval> *iterator_intervals, bool per_row)
{
	ExtractResult res;
	res.values.resize(track_exprs.size());
	if (iterator_intervals)
		scanner.begin(track_exprs, scope, *iterator_intervals, false);
	else
		scanner.begin(track_exprs, scope, iterator_binsize, false);
	const int64_t n = scanner.total_rows();
	res.chrom.reserve(n);
	res.start.reserve(n);
	res.end.reserve(n);
	res.interval_id.reserve(n);
	for (auto &c : res.values)
		c.reserve(n);
	if (per_row) { // the reference's loop, src/GenomeTrackExtract.cpp:1443 (kids) / :905-975 (single process)
		for (; !scanner.isend(); scanner.next()) {
			const Interval iv = scanner.last_interval1d();
			res.chrom.push_back(iv.chromid);
			res.start.push_back(iv.start);
			res.end.push_back(iv.end);
			res.interval_id.push_back((int32_t)scanner.last_scope_idx() + 1);
			for (size_t i = 0; i < track_exprs.size(); ++i)
				res.values[i].push_back(scanner.last_real((int)i));
		}
		return res;
	}
	while (!scanner.isend()) {
		const int64_t m = scanner.batch_size();
		res.chrom.insert(res.chrom.end(), scanner.batch_chroms(), scanner.batch_chroms() + m);
		res.start.insert(res.start.end(), scanner.batch_starts(), scanner.batch_starts() + m);
		res.end.insert(res.end.end(), scanner.batch_ends(), scanner.batch_ends() + m);
		for (int64_t r = 0; r < m; ++r)
			res.interval_id.push_back((int32_t)scanner.batch_scope_idx()[r] + 1);
		for (size_t i = 0; i < track_exprs.size(); ++i)
			res.values[i].insert(res.values[i].end(), scanner.batch_real((int)i), scanner.batch_real((int)i) + m);
		scanner.next_batch();
	}
	return res;
}

void gsummary_dense(mg_ctx *ctx, const mg_dense *track, const std::vector<Interval> &scope, int64_t iterator_binsize, int func, double param,
					int64_t sshift, int64_t eshift, double out[7])
{
	auto check = [&](int rc) {
		if (rc)
			throw MishaGpuException(mg_last_error(ctx));
	};
	// sort + unify the scope as gtracksummary does before scanning
	std::vector<Interval> sc(scope);
	std::stable_sort(sc.begin(), sc.end(), [](const Interval &a, const Interval &b) { return std::tie(a.chromid, a.start) < std::tie(b.chromid, b.start); });
	std::vector<int32_t> c;
	std::vector<int64_t> s, e;
	for (const Interval &iv : sc) {
		if (!c.empty() && c.back() == iv.chromid && e.back() >= iv.start) {
			if (e.back() < iv.end)
				e.back() = iv.end;
			continue;
		}
		c.push_back(iv.chromid);
		s.push_back(iv.start);
		e.push_back(iv.end);
	}
	mg_rows *rows = nullptr;
	check(mg_rows_fixedbin(ctx, iterator_binsize, (int64_t)c.size(), c.data(), s.data(), e.data(), &rows));
	void *d_part = nullptr;
	check(mg_dev_alloc(ctx, 6 * sizeof(double), &d_part));
	check(mg_summary_init(ctx, (double *)d_part, nullptr));
	check(mg_dense_summary(ctx, track, rows, 0, mg_rows_count(rows), sshift, eshift, func, param, (double *)d_part, nullptr));
	double part[6];
	check(mg_copy_d2h(ctx, part, d_part, sizeof(part), nullptr));
	check(mg_sync(ctx, nullptr));
	mg_summary_finalize(part, out);
	mg_dev_free(ctx, d_part);
	mg_rows_free(ctx, rows);
}

// ---- GpuGroupSession ---------------------------------------------------------------------------------------------------------
// sort + unify a scope as gtracksummary / gtrackdist / gquantiles do before scanning (unify_overlaps, src/TrackExpressionScanner.cpp:934-954)
static void unified_scope(const std::vector<Interval> &scope, std::vector<int32_t> &c, std::vector<int64_t> &s, std::vector<int64_t> &e)
{
	std::vector<Interval> sc(scope);
	std::stable_sort(sc.begin(), sc.end(), [](const Interval &a, const Interval &b) { return std::tie(a.chromid, a.start) < std::tie(b.chromid, b.start); });
	for (const Interval &iv : sc) {
		if (!c.empty() && c.back() == iv.chromid && e.back() >= iv.start) {
			if (e.back() < iv.end)
				e.back() = iv.end;
			continue;
		}
		c.push_back(iv.chromid);
		s.push_back(iv.start);
		e.push_back(iv.end);
	}
}

GpuGroupSession::GpuGroupSession(const std::vector<int> &devices, const std::vector<int64_t> &chrom_sizes)
{
	if (mg_group_init((int)devices.size(), devices.data(), (int)chrom_sizes.size(), chrom_sizes.data(), &m_g))
		throw MishaGpuException(mg_group_last_error(nullptr));
}

GpuGroupSession::~GpuGroupSession()
{
	for (mg_gdense *t : m_tracks)
		mg_group_dense_free(m_g, t);
	mg_group_shutdown(m_g);
}

void GpuGroupSession::check(int rc) const
{
	if (rc)
		throw MishaGpuException(mg_group_last_error(m_g));
}

mg_gdense *GpuGroupSession::dense_create(uint32_t bin_size)
{
	mg_gdense *t = nullptr;
	check(mg_group_dense_create(m_g, bin_size, &t));
	m_tracks.push_back(t);
	return t;
}

void GpuGroupSession::dense_stage(mg_gdense *t, int chromid, const float *bins, int64_t nbins) { check(mg_group_dense_stage(m_g, t, chromid, bins, nbins)); }

ExtractResult GpuGroupSession::gextract(const mg_gdense *t, const std::vector<std::pair<int, double>> &funcs, int64_t sshift, int64_t eshift,
										const std::vector<Interval> &scope)
{
	// the scope sorted by (chromid, start) as C_gextract does before scanning (src/GenomeTrackExtract.cpp:557-562); intervalID keeps
	// the caller's order
	std::vector<int64_t> order(scope.size());
	for (size_t i = 0; i < order.size(); ++i)
		order[i] = (int64_t)i;
	std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
		return std::tie(scope[a].chromid, scope[a].start) < std::tie(scope[b].chromid, scope[b].start);
	});
	ExtractResult res;
	const size_t n = scope.size();
	res.chrom.resize(n);
	res.start.resize(n);
	res.end.resize(n);
	res.interval_id.resize(n);
	for (size_t i = 0; i < n; ++i) {
		const Interval &iv = scope[order[i]];
		res.chrom[i] = iv.chromid;
		res.start[i] = iv.start;
		res.end[i] = iv.end;
		res.interval_id[i] = (int32_t)order[i] + 1;
	}
	std::vector<int> ids;
	std::vector<double> params;
	std::vector<double *> out;
	res.values.assign(funcs.size(), std::vector<double>(n));
	for (size_t k = 0; k < funcs.size(); ++k) {
		ids.push_back(funcs[k].first);
		params.push_back(funcs[k].second);
		out.push_back(res.values[k].data());
	}
	check(mg_group_h_dense_window(m_g, t, (int64_t)n, res.chrom.data(), res.start.data(), res.end.data(), sshift, eshift, (int)ids.size(), ids.data(),
								  params.data(), out.data()));
	return res;
}

void GpuGroupSession::gsummary(const mg_gdense *t, const std::vector<Interval> &scope, int64_t iterator_binsize, int func, double param,
							   int64_t sshift, int64_t eshift, double out[7])
{
	std::vector<int32_t> c;
	std::vector<int64_t> s, e;
	unified_scope(scope, c, s, e);
	double part[6];
	check(mg_group_dense_summary(m_g, t, iterator_binsize, (int64_t)c.size(), c.data(), s.data(), e.data(), sshift, eshift, func, param, part));
	mg_summary_finalize(part, out);
}

std::vector<uint64_t> GpuGroupSession::gdist(const mg_gdense *t, const std::vector<Interval> &scope, int64_t iterator_binsize, int func, double param,
											 int64_t sshift, int64_t eshift, const std::vector<double> &breaks, bool include_lowest)
{
	std::vector<int32_t> c;
	std::vector<int64_t> s, e;
	unified_scope(scope, c, s, e);
	std::vector<uint64_t> counts(breaks.size() > 1 ? breaks.size() - 1 : 0);
	check(mg_group_dense_hist(m_g, t, iterator_binsize, (int64_t)c.size(), c.data(), s.data(), e.data(), sshift, eshift, func, param, breaks.data(),
							  (int)breaks.size(), include_lowest ? 1 : 0, counts.data()));
	return counts;
}

std::vector<double> GpuGroupSession::gquantiles(const mg_gdense *t, const std::vector<Interval> &scope, int64_t iterator_binsize, int func,
												double param, int64_t sshift, int64_t eshift, const std::vector<double> &percentiles)
{
	std::vector<int32_t> c;
	std::vector<int64_t> s, e;
	unified_scope(scope, c, s, e);
	std::vector<double> q(percentiles.size());
	check(mg_group_dense_quantiles(m_g, t, iterator_binsize, (int64_t)c.size(), c.data(), s.data(), e.data(), sshift, eshift, func, param,
								   (int)percentiles.size(), percentiles.data(), q.data(), nullptr));
	return q;
}

} // namespace mishagpu
