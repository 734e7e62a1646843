// group.cuh -- several GPUs of one box behind ONE host thread: the multi-device half of the C ABI (mg_group_*).
//
// The reference scales by forking: prepare4multitasking cuts the scope by chromosome / range into one share per child process
// (src/rdbinterval.cpp:446-564), every child scans its share, the parent concatenates the extracted rows
// (src/GenomeTrackExtract.cpp:1461-1509) or merges the children's small partial results (IntervalSummary::merge,
// src/GenomeTrackSummary.cpp:205-216; the histogram sum, src/GenomeTrackDistribution.cpp:131-137; the percentile samples,
// src/GenomeTrackQuantiles.cpp:278-400). An R process is single-threaded, so the drop-in for that cannot be "one process per GPU":
// a group owns one mg_ctx per device and drives them asynchronously from the calling thread --
//   * every chromosome is staged on ONE member, its owner: whole chromosomes, longest-processing-time greedy over their lengths
//     (SURVEY.md 8(e)); work per member is proportional to the bins it owns;
//   * rows are routed to the owner of their chromosome. Sorted rows (every scanner the reference has produces them chromosome by
//     chromosome) fall into a few contiguous runs; each run is ENQUEUED on its owner's copy / kernel / copy pipeline and no member is
//     waited for before all of them have their work. Rows in any other order are gathered per member first and scattered back;
//   * gsummary / gdist partials (48 bytes / 8 bytes per bin) are computed where the data lives, copied to member 0 over NVLink
//     (cudaMemcpyPeerAsync) and merged there by one small kernel in member order -- the order the reference's parent merges its
//     children in -- before the single copy to the host;
//   * gquantiles: the radix-select passes of mg_quantiles run on every member's share of the value column, the per-pass digit
//     histograms (<= 1 MB) meet on the host, which already steers the selection: one exact global selection, no sampling.
// Nothing here touches the data path of a single device: the kernels are the ones the one-device entry points launch.
#pragma once

struct mg_group {
	std::vector<mg_ctx *> ctx;
	std::vector<int> owner; // chromid -> member
	std::vector<int64_t> chrom_sizes;
	std::string err;
	double *d_gather = nullptr;  // on member 0: room for every member's partial record / histogram
	int64_t gather_bytes = 0;
};
struct mg_gdense {
	std::vector<mg_dense *> t; // one per member; a member holds the chromosomes it owns
	uint32_t bin_size = 0;
};

static thread_local std::string g_mg_group_init_err;

static int mg_gfail(mg_group *g, const char *fmt, ...)
{
	char buf[1024];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	if (g)
		g->err = buf;
	else
		g_mg_group_init_err = buf;
	return 1;
}
// a member's failure becomes the group's
#define MG_GMEMBER(g, k, call)                                                                                          \
	do {                                                                                                               \
		if (call)                                                                                                      \
			return mg_gfail(g, "device %d: %s", (g)->ctx[k]->device, (g)->ctx[k]->err.c_str());                        \
	} while (0)
#define MG_GCUDA(g, call)                                                                                              \
	do {                                                                                                               \
		cudaError_t e__ = (call);                                                                                      \
		if (e__ != cudaSuccess)                                                                                        \
			return mg_gfail(g, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__);           \
	} while (0)
#define MG_GREQUIRE(g, cond, ...)                                                                                      \
	do {                                                                                                               \
		if (!(cond))                                                                                                   \
			return mg_gfail(g, __VA_ARGS__);                                                                           \
	} while (0)

extern "C" const char *mg_group_last_error(const mg_group *g) { return g ? g->err.c_str() : g_mg_group_init_err.c_str(); }
extern "C" int mg_group_size(const mg_group *g) { return (int)g->ctx.size(); }
extern "C" mg_ctx *mg_group_ctx(mg_group *g, int k) { return k >= 0 && k < (int)g->ctx.size() ? g->ctx[k] : nullptr; }
extern "C" int mg_group_owner(const mg_group *g, int chromid) { return chromid >= 0 && chromid < (int)g->owner.size() ? g->owner[chromid] : -1; }

extern "C" void mg_group_shutdown(mg_group *g)
{
	if (!g)
		return;
	if (g->d_gather && !g->ctx.empty()) {
		cudaSetDevice(g->ctx[0]->device);
		cudaFree(g->d_gather);
	}
	for (mg_ctx *c : g->ctx)
		mg_shutdown(c);
	delete g;
}

extern "C" int mg_group_init(int ndev, const int *dev_ids, int nchroms, const int64_t *chrom_sizes, mg_group **out)
{
	if (ndev < 1 || !dev_ids || nchroms < 1 || !chrom_sizes || !out)
		return mg_gfail(nullptr, "mg_group_init: bad arguments");
	mg_group *g = new (std::nothrow) mg_group();
	if (!g)
		return mg_gfail(nullptr, "mg_group_init: out of memory");
	for (int k = 0; k < ndev; ++k) {
		for (int j = 0; j < k; ++j)
			if (dev_ids[j] == dev_ids[k]) {
				mg_group_shutdown(g);
				return mg_gfail(nullptr, "mg_group_init: device %d listed twice", dev_ids[k]);
			}
		mg_ctx *c = nullptr;
		if (mg_init(dev_ids[k], nchroms, chrom_sizes, &c)) {
			mg_group_shutdown(g);
			return mg_gfail(nullptr, "mg_group_init: device %d: %s", dev_ids[k], mg_last_error(nullptr));
		}
		g->ctx.push_back(c);
	}
	// peer access towards member 0 (the partials travel there); a pair without it still works, the copy is then staged by the driver
	for (int k = 1; k < ndev; ++k) {
		int can = 0;
		if (cudaDeviceCanAccessPeer(&can, dev_ids[0], dev_ids[k]) == cudaSuccess && can) {
			cudaSetDevice(dev_ids[0]);
			if (cudaDeviceEnablePeerAccess(dev_ids[k], 0) != cudaSuccess)
				cudaGetLastError(); // already enabled
			cudaSetDevice(dev_ids[k]);
			if (cudaDeviceEnablePeerAccess(dev_ids[0], 0) != cudaSuccess)
				cudaGetLastError();
		}
	}
	// whole chromosomes, the longest first, each to the member with the least bins so far
	g->chrom_sizes.assign(chrom_sizes, chrom_sizes + nchroms);
	g->owner.assign(nchroms, 0);
	std::vector<int> order(nchroms);
	for (int i = 0; i < nchroms; ++i)
		order[i] = i;
	std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return chrom_sizes[a] > chrom_sizes[b]; });
	std::vector<int64_t> load(ndev, 0);
	for (int i : order) {
		int best = 0;
		for (int k = 1; k < ndev; ++k)
			if (load[k] < load[best])
				best = k;
		g->owner[i] = best;
		load[best] += chrom_sizes[i];
	}
	if (ndev > 1)
		for (int k = 0; k < ndev; ++k) { // what each member serves, for the device-side check of routed rows (k_rows_sanitize)
			std::vector<uint8_t> served((size_t)nchroms);
			for (int i = 0; i < nchroms; ++i)
				served[i] = g->owner[i] == k;
			mg_ctx *c = g->ctx[k];
			if (cudaSetDevice(c->device) != cudaSuccess || cudaMalloc(&c->d_served, (size_t)nchroms) != cudaSuccess ||
				cudaMemcpy(c->d_served, served.data(), (size_t)nchroms, cudaMemcpyHostToDevice) != cudaSuccess) {
				mg_group_shutdown(g);
				return mg_gfail(nullptr, "mg_group_init: device %d: %s", dev_ids[k], cudaGetErrorString(cudaGetLastError()));
			}
		}
	*out = g;
	return 0;
}

extern "C" int mg_group_set_option(mg_group *g, const char *name, int64_t value)
{
	for (size_t k = 0; k < g->ctx.size(); ++k)
		MG_GMEMBER(g, k, mg_set_option(g->ctx[k], name, value));
	return 0;
}

extern "C" int mg_group_sync(mg_group *g)
{
	for (mg_ctx *c : g->ctx) {
		MG_GCUDA(g, cudaSetDevice(c->device));
		MG_GCUDA(g, cudaDeviceSynchronize());
	}
	return 0;
}

// ---- dense tracks ---------------------------------------------------------------------------------------------------------
extern "C" void mg_group_dense_free(mg_group *g, mg_gdense *t)
{
	if (!t)
		return;
	for (size_t k = 0; k < t->t.size(); ++k)
		if (t->t[k]) {
			cudaSetDevice(g->ctx[k]->device);
			mg_dense_free(g->ctx[k], t->t[k]);
		}
	delete t;
}

extern "C" int mg_group_dense_create(mg_group *g, uint32_t bin_size, mg_gdense **out)
{
	mg_gdense *t = new (std::nothrow) mg_gdense();
	MG_GREQUIRE(g, t, "out of memory");
	t->bin_size = bin_size;
	t->t.assign(g->ctx.size(), nullptr);
	for (size_t k = 0; k < g->ctx.size(); ++k) {
		cudaSetDevice(g->ctx[k]->device);
		if (mg_dense_create(g->ctx[k], bin_size, &t->t[k])) {
			mg_gfail(g, "device %d: %s", g->ctx[k]->device, g->ctx[k]->err.c_str());
			mg_group_dense_free(g, t);
			return 1;
		}
	}
	*out = t;
	return 0;
}

extern "C" int mg_group_dense_stage(mg_group *g, mg_gdense *t, int chromid, const float *h_bins, int64_t nbins)
{
	MG_GREQUIRE(g, chromid >= 0 && chromid < (int)g->owner.size(), "mg_group_dense_stage: chromosome id %d out of range", chromid);
	const int k = g->owner[chromid];
	MG_GCUDA(g, cudaSetDevice(g->ctx[k]->device));
	MG_GMEMBER(g, k, mg_dense_stage(g->ctx[k], t->t[k], chromid, h_bins, nbins));
	return 0;
}

extern "C" int64_t mg_group_dense_bytes(const mg_gdense *t, int member) { return member >= 0 && member < (int)t->t.size() ? mg_dense_bytes(t->t[member]) : 0; }

// ---- routing ----------------------------------------------------------------------------------------------------------------
struct GroupRun {
	int member;
	int64_t first, count;
};
// first j > i with h_chrom[j] != h_chrom[i] (n when there is none): two ids per 64-bit compare, four compares per iteration -- the
// routing pass reads every id of the batch on the host, and for 10 M rows a per-id loop was as long as the batch's transfers
static int64_t mg_run_end(const int32_t *h_chrom, int64_t i, int64_t n)
{
	const uint32_t c = (uint32_t)h_chrom[i];
	const uint64_t pat = ((uint64_t)c << 32) | c;
	int64_t j = i + 1;
	while (j < n && (((uintptr_t)(h_chrom + j)) & 7u)) { // to an 8-byte boundary
		if ((uint32_t)h_chrom[j] != c)
			return j;
		++j;
	}
	const uint64_t *w = reinterpret_cast<const uint64_t *>(h_chrom + j);
	int64_t k = 0;
	const int64_t nw = (n - j) / 2;
	for (; k + 4 <= nw; k += 4)
		if (((w[k] ^ pat) | (w[k + 1] ^ pat) | (w[k + 2] ^ pat) | (w[k + 3] ^ pat)) != 0)
			break;
	j += 2 * k;
	while (j < n && (uint32_t)h_chrom[j] == c)
		++j;
	return j;
}

// maximal runs of rows with one owner; false when a chromosome id is out of range
static bool mg_group_runs(mg_group *g, int64_t n, const int32_t *h_chrom, std::vector<GroupRun> &runs, int64_t max_runs)
{
	runs.clear();
	const int nch = (int)g->owner.size();
	int64_t i = 0;
	while (i < n) {
		const int32_t c = h_chrom[i];
		if (c < 0 || c >= nch) {
			mg_gfail(g, "row %lld names chromosome id %d (the group holds %d)", (long long)i, (int)c, nch);
			return false;
		}
		const int m = g->owner[c];
		const int64_t j = mg_run_end(h_chrom, i, n);
		if (!runs.empty() && runs.back().member == m && runs.back().first + runs.back().count == i)
			runs.back().count += j - i; // the next chromosome has the same owner
		else {
			runs.push_back(GroupRun{m, i, j - i});
			if ((int64_t)runs.size() > max_runs)
				return true; // the caller switches to gather / scatter
		}
		i = j;
	}
	return true;
}

// The same for rows that come chromosome by chromosome in ascending id order (what every scanner of the reference produces from a sorted
// scope): the end of a run is found by bisection, O(chromosomes x log rows) instead of a pass over every id. Nothing here verifies the order --
// the members do, on the device: a row that reaches a member which does not own its chromosome is recorded (k_rows_sanitize) and the caller
// routes the batch again with the exact pass.
static bool mg_group_runs_sorted(mg_group *g, int64_t n, const int32_t *h_chrom, std::vector<GroupRun> &runs)
{
	runs.clear();
	const int nch = (int)g->owner.size();
	int64_t i = 0;
	while (i < n) {
		const int32_t c = h_chrom[i];
		if (c < 0 || c >= nch)
			return false;
		int64_t lo = i, hi = n; // first index > i whose id exceeds c, were the ids ascending
		while (hi - lo > 1) {
			const int64_t mid = lo + (hi - lo) / 2;
			if (h_chrom[mid] <= c)
				lo = mid;
			else
				hi = mid;
		}
		const int m = g->owner[c];
		if (!runs.empty() && runs.back().member == m)
			runs.back().count += hi - i;
		else
			runs.push_back(GroupRun{m, i, hi - i});
		i = hi;
	}
	return true;
}

// gextract: host rows in, host columns out (the group's mg_h_dense_window)
extern "C" int mg_group_h_dense_window(mg_group *g, const mg_gdense *t, int64_t n, const int32_t *h_chrom, const int64_t *h_start, const int64_t *h_end,
									   int64_t sshift, int64_t eshift, int nfuncs, const int *funcs, const double *params, double *const *h_out)
{
	MG_GREQUIRE(g, n >= 0 && nfuncs > 0 && nfuncs <= 32, "mg_group_h_dense_window: bad arguments");
	MG_GREQUIRE(g, n == 0 || (h_chrom && h_start && h_end && h_out), "mg_group_h_dense_window: NULL row or output arrays");
	const int nm = (int)g->ctx.size();
	const int64_t MAX_RUNS = 4096;
	std::vector<GroupRun> runs;
	// gathered batches (rows in no useful order): one per member
	std::vector<std::vector<int64_t>> idx;
	std::vector<std::vector<int32_t>> gc(nm);
	std::vector<std::vector<int64_t>> gs(nm), ge(nm);
	std::vector<std::vector<double>> gout;
	std::vector<char> used(nm, 0);

	// enqueue every run, then drain: no member is waited for before the others have their work. misrouted: a member met a row of a
	// chromosome it does not own, or an id out of range (only the trusting routing below can cause the former)
	auto execute = [&](bool gather, bool &misrouted) -> int {
		int rc = 0;
		std::fill(used.begin(), used.end(), 0);
		for (const GroupRun &r : runs) {
			mg_ctx *ctx = g->ctx[r.member];
			const mg_dense *tt = t->t[r.member];
			double *outp[32];
			const int32_t *pc;
			const int64_t *ps, *pe;
			if (gather) {
				pc = gc[r.member].data(); ps = gs[r.member].data(); pe = ge[r.member].data();
				for (int f = 0; f < nfuncs; ++f)
					outp[f] = gout[(size_t)r.member * nfuncs + f].data();
			} else {
				pc = h_chrom + r.first; ps = h_start + r.first; pe = h_end + r.first;
				for (int f = 0; f < nfuncs; ++f)
					outp[f] = h_out[f] + r.first;
			}
			used[r.member] = 1;
			if (cudaSetDevice(ctx->device) != cudaSuccess || mg_pipe_prepare(ctx, mg_pipe_slot_bytes(ctx, nfuncs)) ||
				mg_h_pipeline_run(ctx, r.count, pc, ps, pe, nfuncs, outp, [&](const mg_rows &rows, int64_t m, double *const *d_cols, cudaStream_t st) {
					DenseQuery q;
					bool need_sweep;
					return mg_fill_dense_query(ctx, tt, &rows, 0, m, sshift, eshift, nfuncs, funcs, params, d_cols, q, need_sweep) ||
						   mg_launch_dense(ctx, q, need_sweep, st);
				})) {
				rc = mg_gfail(g, "device %d: %s", ctx->device, ctx->err.empty() ? "cudaSetDevice failed" : ctx->err.c_str());
				break;
			}
		}
		misrouted = false;
		for (int k = 0; k < nm; ++k) { // whatever happened, nothing stays in flight into the caller's buffers
			if (!used[k])
				continue;
			mg_ctx *ctx = g->ctx[k];
			cudaSetDevice(ctx->device);
			for (int i = 0; i < MG_PIPE_MAX_SLOTS; ++i)
				if (ctx->pipe_stream[i]) {
					const cudaError_t e = cudaStreamSynchronize(ctx->pipe_stream[i]);
					if (e != cudaSuccess && !rc)
						rc = mg_gfail(g, "device %d: %s", ctx->device, cudaGetErrorString(e));
				}
			if (ctx->d_first_bad) { // the members' records of rows they should not have got (reset either way)
				unsigned long long bad = ~0ull;
				if (cudaMemcpy(&bad, ctx->d_first_bad, 8, cudaMemcpyDeviceToHost) == cudaSuccess && bad != ~0ull) {
					misrouted = true;
					cudaMemset(ctx->d_first_bad, 0xff, 8);
				}
			}
		}
		return rc;
	};

	// 1. trust the order (chromosome-major, ascending ids: what a sorted scope yields): routing by bisection, verified on the devices
	bool misrouted = false;
	if (mg_group_runs_sorted(g, n, h_chrom, runs) && (int64_t)runs.size() <= MAX_RUNS) {
		if (execute(false, misrouted))
			return 1;
		if (!misrouted)
			return 0;
	}
	// 2. any other order (or a bad id): the exact pass over every row's id; more owner changes than MAX_RUNS -> gather / scatter
	if (!mg_group_runs(g, n, h_chrom, runs, MAX_RUNS))
		return 1;
	const bool gather = (int64_t)runs.size() > MAX_RUNS;
	if (gather) {
		idx.assign(nm, {});
		const int nch = (int)g->owner.size();
		for (int64_t i = 0; i < n; ++i) {
			const int32_t c = h_chrom[i];
			MG_GREQUIRE(g, c >= 0 && c < nch, "row %lld names chromosome id %d (the group holds %d)", (long long)i, (int)c, nch);
			idx[g->owner[c]].push_back(i);
		}
		gout.assign((size_t)nm * nfuncs, {});
		runs.clear();
		for (int k = 0; k < nm; ++k) {
			const size_t m = idx[k].size();
			gc[k].resize(m); gs[k].resize(m); ge[k].resize(m);
			for (size_t j = 0; j < m; ++j) {
				gc[k][j] = h_chrom[idx[k][j]];
				gs[k][j] = h_start[idx[k][j]];
				ge[k][j] = h_end[idx[k][j]];
			}
			for (int f = 0; f < nfuncs; ++f)
				gout[(size_t)k * nfuncs + f].resize(m);
			if (m)
				runs.push_back(GroupRun{k, 0, (int64_t)m});
		}
	}
	if (execute(gather, misrouted))
		return 1;
	MG_GREQUIRE(g, !misrouted, "mg_group_h_dense_window: a member was handed rows of a chromosome it does not own (internal routing error)");
	if (gather)
		for (int k = 0; k < nm; ++k)
			for (int f = 0; f < nfuncs; ++f) {
				const std::vector<double> &col = gout[(size_t)k * nfuncs + f];
				for (size_t j = 0; j < col.size(); ++j)
					h_out[f][idx[k][j]] = col[j];
			}
	return 0;
}

// ---- gsummary / gdist / gquantiles over the fixed-bin iterator of a scope ---------------------------------------------------------
// The scope's intervals are dealt to the owners of their chromosomes; every member builds its implicit rows (mg_rows_fixedbin).
struct GroupScope {
	std::vector<mg_rows *> rows; // per member (nullptr: no share)
	mg_group *g = nullptr;
	~GroupScope()
	{
		for (size_t k = 0; k < rows.size(); ++k)
			if (rows[k]) {
				cudaSetDevice(g->ctx[k]->device);
				mg_rows_free(g->ctx[k], rows[k]);
			}
	}
};
static int mg_group_scope(mg_group *g, int64_t binsize, int64_t nscope, const int32_t *h_chrom, const int64_t *h_start, const int64_t *h_end, GroupScope &sc)
{
	const int nm = (int)g->ctx.size(), nch = (int)g->owner.size();
	sc.g = g;
	sc.rows.assign(nm, nullptr);
	std::vector<std::vector<int32_t>> c(nm);
	std::vector<std::vector<int64_t>> s(nm), e(nm);
	for (int64_t i = 0; i < nscope; ++i) {
		MG_GREQUIRE(g, h_chrom[i] >= 0 && h_chrom[i] < nch, "scope interval %lld names chromosome id %d", (long long)i, (int)h_chrom[i]);
		const int k = g->owner[h_chrom[i]];
		c[k].push_back(h_chrom[i]); s[k].push_back(h_start[i]); e[k].push_back(h_end[i]);
	}
	for (int k = 0; k < nm; ++k)
		if (!c[k].empty()) {
			MG_GCUDA(g, cudaSetDevice(g->ctx[k]->device));
			MG_GMEMBER(g, k, mg_rows_fixedbin(g->ctx[k], binsize, (int64_t)c[k].size(), c[k].data(), s[k].data(), e[k].data(), &sc.rows[k]));
		}
	return 0;
}
static int mg_group_gather_room(mg_group *g, int64_t bytes)
{
	if (g->gather_bytes >= bytes)
		return 0;
	MG_GCUDA(g, cudaSetDevice(g->ctx[0]->device));
	if (g->d_gather)
		cudaFree(g->d_gather);
	g->d_gather = nullptr;
	g->gather_bytes = 0;
	MG_GCUDA(g, cudaMalloc(&g->d_gather, (size_t)bytes));
	g->gather_bytes = bytes;
	return 0;
}

// records of {n, nn, sum, min, max, ssq} merged in member order (IntervalSummary::merge, src/GenomeTrackSummary.cpp:52-61)
__global__ void k_group_merge_summary(const double *__restrict__ parts, int nm, double *__restrict__ out)
{
	if (threadIdx.x || blockIdx.x)
		return;
	double n = 0, nn = 0, sum = 0, ssq = 0, mn = __longlong_as_double(0x7ff0000000000000LL), mx = -mn;
	for (int k = 0; k < nm; ++k) {
		const double *p = parts + 6 * k;
		n += p[0];
		nn += p[1];
		sum += p[2];
		mn = fmin(mn, p[3]);
		mx = fmax(mx, p[4]);
		ssq += p[5];
	}
	out[0] = n; out[1] = nn; out[2] = sum; out[3] = mn; out[4] = mx; out[5] = ssq;
}
__global__ void k_group_merge_hist(const unsigned long long *__restrict__ parts, int nm, int64_t cells, unsigned long long *__restrict__ out)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= cells)
		return;
	unsigned long long s = 0;
	for (int k = 0; k < nm; ++k)
		s += parts[(int64_t)k * cells + i];
	out[i] = s;
}

// per-member scratch that is released when the call ends
struct GroupScratch {
	mg_group *g;
	std::vector<std::pair<int, void *>> ptrs;
	explicit GroupScratch(mg_group *gg) : g(gg) {}
	~GroupScratch()
	{
		for (auto &p : ptrs) {
			cudaSetDevice(g->ctx[p.first]->device);
			cudaFree(p.second);
		}
	}
	cudaError_t alloc(int member, void **out, size_t bytes)
	{
		cudaError_t e = cudaSetDevice(g->ctx[member]->device);
		if (e != cudaSuccess)
			return e;
		e = cudaMalloc(out, bytes ? bytes : 1);
		if (e == cudaSuccess)
			ptrs.push_back({member, *out});
		return e;
	}
};

// gsummary of one dense-track function under a window over the fixed-bin iterator `binsize` of the scope: h_partial[6]
extern "C" int mg_group_dense_summary(mg_group *g, const mg_gdense *t, int64_t binsize, int64_t nscope, const int32_t *h_chrom, const int64_t *h_start,
									  const int64_t *h_end, int64_t sshift, int64_t eshift, int func, double param, double *h_partial)
{
	const int nm = (int)g->ctx.size();
	GroupScope sc;
	if (mg_group_scope(g, binsize, nscope, h_chrom, h_start, h_end, sc) || mg_group_gather_room(g, (int64_t)(nm + 1) * 48))
		return 1;
	GroupScratch scratch(g);
	std::vector<double *> part(nm, nullptr);
	for (int k = 0; k < nm; ++k) { // every member gets its launch before any is waited for
		MG_GCUDA(g, scratch.alloc(k, (void **)&part[k], 48));
		MG_GMEMBER(g, k, mg_summary_init(g->ctx[k], part[k], nullptr));
		if (sc.rows[k])
			MG_GMEMBER(g, k, mg_dense_summary(g->ctx[k], t->t[k], sc.rows[k], 0, mg_rows_count(sc.rows[k]), sshift, eshift, func, param, part[k], nullptr));
	}
	for (int k = 0; k < nm; ++k) { // 48 bytes per member to member 0, over NVLink where the pair has peer access
		MG_GCUDA(g, cudaSetDevice(g->ctx[k]->device));
		MG_GCUDA(g, cudaMemcpyPeerAsync(g->d_gather + 6 * k, g->ctx[0]->device, part[k], g->ctx[k]->device, 48, nullptr));
		MG_GCUDA(g, cudaStreamSynchronize(nullptr));
	}
	MG_GCUDA(g, cudaSetDevice(g->ctx[0]->device));
	k_group_merge_summary<<<1, 32>>>(g->d_gather, nm, g->d_gather + 6 * nm);
	MG_GCUDA(g, cudaGetLastError());
	g->ctx[0]->launches++;
	MG_GCUDA(g, cudaMemcpy(h_partial, g->d_gather + 6 * nm, 48, cudaMemcpyDeviceToHost));
	return 0;
}

// gdist (one dimension) of one dense-track function: h_counts[nbreaks - 1]
extern "C" int mg_group_dense_hist(mg_group *g, const mg_gdense *t, int64_t binsize, int64_t nscope, const int32_t *h_chrom, const int64_t *h_start,
								   const int64_t *h_end, int64_t sshift, int64_t eshift, int func, double param, const double *h_breaks, int nbreaks,
								   int include_lowest, uint64_t *h_counts)
{
	MG_GREQUIRE(g, nbreaks >= 2 && h_breaks && h_counts, "mg_group_dense_hist: bad arguments");
	const int nm = (int)g->ctx.size();
	const int64_t cells = nbreaks - 1;
	GroupScope sc;
	if (mg_group_scope(g, binsize, nscope, h_chrom, h_start, h_end, sc) || mg_group_gather_room(g, (int64_t)(nm + 1) * cells * 8))
		return 1;
	GroupScratch scratch(g);
	std::vector<uint64_t *> cnt(nm, nullptr);
	for (int k = 0; k < nm; ++k) {
		MG_GCUDA(g, scratch.alloc(k, (void **)&cnt[k], (size_t)cells * 8));
		MG_GCUDA(g, cudaMemsetAsync(cnt[k], 0, (size_t)cells * 8, nullptr));
		if (sc.rows[k])
			MG_GMEMBER(g, k, mg_dense_hist(g->ctx[k], t->t[k], sc.rows[k], 0, mg_rows_count(sc.rows[k]), sshift, eshift, func, param, h_breaks, nbreaks,
										   include_lowest, cnt[k], nullptr));
	}
	unsigned long long *gat = (unsigned long long *)g->d_gather;
	for (int k = 0; k < nm; ++k) {
		MG_GCUDA(g, cudaSetDevice(g->ctx[k]->device));
		MG_GCUDA(g, cudaMemcpyPeerAsync(gat + (int64_t)k * cells, g->ctx[0]->device, cnt[k], g->ctx[k]->device, (size_t)cells * 8, nullptr));
		MG_GCUDA(g, cudaStreamSynchronize(nullptr));
	}
	MG_GCUDA(g, cudaSetDevice(g->ctx[0]->device));
	k_group_merge_hist<<<(unsigned)mg_div_up(cells, 256), 256>>>(gat, nm, cells, gat + (int64_t)nm * cells);
	MG_GCUDA(g, cudaGetLastError());
	g->ctx[0]->launches++;
	MG_GCUDA(g, cudaMemcpy(h_counts, gat + (int64_t)nm * cells, (size_t)cells * 8, cudaMemcpyDeviceToHost));
	return 0;
}

// gquantiles of one dense-track function: every member materialises the column of its rows, the selection runs over all of them
extern "C" int mg_group_dense_quantiles(mg_group *g, const mg_gdense *t, int64_t binsize, int64_t nscope, const int32_t *h_chrom, const int64_t *h_start,
										const int64_t *h_end, int64_t sshift, int64_t eshift, int func, double param, int nq, const double *h_percentiles,
										double *h_out, int64_t *h_nkept)
{
	const int nm = (int)g->ctx.size();
	GroupScope sc;
	if (mg_group_scope(g, binsize, nscope, h_chrom, h_start, h_end, sc))
		return 1;
	GroupScratch scratch(g);
	std::vector<mg_ctx *> ctxs;
	std::vector<const double *> cols;
	std::vector<int64_t> ns;
	for (int k = 0; k < nm; ++k) {
		if (!sc.rows[k])
			continue;
		const int64_t n = mg_rows_count(sc.rows[k]);
		double *col = nullptr;
		MG_GCUDA(g, scratch.alloc(k, (void **)&col, (size_t)n * 8));
		MG_GMEMBER(g, k, mg_dense_window(g->ctx[k], t->t[k], sc.rows[k], 0, n, sshift, eshift, 1, &func, &param, &col, nullptr));
		ctxs.push_back(g->ctx[k]);
		cols.push_back(col);
		ns.push_back(n);
	}
	if (ctxs.empty()) {
		for (int q = 0; q < nq; ++q)
			h_out[q] = std::numeric_limits<double>::quiet_NaN();
		if (h_nkept)
			*h_nkept = 0;
		return 0;
	}
	if (mg_quantiles_multi((int)ctxs.size(), ctxs.data(), cols.data(), ns.data(), nq, h_percentiles, h_out, h_nkept))
		return mg_gfail(g, "%s", ctxs[0]->err.c_str());
	return 0;
}
