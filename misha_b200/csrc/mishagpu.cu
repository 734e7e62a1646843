// mishagpu.cu -- the C ABI declared in include/mishagpu.h: context, staging, row sources and kernel dispatch.
// One translation unit; kernels live in the .cuh files next to it. Built by misha_b200/build.py with
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 --shared ...  ->  misha_b200/libmishagpu.so
#include "common.cuh"
#include "rows.cuh"
#include "dense.cuh"
#include "dense_index.cuh"
#include "dense_tile.cuh"
#include "sparse.cuh"
#include "stats.cuh"
#include "pwm.cuh"

thread_local std::string g_mg_init_err;

// ---- context ---------------------------------------------------------------------------------------------
extern "C" int mg_abi_version(void) { return MG_ABI_VERSION; }

extern "C" int mg_init(int device, int nchroms, const int64_t *chrom_sizes, mg_ctx **out)
{
	if (!out)
		return mg_fail(nullptr, "mg_init: out is NULL");
	*out = nullptr;
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev == 0)
		return mg_fail(nullptr, "mg_init: no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
	if (device < 0 || device >= ndev)
		return mg_fail(nullptr, "mg_init: device %d out of range (%d visible)", device, ndev);
	if (nchroms <= 0 || !chrom_sizes)
		return mg_fail(nullptr, "mg_init: need at least one chromosome");
	mg_ctx *ctx = new (std::nothrow) mg_ctx();
	if (!ctx)
		return mg_fail(nullptr, "mg_init: out of memory");
	ctx->device = device;
	cudaDeviceProp prop;
	if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
		delete ctx;
		return mg_fail(nullptr, "mg_init: %s", cudaGetErrorString(e));
	}
	{ // staging scratch comes from the stream-ordered pool: keep what it frees (re-mapping memory per chromosome costs seconds)
		cudaMemPool_t pool;
		if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
			uint64_t keep = ~0ull;
			cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
		}
	}
	ctx->sm_count = prop.multiProcessorCount;
	ctx->cc = prop.major * 10 + prop.minor;
	ctx->hbm_bytes = (int64_t)prop.totalGlobalMem;
	ctx->nchroms = nchroms;
	ctx->chrom_sizes.assign(chrom_sizes, chrom_sizes + nchroms);
	if ((e = cudaMalloc(&ctx->d_chrom_sizes, sizeof(int64_t) * nchroms)) != cudaSuccess ||
		(e = cudaMemcpy(ctx->d_chrom_sizes, chrom_sizes, sizeof(int64_t) * nchroms, cudaMemcpyHostToDevice)) != cudaSuccess) {
		delete ctx;
		return mg_fail(nullptr, "mg_init: %s", cudaGetErrorString(e));
	}
	*out = ctx;
	return 0;
}

extern "C" void mg_shutdown(mg_ctx *ctx)
{
	if (!ctx)
		return;
	cudaSetDevice(ctx->device);
	for (int i = 0; i < MG_PIPE_MAX_SLOTS; ++i) {
		if (ctx->pipe_dev[i])
			cudaFree(ctx->pipe_dev[i]);
		if (ctx->pipe_stream[i])
			cudaStreamDestroy(ctx->pipe_stream[i]);
	}
	if (ctx->d_first_bad)
		cudaFree(ctx->d_first_bad);
	if (ctx->d_served)
		cudaFree(ctx->d_served);
	for (auto &kv : ctx->stream_scratch)
		cudaFree(kv.second);
	cudaFree(ctx->d_chrom_sizes);
	delete ctx;
}

extern "C" const char *mg_last_error(const mg_ctx *ctx) { return ctx ? ctx->err.c_str() : g_mg_init_err.c_str(); }

extern "C" int mg_sync(mg_ctx *ctx, void *stream)
{
	MG_CUDA(ctx, cudaStreamSynchronize(mg_stream(stream)));
	return 0;
}

extern "C" int mg_set_option(mg_ctx *ctx, const char *name, int64_t value)
{
	if (name && !strcmp(name, "force_sweep")) {
		ctx->force_sweep = value != 0;
		return 0;
	}
	if (name && !strcmp(name, "pwm_strict")) {
		ctx->pwm_strict = value != 0;
		return 0;
	}
	if (name && !strcmp(name, "prefetch")) {
		ctx->prefetch = value != 0;
		return 0;
	}
	if (name && !strcmp(name, "stage_slots")) {
		ctx->stage_records = value != 0; // 0: the per-row kernel walks the block records through L1 (ablation); the slot count is fixed
		return 0;
	}
	if (name && !strcmp(name, "merge_rows")) { // 8: the sparse / coverage merge kernels always take 8 rows per thread (ablation)
		ctx->merge_rows = (int)value;
		return 0;
	}
	if (name && !strcmp(name, "hist_red")) {
		ctx->hist_red = value != 0;
		return 0;
	}
	if (name && !strcmp(name, "cov_raster")) {
		ctx->cov_raster = value != 0;
		return 0;
	}
	if (name && !strcmp(name, "hbm_budget_mb")) { // what the dense tracks may hold on this device; least-recently-used index groups are dropped beyond it
		ctx->hbm_budget = value > 0 ? value << 20 : 0;
		return 0;
	}
	if (name && !strcmp(name, "inline_table")) { // 0: the tile kernels load chromosome sizes / array bases from the tables in HBM (ablation)
		ctx->inline_table = value != 0;
		return 0;
	}
	if (name && !strcmp(name, "tile_kernel")) {
		ctx->tile_kernel = (int)value;
		return 0;
	}
	if (name && !strcmp(name, "block_wm")) {
		ctx->block_wm = value != 0;
		return 0;
	}
	if (name && !strcmp(name, "l2_fetch_granularity")) { // 32 / 64 / 128 bytes: how much L2 pulls from HBM per missed sector
		MG_CUDA(ctx, cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)value));
		return 0;
	}
	if (name && !strcmp(name, "zero_copy")) {
		ctx->zero_copy = value != 0;
		return 0;
	}
	if (name && !strcmp(name, "pipe_slots")) {
		if (value < 2 || value > MG_PIPE_MAX_SLOTS)
			return mg_fail(ctx, "mg_set_option: pipe_slots must be 2 .. %d", MG_PIPE_MAX_SLOTS);
		ctx->pipe_slots = (int)value;
		return 0;
	}
	if (name && !strcmp(name, "pipe_chunk_rows")) {
		if (value < 1024)
			return mg_fail(ctx, "mg_set_option: pipe_chunk_rows must be >= 1024");
		ctx->pipe_chunk_rows = value;
		return 0;
	}
	return mg_fail(ctx, "mg_set_option: unknown option '%s'", name ? name : "(null)");
}

extern "C" int64_t mg_launch_count(const mg_ctx *ctx) { return ctx->launches; }

extern "C" int mg_device_info(const mg_ctx *ctx, int *sm_count, int64_t *hbm_bytes, int *cc)
{
	if (sm_count) *sm_count = ctx->sm_count;
	if (hbm_bytes) *hbm_bytes = ctx->hbm_bytes;
	if (cc) *cc = ctx->cc;
	return 0;
}

extern "C" int mg_dev_alloc(mg_ctx *ctx, int64_t bytes, void **d_ptr)
{
	MG_CUDA(ctx, cudaSetDevice(ctx->device));
	MG_CUDA(ctx, cudaMalloc(d_ptr, (size_t)(bytes > 0 ? bytes : 1)));
	return 0;
}
extern "C" int mg_dev_free(mg_ctx *ctx, void *d_ptr)
{
	MG_CUDA(ctx, cudaFree(d_ptr));
	return 0;
}
extern "C" int mg_host_alloc(mg_ctx *ctx, int64_t bytes, void **h_ptr)
{
	MG_CUDA(ctx, cudaMallocHost(h_ptr, (size_t)(bytes > 0 ? bytes : 1)));
	return 0;
}
extern "C" int mg_host_free(mg_ctx *ctx, void *h_ptr)
{
	MG_CUDA(ctx, cudaFreeHost(h_ptr));
	return 0;
}
extern "C" int mg_copy_h2d(mg_ctx *ctx, void *d_dst, const void *h_src, int64_t bytes, void *stream)
{
	MG_CUDA(ctx, cudaMemcpyAsync(d_dst, h_src, (size_t)bytes, cudaMemcpyHostToDevice, mg_stream(stream)));
	return 0;
}
extern "C" int mg_copy_d2h(mg_ctx *ctx, void *h_dst, const void *d_src, int64_t bytes, void *stream)
{
	MG_CUDA(ctx, cudaMemcpyAsync(h_dst, d_src, (size_t)bytes, cudaMemcpyDeviceToHost, mg_stream(stream)));
	return 0;
}
extern "C" int mg_dev_memset(mg_ctx *ctx, void *d_ptr, int value, int64_t bytes, void *stream)
{
	MG_CUDA(ctx, cudaMemsetAsync(d_ptr, value, (size_t)bytes, mg_stream(stream)));
	return 0;
}

// Stream-ordered scratch of one entry point: whatever exit path the function takes (every MG_CUDA / MG_REQUIRE is a return), what
// was allocated is handed back to the pool in stream order.
struct MgScratch {
	cudaStream_t st;
	std::vector<void *> ptrs;
	explicit MgScratch(cudaStream_t s) : st(s) {}
	MgScratch(const MgScratch &) = delete;
	MgScratch &operator=(const MgScratch &) = delete;
	~MgScratch()
	{
		for (void *p : ptrs)
			cudaFreeAsync(p, st);
	}
	template <typename T>
	cudaError_t alloc(T **out, size_t bytes)
	{
		void *p = nullptr;
		const cudaError_t e = cudaMallocAsync(&p, bytes ? bytes : 1, st);
		if (e == cudaSuccess)
			ptrs.push_back(p);
		*out = (T *)p;
		return e;
	}
};

// Bump allocation out of large chunks: a staged dense chromosome is ~40 arrays, and cudaMalloc of hundreds of megabytes is
// slow enough (it maps physical memory) that one chunk per chromosome instead of one call per array takes seconds off staging.
struct MgArena {
	char *cur = nullptr;
	size_t left = 0, chunk = 0;
};

template <typename T>
static int mg_alloc_tracked(mg_ctx *ctx, std::vector<void *> &allocs, int64_t &bytes, int64_t count, T **out, MgArena *arena = nullptr)
{
	void *p = nullptr;
	size_t sz = sizeof(T) * (size_t)(count > 0 ? count : 1);
	if (arena) {
		sz = (sz + 255) & ~(size_t)255;
		if (sz > arena->left) {
			const size_t chunk = std::max(sz, arena->chunk);
			MG_CUDA(ctx, cudaMalloc(&p, chunk));
			allocs.push_back(p);
			arena->cur = (char *)p;
			arena->left = chunk;
		}
		p = arena->cur;
		arena->cur += sz;
		arena->left -= sz;
		bytes += (int64_t)sz;
		*out = (T *)p;
		return 0;
	}
	MG_CUDA(ctx, cudaMalloc(&p, sz));
	allocs.push_back(p);
	bytes += (int64_t)sz;
	*out = (T *)p;
	return 0;
}

// ---- dense staging ---------------------------------------------------------------------------------------
extern "C" int mg_dense_create(mg_ctx *ctx, uint32_t bin_size, mg_dense **out)
{
	MG_REQUIRE(ctx, bin_size > 0, "mg_dense_create: bin size must be positive");
	mg_dense *t = new (std::nothrow) mg_dense();
	MG_REQUIRE(ctx, t, "out of memory");
	t->bin_size = bin_size;
	t->nchroms = ctx->nchroms;
	{
		DenseChrom z;
		memset(&z, 0, sizeof(z));
		t->h_table.assign(ctx->nchroms, z);
	}
	if (mg_alloc_tracked(ctx, t->allocs, t->bytes, ctx->nchroms, &t->d_table)) {
		delete t;
		return 1;
	}
	cudaError_t e = cudaMemcpy(t->d_table, t->h_table.data(), sizeof(DenseChrom) * ctx->nchroms, cudaMemcpyHostToDevice);
	if (e != cudaSuccess) {
		mg_dense_free(ctx, t);
		return mg_fail(ctx, "mg_dense_create: %s", cudaGetErrorString(e));
	}
	ctx->dense_tracks.push_back(t);
	*out = t;
	return 0;
}

// min/max pyramids + wavelet matrix of one staged chromosome (dense_index.cuh)
// MISHAGPU_STAGE_TIMING=1 prints the phases of the index build (development aid)
struct StagePhase {
	const char *name;
	std::chrono::steady_clock::time_point t0;
	bool on;
	explicit StagePhase(const char *n) : name(n), t0(std::chrono::steady_clock::now()), on(getenv("MISHAGPU_STAGE_TIMING") != nullptr) {}
	~StagePhase()
	{
		if (!on)
			return;
		cudaDeviceSynchronize();
		fprintf(stderr, "[stage] %-28s %8.2f ms\n", name, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
	}
};

// ---- index groups: built per staged chromosome, on first use (mg_dense_ensure) --------------------------------------------------
// allocation into a group of the track
template <typename T>
static int mg_group_alloc(mg_ctx *ctx, mg_dense *t, int g, int64_t count, T **out, MgArena *arena)
{
	return mg_alloc_tracked(ctx, t->group[g].allocs, t->group[g].bytes, count, out, arena);
}

// the per-bin prefix {sum, count} (pref, nbins + 1 entries) and / or the prefix of squares (pref2): staging scratch for p8 and the
// chromosome-wide wavelet matrix, and the stddev group itself. Either output may be null.
static int mg_dense_prefix_pass(mg_ctx *ctx, float *vals, int64_t nbins, Pref *pref, double *pref2, double *totals /* host: sum |v|, sum v^2; may be null */)
{
	const int64_t ntiles = mg_div_up(nbins + 1, MG_ST_TILE); // +1: the closing entry pref[nbins]
	double *tile_sum = nullptr, *tile_abs = nullptr, *tile_sq = nullptr, *abs_total = nullptr;
	long long *tile_cnt = nullptr;
	MgScratch scratch(nullptr);
	MG_CUDA(ctx, scratch.alloc(&tile_sum, sizeof(double) * ntiles));
	MG_CUDA(ctx, scratch.alloc(&tile_abs, sizeof(double) * ntiles));
	MG_CUDA(ctx, scratch.alloc(&tile_sq, sizeof(double) * ntiles));
	MG_CUDA(ctx, scratch.alloc(&tile_cnt, sizeof(long long) * ntiles));
	MG_CUDA(ctx, scratch.alloc(&abs_total, sizeof(double) * 2));
	k_dense_clean_reduce<<<(unsigned)ntiles, MG_ST_THREADS>>>(vals, nbins, tile_sum, tile_cnt, tile_abs, tile_sq);
	MG_LAUNCH_CHECK(ctx);
	k_dense_tile_scan<<<1, 1024>>>(tile_sum, tile_cnt, tile_abs, tile_sq, ntiles, abs_total);
	MG_LAUNCH_CHECK(ctx);
	if (pref || pref2) {
		k_dense_prefix_write<<<(unsigned)ntiles, MG_ST_THREADS>>>(vals, nbins, tile_sum, tile_cnt, tile_sq, pref, pref2);
		MG_LAUNCH_CHECK(ctx);
	}
	if (totals)
		MG_CUDA(ctx, cudaMemcpy(totals, abs_total, sizeof(double) * 2, cudaMemcpyDeviceToHost));
	return 0;
}

static int mg_build_sd(mg_ctx *ctx, mg_dense *t, DenseChrom &dc, MgArena &arena)
{
	StagePhase ph("prefix of squares");
	double *pref2 = nullptr;
	if (mg_group_alloc(ctx, t, MG_G_SD, dc.nbins + 1, &pref2, &arena) || mg_dense_prefix_pass(ctx, const_cast<float *>(dc.vals), dc.nbins, nullptr, pref2, nullptr))
		return 1;
	dc.pref2 = pref2;
	return 0;
}

static int mg_build_ext(mg_ctx *ctx, mg_dense *t, DenseChrom &dc, MgArena &arena)
{
	const int64_t nbins = dc.nbins;
	// pyramids: level j+1 exists while level j has more than one group of 8
	{
		StagePhase ph("pyramids");
		const float *cmax = dc.vals, *cmin = dc.vals;
		int64_t nchild = nbins;
		for (int lv = 0; lv < MG_PYR_MAX && nchild > 8; ++lv) {
			const int64_t nparent = mg_div_up(nchild, 8), padded = (nparent + 7) / 8 * 8 + 8;
			float *pmax = nullptr, *pmin = nullptr;
			if (mg_group_alloc(ctx, t, MG_G_EXT, padded, &pmax, &arena) || mg_group_alloc(ctx, t, MG_G_EXT, padded, &pmin, &arena))
				return 1;
			k_pyr_build<<<(unsigned)mg_div_up(padded, 256), 256>>>(cmax, cmin, nchild, pmax, pmin, padded);
			MG_LAUNCH_CHECK(ctx);
			dc.pmax[lv] = pmax;
			dc.pmin[lv] = pmin;
			cmax = pmax;
			cmin = pmin;
			nchild = nparent;
		}
	}
	// sparse tables over the groups of 8 (level 0 = pyramid level 1)
	StagePhase ph("sparse tables");
	const int64_t ng = mg_div_up(nbins > 0 ? nbins : 1, 8);
	dc.ngroups = ng;
	// every level in ONE slab per extremum, pitch entries apart (level 0 is a copy of pyramid level 0): a kernel that holds the slab's base and
	// the pitch in its parameters (DenseQuery::inl_smax) reaches smax[j][g] without first loading the level's pointer from the table
	int nlev = 1;
	while (nlev < MG_ST_LEVELS && (1ll << nlev) <= ng)
		++nlev;
	const int64_t pitch = (ng + 31) / 32 * 32;
	float *slab_max = nullptr, *slab_min = nullptr;
	if (mg_group_alloc(ctx, t, MG_G_EXT, pitch * nlev, &slab_max, &arena) || mg_group_alloc(ctx, t, MG_G_EXT, pitch * nlev, &slab_min, &arena))
		return 1;
	if (!dc.pmax[0]) { // chromosomes of <= 8 bins have no pyramid level: build the single group entry
		float *pmax = nullptr, *pmin = nullptr;
		if (mg_group_alloc(ctx, t, MG_G_EXT, 16, &pmax, &arena) || mg_group_alloc(ctx, t, MG_G_EXT, 16, &pmin, &arena))
			return 1;
		k_pyr_build<<<1, 256>>>(dc.vals, dc.vals, nbins, pmax, pmin, 16);
		MG_LAUNCH_CHECK(ctx);
		MG_CUDA(ctx, cudaMemcpyAsync(slab_max, pmax, sizeof(float) * ng, cudaMemcpyDeviceToDevice));
		MG_CUDA(ctx, cudaMemcpyAsync(slab_min, pmin, sizeof(float) * ng, cudaMemcpyDeviceToDevice));
	} else {
		MG_CUDA(ctx, cudaMemcpyAsync(slab_max, dc.pmax[0], sizeof(float) * ng, cudaMemcpyDeviceToDevice));
		MG_CUDA(ctx, cudaMemcpyAsync(slab_min, dc.pmin[0], sizeof(float) * ng, cudaMemcpyDeviceToDevice));
	}
	dc.smax[0] = slab_max;
	dc.smin[0] = slab_min;
	dc.st_pitch = pitch;
	for (int j = 1; j < MG_ST_LEVELS; ++j) {
		if (j >= nlev) { // never selected: a run of len groups uses level floor(log2 len)
			dc.smax[j] = dc.smax[j - 1];
			dc.smin[j] = dc.smin[j - 1];
			continue;
		}
		float *smax = slab_max + pitch * j, *smin = slab_min + pitch * j;
		k_st_build<<<(unsigned)mg_div_up(ng, 256), 256>>>(dc.smax[j - 1], dc.smin[j - 1], ng, 1ll << (j - 1), smax, smin);
		MG_LAUNCH_CHECK(ctx);
		dc.smax[j] = smax;
		dc.smin[j] = smin;
	}
	return 0;
}

static int mg_build_lse(mg_ctx *ctx, mg_dense *t, DenseChrom &dc, MgArena &arena)
{
	// lse pyramid (double): level j+1 exists while level j has more than one group of 8
	StagePhase ph("lse pyramid");
	const double *child = nullptr;
	int64_t nchild = dc.nbins;
	for (int lv = 0; lv < MG_PYR_MAX && nchild > 8; ++lv) {
		const int64_t nparent = mg_div_up(nchild, 8);
		double *pl = nullptr;
		if (mg_group_alloc(ctx, t, MG_G_LSE, nparent, &pl, &arena))
			return 1;
		if (lv == 0)
			k_lse_build<float><<<(unsigned)mg_div_up(nparent, 256), 256>>>(dc.vals, nchild, pl, nparent);
		else
			k_lse_build<double><<<(unsigned)mg_div_up(nparent, 256), 256>>>(child, nchild, pl, nparent);
		MG_LAUNCH_CHECK(ctx);
		dc.plse[lv] = pl;
		child = pl;
		nchild = nparent;
	}
	return 0;
}

static int mg_build_bw(mg_ctx *ctx, mg_dense *t, DenseChrom &dc, MgArena &arena)
{
	// block-local wavelet matrices (quantiles of windows <= MG_BW_S bins)
	StagePhase ph("block wavelet matrices");
	const int64_t nblk = mg_div_up(dc.nbins > 0 ? dc.nbins : 1, MG_BW_STRIDE);
	uint8_t *bw = nullptr;
	if (mg_group_alloc(ctx, t, MG_G_BW, nblk * MG_BW_RECORD_BYTES, &bw, &arena))
		return 1;
	k_bw_build<<<(unsigned)nblk, MG_BW_BUILD_THREADS>>>(dc.vals, dc.nbins, bw);
	MG_LAUNCH_CHECK(ctx);
	dc.bw = bw;
	return 0;
}

static int mg_build_wm(mg_ctx *ctx, mg_dense *t, DenseChrom &dc, MgArena &arena)
{
	// wavelet matrix over the dense ranks of the non-NaN values
	StagePhase ph_wm("chromosome wavelet matrix");
	const int64_t nbins = dc.nbins;
	MgScratch tmp(nullptr);
	Pref *pref = nullptr; // per-bin prefix: where each non-NaN bin lands in the compacted order
	MG_CUDA(ctx, tmp.alloc(&pref, sizeof(Pref) * (size_t)(nbins + 1)));
	if (mg_dense_prefix_pass(ctx, const_cast<float *>(dc.vals), nbins, pref, nullptr, nullptr))
		return 1;
	long long m = 0;
	MG_CUDA(ctx, cudaMemcpy(&m, &pref[nbins].cnt, sizeof(long long), cudaMemcpyDeviceToHost));
	dc.wm_m = m;
	dc.wm_blocks = m / 32 + 1; // position m itself must be addressable
	int nlevels = 0;
	while (nlevels < 32 && (1ll << nlevels) < m)
		++nlevels;
	dc.wm_levels = nlevels;
	uint2 *wm = nullptr;
	uint32_t *zeros = nullptr;
	float *sorted = nullptr;
	if (mg_group_alloc(ctx, t, MG_G_WM, dc.wm_blocks * (nlevels > 0 ? nlevels : 1), &wm, &arena) ||
		mg_group_alloc(ctx, t, MG_G_WM, MG_WM_LEVELS, &zeros, &arena) || mg_group_alloc(ctx, t, MG_G_WM, m + 16, &sorted, &arena))
		return 1;
	dc.wm = wm;
	dc.wm_zeros = zeros;
	dc.wm_sorted = sorted;
	const size_t mm = (size_t)(m > 0 ? m : 1);
	uint32_t *keys[2] = {nullptr, nullptr}, *pos[2] = {nullptr, nullptr}, *d_zero = nullptr;
	uint2 *scratch = nullptr;
	MG_CUDA(ctx, tmp.alloc(&keys[0], sizeof(uint32_t) * mm));
	MG_CUDA(ctx, tmp.alloc(&keys[1], sizeof(uint32_t) * mm));
	MG_CUDA(ctx, tmp.alloc(&pos[0], sizeof(uint32_t) * mm));
	MG_CUDA(ctx, tmp.alloc(&pos[1], sizeof(uint32_t) * mm));
	MG_CUDA(ctx, tmp.alloc(&scratch, sizeof(uint2) * dc.wm_blocks));
	MG_CUDA(ctx, tmp.alloc(&d_zero, sizeof(uint32_t)));
	if (nbins) {
		k_wm_compact<<<(unsigned)mg_div_up(nbins, 256), 256>>>(dc.vals, pref, nbins, keys[0], pos[0]);
		MG_LAUNCH_CHECK(ctx);
	}
	const unsigned gm = (unsigned)mg_div_up(m > 0 ? m : 1, 256), gb = (unsigned)mg_div_up(dc.wm_blocks, 256);
	// (1) rank the values: LSD bit sort of (key, position) pairs, one stable partition per non-constant bit
	int cur = 0;
	for (int bit = 0; bit < 32 && m > 1; ++bit) {
		uint32_t z = 0;
		k_wm_bits<<<gb, 256>>>(keys[cur], m, bit, scratch, dc.wm_blocks);
		MG_LAUNCH_CHECK(ctx);
		k_wm_scan<<<1, 1024>>>(scratch, dc.wm_blocks, m, d_zero);
		MG_LAUNCH_CHECK(ctx);
		MG_CUDA(ctx, cudaMemcpy(&z, d_zero, sizeof(uint32_t), cudaMemcpyDeviceToHost));
		if (z == 0 || z == (uint32_t)m)
			continue; // constant bit: the partition is the identity
		k_wm_sort_scatter<<<gm, 256>>>(keys[cur], pos[cur], m, scratch, z, keys[cur ^ 1], pos[cur ^ 1]);
		MG_LAUNCH_CHECK(ctx);
		cur ^= 1;
	}
	uint32_t *rank_of = keys[cur ^ 1]; // reuse: rank key of every compacted position, in bin order
	if (m) {
		k_wm_assign_ranks<<<gm, 256>>>(keys[cur], pos[cur], m, sorted, rank_of);
		MG_LAUNCH_CHECK(ctx);
	}
	// (2) the matrix: level lv partitions by bit (nlevels - 1 - lv) of the rank
	uint32_t *wk[2] = {rank_of, keys[cur]};
	std::vector<uint32_t> h_zeros(MG_WM_LEVELS, 0);
	for (int lv = 0; lv < nlevels; ++lv) {
		uint2 *level = wm + (size_t)lv * dc.wm_blocks;
		const uint32_t *src = wk[lv & 1];
		if (lv >= 4 && (lv & 3) == 0) { // checkpoint: the rank keys in the order entering this level
			uint32_t *ck = nullptr;
			if (mg_group_alloc(ctx, t, MG_G_WM, m + 16, &ck, &arena))
				return 1;
			MG_CUDA(ctx, cudaMemset(ck + m, 0xff, sizeof(uint32_t) * 16));
			k_wm_ckpt<<<gm, 256>>>(src, m, sorted, ck);
			MG_LAUNCH_CHECK(ctx);
			dc.wm_ckpt[(lv >> 2) - 1] = ck;
		}
		k_wm_bits<<<gb, 256>>>(src, m, nlevels - 1 - lv, level, dc.wm_blocks);
		MG_LAUNCH_CHECK(ctx);
		k_wm_scan<<<1, 1024>>>(level, dc.wm_blocks, m, zeros + lv);
		MG_LAUNCH_CHECK(ctx);
		MG_CUDA(ctx, cudaMemcpy(&h_zeros[lv], zeros + lv, sizeof(uint32_t), cudaMemcpyDeviceToHost));
		k_wm_scatter<<<gm, 256>>>(src, m, level, h_zeros[lv], wk[(lv + 1) & 1]);
		MG_LAUNCH_CHECK(ctx);
	}
	MG_CUDA(ctx, cudaDeviceSynchronize());
	return 0;
}

static const int64_t MG_GROUP_EST_BYTES_PER_BIN[MG_G_COUNT] = {8, 13, 2, 8, 32}; // what a group takes, rounded up (budget checks before building)
static const char *const MG_GROUP_NAME[MG_G_COUNT] = {"stddev prefix", "min / max tables", "lse pyramid", "block wavelet matrices", "chromosome wavelet matrix"};

static int mg_build_group_chrom(mg_ctx *ctx, mg_dense *t, int g, DenseChrom &dc)
{
	MgArena arena;
	arena.chunk = (size_t)MG_GROUP_EST_BYTES_PER_BIN[g] * (size_t)dc.nbins + ((size_t)1 << 20); // one chunk per chromosome and group
	switch (g) {
	case MG_G_SD: return mg_build_sd(ctx, t, dc, arena);
	case MG_G_EXT: return mg_build_ext(ctx, t, dc, arena);
	case MG_G_LSE: return mg_build_lse(ctx, t, dc, arena);
	case MG_G_BW: return mg_build_bw(ctx, t, dc, arena);
	default: return mg_build_wm(ctx, t, dc, arena);
	}
}

static void mg_clear_group_pointers(DenseChrom &dc, int g)
{
	switch (g) {
	case MG_G_SD: dc.pref2 = nullptr; break;
	case MG_G_EXT:
		for (int j = 0; j < MG_PYR_MAX; ++j)
			dc.pmax[j] = dc.pmin[j] = nullptr;
		for (int j = 0; j < MG_ST_LEVELS; ++j)
			dc.smax[j] = dc.smin[j] = nullptr;
		break;
	case MG_G_LSE:
		for (int j = 0; j < MG_PYR_MAX; ++j)
			dc.plse[j] = nullptr;
		break;
	case MG_G_BW: dc.bw = nullptr; break;
	default:
		dc.wm = nullptr;
		dc.wm_zeros = nullptr;
		dc.wm_sorted = nullptr;
		for (int j = 0; j < 7; ++j)
			dc.wm_ckpt[j] = nullptr;
		dc.wm_levels = 0;
		dc.wm_blocks = 0;
		dc.wm_m = 0;
	}
}

static int64_t mg_dense_held(const mg_ctx *ctx)
{
	int64_t b = 0;
	for (const mg_dense *t : ctx->dense_tracks)
		b += t->bytes;
	return b;
}
static int64_t mg_dense_budget(const mg_ctx *ctx) { return ctx->hbm_budget > 0 ? ctx->hbm_budget : (int64_t)(0.85 * (double)ctx->hbm_bytes); }

// drops one built group of a track (every chromosome's part)
static int mg_drop_group(mg_ctx *ctx, mg_dense *t, int g)
{
	DenseGroup &gr = t->group[g];
	MG_CUDA(ctx, cudaDeviceSynchronize()); // nothing in flight reads what goes
	for (DenseChrom &dc : t->h_table)
		mg_clear_group_pointers(dc, g);
	MG_CUDA(ctx, cudaMemcpy(t->d_table, t->h_table.data(), sizeof(DenseChrom) * t->nchroms, cudaMemcpyHostToDevice));
	for (void *p : gr.allocs)
		cudaFree(p);
	gr.allocs.clear();
	t->bytes -= gr.bytes;
	gr.bytes = 0;
	gr.built = false;
	return 0;
}

// makes room for `need` more bytes: least-recently-used groups go first, never one of `keep_t`'s groups in `keep_mask`
static int mg_make_room(mg_ctx *ctx, int64_t need, const mg_dense *keep_t, uint32_t keep_mask)
{
	const int64_t budget = mg_dense_budget(ctx);
	while (mg_dense_held(ctx) + need > budget) {
		mg_dense *vt = nullptr;
		int vg = -1;
		for (mg_dense *t : ctx->dense_tracks)
			for (int g = 0; g < MG_G_COUNT; ++g) {
				if (!t->group[g].built || (t == keep_t && ((keep_mask >> g) & 1u)))
					continue;
				if (!vt || t->group[g].last_use < vt->group[vg].last_use) {
					vt = t;
					vg = g;
				}
			}
		if (!vt)
			return mg_fail(ctx, "dense tracks hold %lld MB and %lld MB more are needed, beyond the budget of %lld MB (option hbm_budget_mb) with nothing left to drop",
						   (long long)(mg_dense_held(ctx) >> 20), (long long)(need >> 20), (long long)(budget >> 20));
		if (mg_drop_group(ctx, vt, vg))
			return 1;
	}
	return 0;
}

// Every group of `mask` built for every staged chromosome of the track; a no-op (one mask test) when they are.
static int mg_dense_ensure(mg_ctx *ctx, mg_dense *t, uint32_t mask)
{
	const int64_t tick = ++ctx->use_tick;
	uint32_t missing = 0;
	for (int g = 0; g < MG_G_COUNT; ++g)
		if ((mask >> g) & 1u) {
			t->group[g].last_use = tick;
			if (!t->group[g].built)
				missing |= 1u << g;
		}
	if (!missing)
		return 0;
	MG_CUDA(ctx, cudaSetDevice(ctx->device));
	int64_t staged_bins = 0;
	for (const DenseChrom &dc : t->h_table)
		if (dc.vals)
			staged_bins += dc.nbins;
	for (int g = 0; g < MG_G_COUNT; ++g) {
		if (!((missing >> g) & 1u))
			continue;
		if (mg_make_room(ctx, MG_GROUP_EST_BYTES_PER_BIN[g] * staged_bins, t, mask))
			return 1;
		StagePhase ph(MG_GROUP_NAME[g]);
		for (DenseChrom &dc : t->h_table)
			if (dc.vals && mg_build_group_chrom(ctx, t, g, dc))
				return 1;
		t->bytes += t->group[g].bytes;
		t->group[g].built = true;
	}
	MG_CUDA(ctx, cudaDeviceSynchronize());
	MG_CUDA(ctx, cudaMemcpy(t->d_table, t->h_table.data(), sizeof(DenseChrom) * t->nchroms, cudaMemcpyHostToDevice));
	return 0;
}

extern "C" int mg_dense_stage(mg_ctx *ctx, mg_dense *t, int chromid, const float *h_bins, int64_t nbins)
{
	MG_REQUIRE(ctx, chromid >= 0 && chromid < t->nchroms, "mg_dense_stage: chromid %d out of range", chromid);
	MG_REQUIRE(ctx, nbins >= 0, "mg_dense_stage: negative bin count");
	MG_REQUIRE(ctx, t->h_table[chromid].vals == nullptr, "mg_dense_stage: chromosome %d already staged", chromid);
	// the reference validates the bin count against the chromosome size on every scan (src/TrackExpressionScanner.cpp:1160-1166)
	const int64_t expect = mg_div_up(ctx->chrom_sizes[chromid], (int64_t)t->bin_size);
	MG_REQUIRE(ctx, nbins == expect, "mg_dense_stage: chromosome %d holds %lld bins, expected %lld for bin size %u", chromid,
			   (long long)nbins, (long long)expect, t->bin_size);
	MG_CUDA(ctx, cudaSetDevice(ctx->device));
	if (mg_make_room(ctx, 6 * nbins + (1 << 20), t, ~0u))
		return 1;
	// the core: the bins and the per-group-of-8 prefix (6 bytes per bin); everything else waits for the function that needs it
	MgArena arena;
	arena.chunk = (size_t)6 * (size_t)nbins + ((size_t)1 << 20);
	float *vals = nullptr;
	Pref *pref = nullptr;
	const int64_t padded = (nbins + 15) / 16 * 16 + 16; // group-of-16 / float4-safe tail, NaN filled
	if (mg_alloc_tracked(ctx, t->allocs, t->bytes, padded, &vals, &arena))
		return 1;
	MgScratch tmp(nullptr);
	MG_CUDA(ctx, tmp.alloc(&pref, sizeof(Pref) * (size_t)(nbins + 1))); // per-bin prefix: staging scratch
	MG_CUDA(ctx, cudaMemset(vals, 0xff, sizeof(float) * padded)); // 0xffffffff is a NaN
	if (nbins)
		MG_CUDA(ctx, cudaMemcpy(vals, h_bins, sizeof(float) * nbins, cudaMemcpyHostToDevice));
	double tot[2] = {0, 0};
	if (mg_dense_prefix_pass(ctx, vals, nbins, pref, nullptr, tot)) // also turns +-inf into NaN in place
		return 1;
	DenseChrom dc;
	memset(&dc, 0, sizeof(dc));
	dc.vals = vals;
	{
		const int64_t nent = mg_div_up(nbins, 8) + 1;
		Pref *p8 = nullptr;
		if (mg_alloc_tracked(ctx, t->allocs, t->bytes, nent, &p8, &arena))
			return 1;
		k_p8_build<<<(unsigned)mg_div_up(nent, 256), 256>>>(pref, vals, nbins, p8, nent);
		MG_LAUNCH_CHECK(ctx);
		dc.p8 = p8;
	}
	dc.nbins = nbins;
	dc.ngroups = mg_div_up(nbins > 0 ? nbins : 1, 8);
	dc.abs_sum = tot[0];
	dc.sq_sum = tot[1];
	// groups this track already holds for its other chromosomes
	for (int g = 0; g < MG_G_COUNT; ++g)
		if (t->group[g].built) {
			const int64_t b0 = t->group[g].bytes;
			if (mg_build_group_chrom(ctx, t, g, dc))
				return 1;
			t->bytes += t->group[g].bytes - b0;
		}
	MG_CUDA(ctx, cudaDeviceSynchronize());
	t->h_table[chromid] = dc;
	MG_CUDA(ctx, cudaMemcpy(t->d_table + chromid, &dc, sizeof(DenseChrom), cudaMemcpyHostToDevice));
	return 0;
}

extern "C" void mg_dense_free(mg_ctx *ctx, mg_dense *t)
{
	if (!t)
		return;
	if (ctx)
		ctx->dense_tracks.erase(std::remove(ctx->dense_tracks.begin(), ctx->dense_tracks.end(), t), ctx->dense_tracks.end());
	for (int g = 0; g < MG_G_COUNT; ++g)
		for (void *p : t->group[g].allocs)
			cudaFree(p);
	for (void *p : t->allocs)
		cudaFree(p);
	delete t;
}

extern "C" int64_t mg_dense_bytes(const mg_dense *t) { return t->bytes; }

// ---- rows ------------------------------------------------------------------------------------------------
static int mg_check_rows(mg_ctx *ctx, int64_t n, const int32_t *h_chrom, const int64_t *h_start, const int64_t *h_end, const char *who);
extern "C" int mg_rows_upload(mg_ctx *ctx, int64_t n, const int32_t *h_chrom, const int64_t *h_start, const int64_t *h_end, void *stream,
							  mg_rows **out)
{
	MG_REQUIRE(ctx, n >= 0, "mg_rows_upload: negative row count");
	MG_REQUIRE(ctx, n == 0 || (h_chrom && h_start && h_end), "mg_rows_upload: NULL row arrays");
	if (mg_check_rows(ctx, n, h_chrom, h_start, h_end, "mg_rows_upload"))
		return 1;
	mg_rows *r = new (std::nothrow) mg_rows();
	MG_REQUIRE(ctx, r, "out of memory");
	const size_t n8 = (size_t)(n > 0 ? n : 1);
	const size_t off_end = n8 * 8, off_chrom = n8 * 16, total = n8 * 20;
	cudaError_t e = cudaMalloc(&r->d_block, total);
	if (e != cudaSuccess) {
		delete r;
		return mg_fail(ctx, "mg_rows_upload: %s", cudaGetErrorString(e));
	}
	char *base = (char *)r->d_block;
	r->owns = true;
	r->src.kind = 0;
	r->src.n = n;
	r->src.start = (const int64_t *)base;
	r->src.end = (const int64_t *)(base + off_end);
	r->src.chrom = (const int32_t *)(base + off_chrom);
	r->src.scope_idx = nullptr;
	cudaStream_t st = mg_stream(stream);
	if (n) {
		if ((e = cudaMemcpyAsync(base, h_start, (size_t)n * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess ||
			(e = cudaMemcpyAsync(base + off_end, h_end, (size_t)n * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess ||
			(e = cudaMemcpyAsync(base + off_chrom, h_chrom, (size_t)n * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) {
			mg_rows_free(ctx, r);
			return mg_fail(ctx, "mg_rows_upload: %s", cudaGetErrorString(e));
		}
	}
	*out = r;
	return 0;
}

extern "C" int mg_rows_wrap(mg_ctx *ctx, int64_t n, const int32_t *d_chrom, const int64_t *d_start, const int64_t *d_end, mg_rows **out)
{
	mg_rows *r = new (std::nothrow) mg_rows();
	MG_REQUIRE(ctx, r, "out of memory");
	r->src.kind = 0;
	r->src.n = n;
	r->src.chrom = d_chrom;
	r->src.start = d_start;
	r->src.end = d_end;
	*out = r;
	return 0;
}

extern "C" int mg_rows_fixedbin(mg_ctx *ctx, int64_t binsize, int64_t nscope, const int32_t *h_chrom, const int64_t *h_start,
								const int64_t *h_end, mg_rows **out)
{
	MG_REQUIRE(ctx, binsize > 0, "Bin size of a fixed bin iterator (%lld) must be positive", (long long)binsize);
	MG_REQUIRE(ctx, nscope >= 0, "mg_rows_fixedbin: negative scope size");
	std::vector<int64_t> row0((size_t)nscope + 1, 0);
	for (int64_t k = 0; k < nscope; ++k) {
		MG_REQUIRE(ctx, h_chrom[k] >= 0 && h_chrom[k] < ctx->nchroms && h_start[k] >= 0 && h_start[k] < h_end[k] &&
							h_end[k] <= ctx->chrom_sizes[h_chrom[k]],
				   "mg_rows_fixedbin: invalid scope interval %lld", (long long)k);
		MG_REQUIRE(ctx, k == 0 || h_chrom[k] > h_chrom[k - 1] || (h_chrom[k] == h_chrom[k - 1] && h_start[k] >= h_start[k - 1]),
				   "mg_rows_fixedbin: scope is not sorted at %lld", (long long)k);
		row0[k + 1] = row0[k] + (mg_div_up(h_end[k], binsize) - h_start[k] / binsize);
	}
	mg_rows *r = new (std::nothrow) mg_rows();
	MG_REQUIRE(ctx, r, "out of memory");
	const size_t ns = (size_t)(nscope > 0 ? nscope : 1);
	const size_t total = ns * 8 * 2 + (ns + 1) * 8 * 2 + ns * 4;
	cudaError_t e = cudaMalloc(&r->d_block, total);
	if (e != cudaSuccess) {
		delete r;
		return mg_fail(ctx, "mg_rows_fixedbin: %s", cudaGetErrorString(e));
	}
	char *base = (char *)r->d_block;
	r->owns = true;
	r->src.kind = 1;
	r->src.n = row0[nscope];
	r->src.binsize = binsize;
	r->src.nscope = nscope;
	r->src.sstart = (const int64_t *)base;
	r->src.send = (const int64_t *)(base + ns * 8);
	r->src.srow0 = (const int64_t *)(base + ns * 16);
	r->src.sbin0 = (const int64_t *)(base + ns * 16 + (ns + 1) * 8);
	r->src.schrom = (const int32_t *)(base + ns * 16 + (ns + 1) * 16);
	r->h_schrom.assign(h_chrom, h_chrom + nscope);
	r->h_sstart.assign(h_start, h_start + nscope);
	r->h_send.assign(h_end, h_end + nscope);
	r->h_srow0 = row0;
	for (int64_t k = 0; k <= nscope && k <= MG_SCOPE_INLINE; ++k)
		r->src.srow0_inl[k] = row0[k];
	if (nscope) {
		if ((e = cudaMemcpy((void *)r->src.sstart, h_start, (size_t)nscope * 8, cudaMemcpyHostToDevice)) != cudaSuccess ||
			(e = cudaMemcpy((void *)r->src.send, h_end, (size_t)nscope * 8, cudaMemcpyHostToDevice)) != cudaSuccess ||
			(e = cudaMemcpy((void *)r->src.schrom, h_chrom, (size_t)nscope * 4, cudaMemcpyHostToDevice)) != cudaSuccess) {
			mg_rows_free(ctx, r);
			return mg_fail(ctx, "mg_rows_fixedbin: %s", cudaGetErrorString(e));
		}
	}
	std::vector<int64_t> bin0((size_t)nscope + 1, 0);
	for (int64_t k = 0; k < nscope; ++k)
		bin0[k] = h_start[k] / binsize; // first bin of the scope interval: no 64-bit division on the device
	if ((e = cudaMemcpy((void *)r->src.sbin0, bin0.data(), ((size_t)nscope + 1) * 8, cudaMemcpyHostToDevice)) != cudaSuccess ||
		(e = cudaMemcpy((void *)r->src.srow0, row0.data(), ((size_t)nscope + 1) * 8, cudaMemcpyHostToDevice)) != cudaSuccess) {
		mg_rows_free(ctx, r);
		return mg_fail(ctx, "mg_rows_fixedbin: %s", cudaGetErrorString(e));
	}
	*out = r;
	return 0;
}

extern "C" int64_t mg_rows_count(const mg_rows *r) { return r->src.n; }

__global__ void k_rows_coords(const RowSrc rs, int64_t first, int64_t count, int32_t *chrom, int64_t *start, int64_t *end, int64_t *scope)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count)
		return;
	int c;
	int64_t s, e, k;
	mg_get_row(rs, first + i, c, s, e, k);
	if (chrom) chrom[i] = c;
	if (start) start[i] = s;
	if (end) end[i] = e;
	if (scope) scope[i] = k;
}

extern "C" int mg_rows_coords(mg_ctx *ctx, const mg_rows *r, int64_t first, int64_t count, int32_t *d_chrom, int64_t *d_start,
							  int64_t *d_end, int64_t *d_scope_idx, void *stream)
{
	MG_REQUIRE(ctx, first >= 0 && count >= 0 && first + count <= r->src.n, "mg_rows_coords: row range out of bounds");
	if (!count)
		return 0;
	k_rows_coords<<<(unsigned)mg_div_up(count, 256), 256, 0, mg_stream(stream)>>>(r->src, first, count, d_chrom, d_start, d_end, d_scope_idx);
	MG_LAUNCH_CHECK(ctx);
	return 0;
}

extern "C" void mg_rows_free(mg_ctx *ctx, mg_rows *r)
{
	if (!r)
		return;
	if (r->owns && r->d_block)
		cudaFree(r->d_block);
	delete r;
}

// ---- dense window functions --------------------------------------------------------------------------------
static int mg_fill_dense_query(mg_ctx *ctx, const mg_dense *t, const mg_rows *rows, int64_t first, int64_t count, int64_t sshift,
							   int64_t eshift, int nfuncs, const int *funcs, const double *params, double *const *d_out, DenseQuery &q,
							   bool &need_sweep)
{
	MG_REQUIRE(ctx, first >= 0 && count >= 0 && first + count <= rows->src.n, "row range [%lld, +%lld) out of bounds (%lld rows)",
			   (long long)first, (long long)count, (long long)rows->src.n);
	MG_REQUIRE(ctx, nfuncs > 0, "no functions requested");
	memset(&q, 0, sizeof(q));
	q.rows = rows->src;
	q.first = first;
	q.count = count;
	q.sshift = sshift;
	q.eshift = eshift;
	q.table = t->d_table;
	q.chrom_sizes = ctx->d_chrom_sizes;
	q.bin_size = t->bin_size;
	q.bin_magic = t->bin_size > 1 ? ~0ull / t->bin_size + 1 : 0;
	if (t->bin_size > 1) { // 2^(s-1) < d <= 2^s, m = floor(2^(31+s) / d) + 1 (< 2^32): floor(x / d) == (x m) >> (31 + s) for every x < 2^31
		uint32_t sft = 1;
		while ((1ull << sft) < t->bin_size)
			++sft;
		q.bin_magic32 = (uint32_t)(((1ull << (31 + sft)) / t->bin_size) + 1ull);
		q.bin_shift32 = sft - 1;
	}
	q.use_bw = ctx->block_wm ? 1 : 0;
	q.stage = ctx->stage_records ? 1 : 0;
	q.prefetch = ctx->prefetch ? 1 : 0;
	need_sweep = false;
	{ // the index groups these functions read, built now if this is their first use on this track (a mask test otherwise)
		uint32_t groups = 0;
		for (int k = 0; k < nfuncs; ++k) {
			const int f = funcs[k] & ~MG_FUNC_GENERIC;
			if (f == MG_MIN || f == MG_MAX || f == MG_MIN_POS || f == MG_MAX_POS)
				groups |= 1u << MG_G_EXT;
			else if (f == MG_STDDEV)
				groups |= 1u << MG_G_SD;
			else if (f == MG_LSE)
				groups |= 1u << MG_G_LSE;
			else if (f == MG_QUANTILE) {
				if (ctx->block_wm)
					groups |= 1u << MG_G_BW;
				// windows of more than MG_BW_S bins walk the chromosome-wide matrix. Only the fixed-bin iterator bounds its rows'
				// length without looking at them: iterator bin + shifts
				const bool bounded = ctx->block_wm && rows->src.kind == 1 &&
									 (rows->src.binsize + (eshift > 0 ? eshift : 0) - (sshift < 0 ? sshift : 0)) / (int64_t)t->bin_size + 2 <= MG_BW_S;
				if (!bounded)
					groups |= 1u << MG_G_WM;
			}
		}
		if (groups && mg_dense_ensure(ctx, const_cast<mg_dense *>(t), groups))
			return 1;
	}
	if (ctx->nchroms <= MG_INL_CHROMS && ctx->inline_table) { // after mg_dense_ensure: the group pointers are the current ones
		bool fits = true;
		for (int c = 0; c < ctx->nchroms; ++c)
			fits = fits && ctx->chrom_sizes[c] < (int64_t)0x7fffffff - (int64_t)t->bin_size && t->h_table[c].nbins < (int64_t)0x7fffffff;
		if (fits) {
			q.inl_n = ctx->nchroms;
			for (int c = 0; c < ctx->nchroms; ++c) {
				q.inl_csize[c] = (uint32_t)ctx->chrom_sizes[c];
				q.inl_nbins[c] = (uint32_t)t->h_table[c].nbins;
				q.inl_vals[c] = t->h_table[c].vals;
				q.inl_p8[c] = t->h_table[c].p8;
				q.inl_bw[c] = t->h_table[c].bw;
				q.inl_smax[c] = t->h_table[c].smax[0];
				q.inl_smin[c] = t->h_table[c].smin[0];
				q.inl_stpitch[c] = (uint32_t)t->h_table[c].st_pitch;
			}
		}
	}
	for (int k = 0; k < nfuncs; ++k) {
		const int f = funcs[k] & ~MG_FUNC_GENERIC;
		MG_REQUIRE(ctx, f >= 0 && f < MG_NUM_FUNCS, "unknown function id %d", f);
		if (funcs[k] & MG_FUNC_GENERIC) {
			MG_REQUIRE(ctx, f == MG_SIZE || f == MG_EXISTS || f == MG_LSE, "MG_FUNC_GENERIC applies to size, exists and lse (function id %d)", f);
			q.generic_mask |= 1u << f;
		}
		MG_REQUIRE(ctx, d_out[k], "output column %d is NULL", k);
		if (f == MG_QUANTILE) {
			MG_REQUIRE(ctx, q.nq < MG_MAX_Q, "at most %d quantile columns per call", MG_MAX_Q);
			MG_REQUIRE(ctx, params, "quantile needs a percentile parameter");
			q.q_pct[q.nq] = params[k] < 0. ? 0. : (params[k] > 1. ? 1. : params[k]); // src/StreamPercentiler.h:193-194
			q.q_out[q.nq] = d_out[k];
			q.nq++;
			need_sweep = true;
		} else {
			MG_REQUIRE(ctx, !q.out[f], "function %d requested twice in one call", f);
			q.out[f] = d_out[k];
			q.out_mask |= 1u << f;
			if (f == MG_MIN || f == MG_MAX || f == MG_STDDEV)
				need_sweep = true;
			if (f >= MG_FIRST_POS && f <= MG_MAX_POS && params && params[k] != 0.)
				q.pos_rel |= 1u << f;
		}
	}
	return 0;
}

static bool mg_any_out(double *const *out, int lo, int hi)
{
	for (int f = lo; f <= hi; ++f)
		if (out[f])
			return true;
	return false;
}

static int mg_launch_dense(mg_ctx *ctx, const DenseQuery &q, bool need_sweep, cudaStream_t st)
{
	if (!q.count)
		return 0;
	const bool want_sum = q.out[MG_AVG] || q.out[MG_SUM] || q.out[MG_NEAREST] || q.out[MG_SIZE] || q.out[MG_EXISTS];
	const bool want_mm = q.out[MG_MIN] || q.out[MG_MAX];
	if (mg_any_out(q.out, MG_LSE, MG_MAX_POS)) {
		// the order-dependent functions: first / last and min.pos / max.pos by a thread per row (early exit / sparse-table descent),
		// lse by a warp per row; force_sweep keeps everything on the warp-per-row kernel
		const int64_t blocks_needed = mg_div_up(q.count, MG_EXT_WARPS);
		const int64_t cap = (int64_t)ctx->sm_count * 64;
		const unsigned wgrid = (unsigned)(blocks_needed < cap ? blocks_needed : cap), tgrid = (unsigned)mg_div_up(q.count, 256);
		if (ctx->force_sweep) {
			if (q.out[MG_LSE] || q.out[MG_MIN_POS] || q.out[MG_MAX_POS])
				k_dense_ext<true><<<wgrid, MG_EXT_WARPS * 32, 0, st>>>(q);
			else
				k_dense_ext<false><<<wgrid, MG_EXT_WARPS * 32, 0, st>>>(q);
			MG_LAUNCH_CHECK(ctx);
		} else {
			if (mg_any_out(q.out, MG_FIRST, MG_LAST_POS)) {
				k_dense_firstlast<<<tgrid, 256, 0, st>>>(q);
				MG_LAUNCH_CHECK(ctx);
			}
			if (q.out[MG_MIN_POS] || q.out[MG_MAX_POS]) {
				k_dense_argext<<<tgrid, 256, 0, st>>>(q);
				MG_LAUNCH_CHECK(ctx);
			}
			if (q.out[MG_LSE]) {
				if ((q.generic_mask >> MG_LSE) & 1u)
					k_dense_lse_generic<<<tgrid, 256, 0, st>>>(q);
				else
					k_dense_lse<<<tgrid, 256, 0, st>>>(q);
				MG_LAUNCH_CHECK(ctx);
			}
		}
		if (!want_sum && !want_mm && !q.out[MG_STDDEV] && q.nq == 0)
			return 0;
	}
	const int ngroups = (want_sum ? 1 : 0) + (want_mm ? 1 : 0) + (q.nq > 0 ? 1 : 0);
	const bool tile_ok = ctx->tile_kernel && !ctx->force_sweep && !q.out[MG_STDDEV] && (q.nq == 0 || q.use_bw) &&
						 (ctx->tile_kernel == 3 ? (ngroups >= 2 || q.nq > 0) : (need_sweep || ctx->tile_kernel == 2 || ctx->tile_kernel >= 4));
	if (tile_ok) {
		// a thread block per tile of rows, everything from shared memory (dense_tile.cuh)
		const unsigned grid = (unsigned)mg_div_up(q.count, MG_TL_THREADS);
		const int key = (want_sum ? 1 : 0) | (q.out[MG_MAX] ? 2 : 0) | (q.out[MG_MIN] ? 4 : 0) | (q.nq > 0 ? 8 : 0);
#define MG_TILE_CASE(K)                                                                                                  \
	case K: {                                                                                                           \
		auto kern = k_dense_tile<((K) & 1) != 0, ((K) & 2) != 0, ((K) & 4) != 0, ((K) & 8) != 0>;                       \
		constexpr int smem_bytes = mg_tile_layout((((K) & 2) ? 1 : 0) + (((K) & 4) ? 1 : 0), ((K) & 8) != 0).total;                                \
		if (!(ctx->tile_attr & (1u << (K)))) {                                                                          \
			MG_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));            \
			ctx->tile_attr |= 1u << (K);                                                                                \
		}                                                                                                               \
		kern<<<grid, MG_TL_THREADS, smem_bytes, st>>>(q);                                                               \
		break;                                                                                                          \
	}
		if (ctx->tile_kernel == 5 || ctx->tile_kernel == 6) { // k_dense_thin's arithmetic in a persistent, multi-buffered thread block (k_dense_pipe)
#define MG_PIPE_CASE2(K, D)                                                                                              \
	{                                                                                                                   \
		auto kern = k_dense_pipe<((K) & 1) != 0, ((K) & 2) != 0, ((K) & 4) != 0, ((K) & 8) != 0, D>;                    \
		constexpr int smem_bytes = ((D) + 1) * MG_PP_STAGE(((K) & 8) != 0);                                             \
		const unsigned pgrid = (unsigned)std::min<int64_t>(grid, (int64_t)ctx->sm_count * MG_PP_BPS(D));               \
		if (!(ctx->pipe_attr & (1u << ((K) + 16 * ((D) - 1))))) {                                                       \
			MG_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));            \
			MG_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));                \
			ctx->pipe_attr |= 1u << ((K) + 16 * ((D) - 1));                                                             \
			if (getenv("MG_DEBUG_OCC")) {                                                                               \
				int nb_ = 0;                                                                                            \
				cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb_, kern, MG_TL_THREADS, smem_bytes);                   \
				fprintf(stderr, "k_dense_pipe<%d,%d>: %d resident blocks per SM, %d bytes dynamic smem, grid %u\n", (K), (D), nb_, smem_bytes, pgrid); \
			}                                                                                                           \
		}                                                                                                               \
		kern<<<pgrid, MG_TL_THREADS, smem_bytes, st>>>(q);                                                              \
	}
#define MG_PIPE_CASE(K)                                                                                                  \
	case K:                                                                                                             \
		if (ctx->tile_kernel == 5)                                                                                      \
			MG_PIPE_CASE2(K, 1)                                                                                         \
		else                                                                                                            \
			MG_PIPE_CASE2(K, 2)                                                                                         \
		break;
			switch (key) {
				MG_PIPE_CASE(1) MG_PIPE_CASE(2) MG_PIPE_CASE(3) MG_PIPE_CASE(4) MG_PIPE_CASE(5) MG_PIPE_CASE(6) MG_PIPE_CASE(7) MG_PIPE_CASE(8)
				MG_PIPE_CASE(9) MG_PIPE_CASE(10) MG_PIPE_CASE(11) MG_PIPE_CASE(12) MG_PIPE_CASE(13) MG_PIPE_CASE(14) MG_PIPE_CASE(15)
			default: break;
			}
#undef MG_PIPE_CASE2
#undef MG_PIPE_CASE
		} else if (ctx->tile_kernel >= 3) { // staging only (k_dense_thin): sums / extrema from the chromosome's structures
#define MG_THIN_CASE(K)                                                                                                  \
	case K: {                                                                                                           \
		auto kern = k_dense_thin<((K) & 1) != 0, ((K) & 2) != 0, ((K) & 4) != 0, ((K) & 8) != 0>;                       \
		constexpr int smem_bytes = MG_TH_SZ_VALS + (((K) & 8) ? MG_TL_SLOTS * MG_BW_STAGE_BYTES : 0);                    \
		if (!(ctx->tile_attr & (1u << (16 + (K))))) {                                                                   \
			MG_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));            \
			ctx->tile_attr |= 1u << (16 + (K));                                                                         \
		}                                                                                                               \
		kern<<<grid, MG_TL_THREADS, smem_bytes, st>>>(q);                                                               \
		break;                                                                                                          \
	}
			switch (key) {
				MG_THIN_CASE(1) MG_THIN_CASE(2) MG_THIN_CASE(3) MG_THIN_CASE(4) MG_THIN_CASE(5) MG_THIN_CASE(6) MG_THIN_CASE(7) MG_THIN_CASE(8)
				MG_THIN_CASE(9) MG_THIN_CASE(10) MG_THIN_CASE(11) MG_THIN_CASE(12) MG_THIN_CASE(13) MG_THIN_CASE(14) MG_THIN_CASE(15)
			default: break;
			}
#undef MG_THIN_CASE
		} else
		switch (key) {
			MG_TILE_CASE(1) MG_TILE_CASE(2) MG_TILE_CASE(3) MG_TILE_CASE(4) MG_TILE_CASE(5) MG_TILE_CASE(6) MG_TILE_CASE(7) MG_TILE_CASE(8)
			MG_TILE_CASE(9) MG_TILE_CASE(10) MG_TILE_CASE(11) MG_TILE_CASE(12) MG_TILE_CASE(13) MG_TILE_CASE(14) MG_TILE_CASE(15)
		default: break;
		}
#undef MG_TILE_CASE
	} else if (!need_sweep) {
		k_dense_prefix<<<(unsigned)mg_div_up(q.count, 256), 256, 0, st>>>(q);
	} else if (!ctx->force_sweep) {
		// everything else comes from the index structures, one thread per row
		const unsigned grid = (unsigned)mg_div_up(q.count, MG_RQ_THREADS);
		const bool want_sd = q.out[MG_STDDEV] != nullptr;
		const int key = (want_sum ? 1 : 0) | (want_mm ? 2 : 0) | (q.nq > 0 ? 4 : 0) | (want_sd ? 8 : 0);
#define MG_RQ_CASE(K)                                                                                                    \
	case K:                                                                                                             \
		k_dense_rowquery<((K) & 1) != 0, ((K) & 2) != 0, ((K) & 4) != 0, ((K) & 8) != 0><<<grid, MG_RQ_THREADS, 0, st>>>(q); \
		break;
		switch (key) {
			MG_RQ_CASE(1) MG_RQ_CASE(2) MG_RQ_CASE(3) MG_RQ_CASE(4) MG_RQ_CASE(5) MG_RQ_CASE(6) MG_RQ_CASE(7) MG_RQ_CASE(8) MG_RQ_CASE(9)
			MG_RQ_CASE(10) MG_RQ_CASE(11) MG_RQ_CASE(12) MG_RQ_CASE(13) MG_RQ_CASE(14) MG_RQ_CASE(15)
		default: break;
		}
#undef MG_RQ_CASE
	} else {
		const int64_t blocks_needed = mg_div_up(q.count, MG_SW_WARPS);
		const int64_t cap = (int64_t)ctx->sm_count * 32; // grid-stride beyond that
		const unsigned grid = (unsigned)(blocks_needed < cap ? blocks_needed : cap);
		const bool sum = q.out[MG_AVG] || q.out[MG_SUM] || q.out[MG_NEAREST], mm = q.out[MG_MIN] || q.out[MG_MAX], sd = q.out[MG_STDDEV] != nullptr,
				   qq = q.nq > 0;
		const int key = (sum ? 1 : 0) | (mm ? 2 : 0) | (sd ? 4 : 0) | (qq ? 8 : 0);
#define MG_SWEEP_CASE(K)                                                                                                \
	case K:                                                                                                             \
		k_dense_sweep<((K) & 1) != 0, ((K) & 2) != 0, ((K) & 4) != 0, ((K) & 8) != 0><<<grid, MG_SW_WARPS * 32, 0, st>>>(q);   \
		break;
		switch (key) {
			MG_SWEEP_CASE(0) MG_SWEEP_CASE(1) MG_SWEEP_CASE(2) MG_SWEEP_CASE(3) MG_SWEEP_CASE(4) MG_SWEEP_CASE(5) MG_SWEEP_CASE(6) MG_SWEEP_CASE(7)
			MG_SWEEP_CASE(8) MG_SWEEP_CASE(9) MG_SWEEP_CASE(10) MG_SWEEP_CASE(11) MG_SWEEP_CASE(12) MG_SWEEP_CASE(13) MG_SWEEP_CASE(14) MG_SWEEP_CASE(15)
		}
#undef MG_SWEEP_CASE
	}
	MG_LAUNCH_CHECK(ctx);
	if (q.generic_mask & ((1u << MG_SIZE) | (1u << MG_EXISTS))) { // the rows whose window lies inside one bin, as the generic path counts them
		k_dense_generic_counts<<<(unsigned)mg_div_up(q.count, 256), 256, 0, st>>>(q);
		MG_LAUNCH_CHECK(ctx);
	}
	return 0;
}

extern "C" int mg_dense_window(mg_ctx *ctx, const mg_dense *t, const mg_rows *rows, int64_t first, int64_t count, int64_t sshift,
							   int64_t eshift, int nfuncs, const int *funcs, const double *params, double *const *d_out, void *stream)
{
	DenseQuery q;
	bool need_sweep;
	if (mg_fill_dense_query(ctx, t, rows, first, count, sshift, eshift, nfuncs, funcs, params, d_out, q, need_sweep))
		return 1;
	return mg_launch_dense(ctx, q, need_sweep, mg_stream(stream));
}

// ---- host-buffer convenience (the end-to-end path) -----------------------------------------------------------
static int mg_pipe_prepare(mg_ctx *ctx, int64_t bytes_per_slot)
{
	MG_CUDA(ctx, cudaSetDevice(ctx->device));
	if (!ctx->d_first_bad) {
		MG_CUDA(ctx, cudaMalloc(&ctx->d_first_bad, 8));
		MG_CUDA(ctx, cudaMemset(ctx->d_first_bad, 0xff, 8)); // "no bad row"
	}
	for (int i = 0; i < ctx->pipe_slots; ++i) {
		if (!ctx->pipe_stream[i])
			MG_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->pipe_stream[i], cudaStreamNonBlocking));
		if (ctx->pipe_dev_bytes[i] < bytes_per_slot) {
			if (ctx->pipe_dev[i])
				cudaFree(ctx->pipe_dev[i]);
			ctx->pipe_dev[i] = nullptr;
			ctx->pipe_dev_bytes[i] = 0;
			MG_CUDA(ctx, cudaMalloc(&ctx->pipe_dev[i], (size_t)bytes_per_slot));
			ctx->pipe_dev_bytes[i] = bytes_per_slot;
		}
	}
	return 0;
}

// Rows in host memory -> result columns in host memory: chunks of pipe_chunk_rows rows travel through three device slots, each
// on its own stream (upload, the launch the caller supplies, download), so the copies of one chunk overlap the kernel of
// another in both directions. launch(rows, m, d_cols, stream) enqueues the computation of one chunk of m rows.
// chrom ids index the per-chromosome tables on the device: an id out of range there is an out-of-bounds read and a sticky error for
// the whole context, so explicit rows are checked where they are still host arrays
static int mg_check_rows(mg_ctx *ctx, int64_t n, const int32_t *h_chrom, const int64_t *, const int64_t *, const char *who)
{ // coordinates need no check: the window transform clamps them to the chromosome
	// one branch-free pass (the largest id as unsigned: negative ids become huge), the offending row is looked for only on failure
	uint32_t worst = 0;
	for (int64_t i = 0; i < n; ++i) {
		const uint32_t c = (uint32_t)h_chrom[i];
		worst = c > worst ? c : worst;
	}
	if (worst < (uint32_t)ctx->nchroms)
		return 0;
	for (int64_t i = 0; i < n; ++i) {
		const int32_t c = h_chrom[i];
		if (c < 0 || c >= ctx->nchroms)
			return mg_fail(ctx, "%s: row %lld names chromosome id %d (the context holds %d)", who, (long long)i, (int)c, ctx->nchroms);
	}
	return 0;
}

// Chromosome ids of a chunk already on the device: an id out of range is replaced by 0 -- the kernels index per-chromosome tables with
// it -- and the first offending row is recorded; the host looks at the record once, after the drain. (A pass over the ids on the
// host, even a branch-free one per chunk, cost more than a millisecond of the 5 ms a 10 M-row batch takes end to end.)
// served: when not null, served[id] != 0 marks the chromosomes this device holds (a member of a group is handed the rows of its own
// chromosomes by a routing that trusts the rows' order; a row of somebody else's chromosome is recorded the same way and the host re-routes).
__global__ void __launch_bounds__(256) k_rows_sanitize(int32_t *chrom, int64_t m, int nchroms, int64_t base_row, unsigned long long *first_bad,
													 const uint8_t *__restrict__ served)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m)
		return;
	const uint32_t c = (uint32_t)chrom[i];
	if (c >= (uint32_t)nchroms || (served && !served[c])) {
		atomicMin(first_bad, (unsigned long long)(base_row + i));
		chrom[i] = c < (uint32_t)nchroms ? (int32_t)c : 0; // a valid id of another member: tables exist (empty there), nothing is out of bounds
	}
}

static int64_t mg_pipe_slot_bytes(const mg_ctx *ctx, int ncols) { return ctx->pipe_chunk_rows * (20 + 8 * (int64_t)ncols); }

// Enqueues the chunks of one batch on the three slots and returns without waiting (mg_h_pipeline and the group calls drain the
// streams). The slots are taken in turn ACROSS calls (ctx->pipe_next), so consecutive batches on one device keep all three streams
// busy. A batch shorter than eight full chunks is cut into eight (not below 2^16 rows a chunk): with two or three chunks there is
// no steady state in which one chunk's upload overlaps another's download -- the 1.25 M rows a rank holds of the 10 M at 8 GPUs
// were 2.4 chunks.
template <typename Launch>
static int mg_h_pipeline_run(mg_ctx *ctx, int64_t n, const int32_t *h_chrom, const int64_t *h_start, const int64_t *h_end, int ncols,
							 double *const *h_out, Launch launch, int64_t row_base = 0)
{
	const int64_t CHUNK = ctx->pipe_chunk_rows; // rows a slot can hold
	int64_t step = ((n + 7) / 8 + 1023) & ~(int64_t)1023;
	step = step < 65536 ? 65536 : step;
	step = step > CHUNK ? CHUNK : step;
	for (int64_t done = 0; done < n; done += step) {
		const int64_t m = n - done < step ? n - done : step;
		const int slot = (int)(ctx->pipe_next++ % ctx->pipe_slots);
		cudaStream_t st = ctx->pipe_stream[slot];
		char *base = (char *)ctx->pipe_dev[slot];
		int64_t *d_start = (int64_t *)base, *d_end = (int64_t *)(base + CHUNK * 8);
		int32_t *d_chrom = (int32_t *)(base + CHUNK * 16);
		double *d_cols[32];
		for (int k = 0; k < ncols; ++k)
			d_cols[k] = (double *)(base + CHUNK * 20 + (int64_t)k * CHUNK * 8);
		MG_CUDA(ctx, cudaMemcpyAsync(d_start, h_start + done, (size_t)m * 8, cudaMemcpyHostToDevice, st));
		MG_CUDA(ctx, cudaMemcpyAsync(d_end, h_end + done, (size_t)m * 8, cudaMemcpyHostToDevice, st));
		MG_CUDA(ctx, cudaMemcpyAsync(d_chrom, h_chrom + done, (size_t)m * 4, cudaMemcpyHostToDevice, st));
		k_rows_sanitize<<<(unsigned)mg_div_up(m, 256), 256, 0, st>>>(d_chrom, m, ctx->nchroms, row_base + done, ctx->d_first_bad, ctx->d_served);
		MG_LAUNCH_CHECK(ctx);
		mg_rows rows;
		rows.src.kind = 0;
		rows.src.n = m;
		rows.src.chrom = d_chrom;
		rows.src.start = d_start;
		rows.src.end = d_end;
		if (launch(rows, m, d_cols, st))
			return 1;
		for (int k = 0; k < ncols; ++k)
			MG_CUDA(ctx, cudaMemcpyAsync(h_out[k] + done, d_cols[k], (size_t)m * 8, cudaMemcpyDeviceToHost, st));
	}
	return 0;
}

// After the drain: did k_rows_sanitize meet a chromosome id out of range? (the record is reset either way)
static int mg_pipe_bad_row(mg_ctx *ctx, const int32_t *h_chrom, const char *who)
{
	unsigned long long bad = ~0ull;
	MG_CUDA(ctx, cudaMemcpy(&bad, ctx->d_first_bad, 8, cudaMemcpyDeviceToHost));
	if (bad == ~0ull)
		return 0;
	MG_CUDA(ctx, cudaMemset(ctx->d_first_bad, 0xff, 8));
	return mg_fail(ctx, "%s: row %llu names chromosome id %d (the context holds %d)", who, bad, (int)h_chrom[bad], ctx->nchroms);
}

// Rows in host memory -> result columns in host memory: chunks of pipe_chunk_rows rows travel through three device slots, each
// on its own stream (upload, the launch the caller supplies, download), so the copies of one chunk overlap the kernel of
// another in both directions. launch(rows, m, d_cols, stream) enqueues the computation of one chunk of m rows.
// Whatever happens on the way, the three streams are drained before the call returns: earlier chunks' copies into h_out and out of
// the row arrays may still be in flight when a later step fails, and the caller is free to release its buffers on error.
template <typename Launch>
static int mg_h_pipeline(mg_ctx *ctx, int64_t n, const int32_t *h_chrom, const int64_t *h_start, const int64_t *h_end, int ncols,
						 double *const *h_out, Launch launch)
{
	MG_REQUIRE(ctx, n >= 0 && ncols > 0 && ncols <= 32, "host pipeline: bad arguments");
	MG_REQUIRE(ctx, n == 0 || (h_chrom && h_start && h_end && h_out), "host pipeline: NULL row or output arrays");
	if (mg_pipe_prepare(ctx, mg_pipe_slot_bytes(ctx, ncols)))
		return 1;
	int rc = mg_h_pipeline_run(ctx, n, h_chrom, h_start, h_end, ncols, h_out, launch);
	const std::string first_err = rc ? ctx->err : std::string();
	for (int i = 0; i < MG_PIPE_MAX_SLOTS; ++i) {
		if (!ctx->pipe_stream[i])
			continue;
		const cudaError_t e = cudaStreamSynchronize(ctx->pipe_stream[i]);
		if (e != cudaSuccess && !rc)
			rc = mg_fail(ctx, "host pipeline: %s", cudaGetErrorString(e));
	}
	if (!first_err.empty())
		ctx->err = first_err; // the failure that started it, not what the drain reported
	if (!rc)
		rc = mg_pipe_bad_row(ctx, h_chrom, "host pipeline");
	return rc;
}

extern "C" int mg_h_dense_window(mg_ctx *ctx, const mg_dense *t, int64_t n, const int32_t *h_chrom, const int64_t *h_start,
								 const int64_t *h_end, int64_t sshift, int64_t eshift, int nfuncs, const int *funcs, const double *params,
								 double *const *h_out)
{
	MG_REQUIRE(ctx, n >= 0 && nfuncs > 0 && nfuncs <= 32, "mg_h_dense_window: bad arguments");
	if (ctx->zero_copy && n > 0) {
		// Every buffer pinned (cudaMallocHost / cudaHostRegister / torch pin_memory)? Then the kernel reads the rows and
		// writes the result columns directly over PCIe: reads and posted writes overlap in both directions and there is no
		// staging copy at all. Otherwise fall through to the chunked copy pipeline.
		const void *hp[3 + 32] = {h_chrom, h_start, h_end};
		for (int k = 0; k < nfuncs; ++k)
			hp[3 + k] = h_out[k];
		void *dp[3 + 32];
		bool pinned = true;
		for (int k = 0; k < 3 + nfuncs && pinned; ++k) {
			cudaPointerAttributes at;
			if (cudaPointerGetAttributes(&at, hp[k]) != cudaSuccess || at.type != cudaMemoryTypeHost || !at.devicePointer) {
				cudaGetLastError();
				pinned = false;
			} else
				dp[k] = at.devicePointer;
		}
		if (pinned) {
			MG_CUDA(ctx, cudaSetDevice(ctx->device));
			if (mg_pipe_prepare(ctx, 1))
				return 1;
			cudaStream_t st = ctx->pipe_stream[0];
			mg_rows rows;
			rows.src.kind = 0;
			rows.src.n = n;
			rows.src.chrom = (const int32_t *)dp[0];
			rows.src.start = (const int64_t *)dp[1];
			rows.src.end = (const int64_t *)dp[2];
			double *d_cols[32];
			for (int k = 0; k < nfuncs; ++k)
				d_cols[k] = (double *)dp[3 + k];
			DenseQuery q;
			bool need_sweep;
			if (mg_fill_dense_query(ctx, t, &rows, 0, n, sshift, eshift, nfuncs, funcs, params, d_cols, q, need_sweep) ||
				mg_launch_dense(ctx, q, need_sweep, st))
				return 1;
			MG_CUDA(ctx, cudaStreamSynchronize(st));
			return 0;
		}
	}
	return mg_h_pipeline(ctx, n, h_chrom, h_start, h_end, nfuncs, h_out, [&](const mg_rows &rows, int64_t m, double *const *d_cols, cudaStream_t st) {
		DenseQuery q;
		bool need_sweep;
		return mg_fill_dense_query(ctx, t, &rows, 0, m, sshift, eshift, nfuncs, funcs, params, d_cols, q, need_sweep) ||
			   mg_launch_dense(ctx, q, need_sweep, st);
	});
}

// ---- generic device exclusive scan of int64 (tile reduce -> one-block scan -> tile write) -------------------------
#define MG_SCAN_THREADS 256
#define MG_SCAN_PER_THREAD 8
#define MG_SCAN_TILE (MG_SCAN_THREADS * MG_SCAN_PER_THREAD)

__global__ void __launch_bounds__(MG_SCAN_THREADS) k_scan_tile_reduce(const long long *in, int64_t n, long long *tile_tot)
{
	__shared__ long long sc[33];
	const int64_t base = (int64_t)blockIdx.x * MG_SCAN_TILE + (int64_t)threadIdx.x * MG_SCAN_PER_THREAD;
	long long s = 0;
#pragma unroll
	for (int j = 0; j < MG_SCAN_PER_THREAD; ++j)
		if (base + j < n)
			s += in[base + j];
	long long tot;
	mg_block_excl_scan<long long>(s, sc, tot);
	if (threadIdx.x == 0)
		tile_tot[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024) k_scan_tiles(long long *tile_tot, int64_t ntiles, long long *total)
{
	__shared__ long long sc[33];
	const int64_t chunk = (ntiles + blockDim.x - 1) / blockDim.x;
	const int64_t lo = (int64_t)threadIdx.x * chunk, hi = min(lo + chunk, ntiles);
	long long s = 0;
	for (int64_t i = lo; i < hi; ++i)
		s += tile_tot[i];
	long long tot;
	long long s0 = mg_block_excl_scan<long long>(s, sc, tot);
	for (int64_t i = lo; i < hi; ++i) {
		const long long v = tile_tot[i];
		tile_tot[i] = s0;
		s0 += v;
	}
	if (threadIdx.x == 0)
		*total = tot;
}

// out[i] = sum in[0..i), out[n] = total (out has n + 1 entries; may alias nothing)
__global__ void __launch_bounds__(MG_SCAN_THREADS) k_scan_write(const long long *in, int64_t n, const long long *tile_tot, long long *out)
{
	__shared__ long long sc[33];
	const int64_t base = (int64_t)blockIdx.x * MG_SCAN_TILE + (int64_t)threadIdx.x * MG_SCAN_PER_THREAD;
	long long v[MG_SCAN_PER_THREAD], s = 0;
#pragma unroll
	for (int j = 0; j < MG_SCAN_PER_THREAD; ++j) {
		v[j] = base + j < n ? in[base + j] : 0;
		s += v[j];
	}
	long long tot;
	long long s0 = tile_tot[blockIdx.x] + mg_block_excl_scan<long long>(s, sc, tot);
#pragma unroll
	for (int j = 0; j < MG_SCAN_PER_THREAD; ++j) {
		if (base + j <= n)
			out[base + j] = s0;
		s0 += v[j];
	}
}

// d_out[0..n] = exclusive prefix of d_in[0..n); *h_total = d_out[n]. Synchronises the stream.
static int mg_excl_scan_i64(mg_ctx *ctx, const long long *d_in, int64_t n, long long *d_out, long long *h_total, cudaStream_t st)
{
	const int64_t ntiles = mg_div_up(n + 1, MG_SCAN_TILE);
	long long *tile_tot = nullptr, *d_total = nullptr;
	MG_CUDA(ctx, cudaMalloc(&tile_tot, sizeof(long long) * (ntiles + 1)));
	d_total = tile_tot + ntiles;
	k_scan_tile_reduce<<<(unsigned)ntiles, MG_SCAN_THREADS, 0, st>>>(d_in, n, tile_tot);
	MG_LAUNCH_CHECK(ctx);
	k_scan_tiles<<<1, 1024, 0, st>>>(tile_tot, ntiles, d_total);
	MG_LAUNCH_CHECK(ctx);
	k_scan_write<<<(unsigned)ntiles, MG_SCAN_THREADS, 0, st>>>(d_in, n, tile_tot, d_out);
	MG_LAUNCH_CHECK(ctx);
	MG_CUDA(ctx, cudaMemcpyAsync(h_total, d_total, sizeof(long long), cudaMemcpyDeviceToHost, st));
	MG_CUDA(ctx, cudaStreamSynchronize(st));
	cudaFree(tile_tot);
	return 0;
}

// ---- intervals iterator rows (src/TrackExpressionIntervals1DIterator.cpp:21-85) ------------------------------------
// For scope interval k the overlapping iterator intervals are the contiguous range [lo_k, hi_k): the iterator set is
// sorted and disjoint, so lo_k = first interval with (chrom, end) > (chrom_k, start_k) and hi_k = first with
// (chrom, start) >= (chrom_k, end_k).
struct IvIterArgs {
	int64_t niter, nscope;
	const int32_t *ichrom;
	const int64_t *istart, *iend;
	const int32_t *schrom;
	const int64_t *sstart, *send;
};

__global__ void __launch_bounds__(256) k_ivrows_range(const IvIterArgs a, long long *lo_out, long long *cnt_out)
{
	const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= a.nscope)
		return;
	const int32_t c = a.schrom[k];
	const int64_t s = a.sstart[k], e = a.send[k];
	int64_t lo = 0, hi = a.niter;
	while (lo < hi) {
		const int64_t mid = lo + ((hi - lo) >> 1);
		const int32_t mc = __ldg(a.ichrom + mid);
		if (mc < c || (mc == c && __ldg(a.iend + mid) <= s))
			lo = mid + 1;
		else
			hi = mid;
	}
	const int64_t first = lo;
	hi = a.niter;
	while (lo < hi) {
		const int64_t mid = lo + ((hi - lo) >> 1);
		const int32_t mc = __ldg(a.ichrom + mid);
		if (mc < c || (mc == c && __ldg(a.istart + mid) < e))
			lo = mid + 1;
		else
			hi = mid;
	}
	lo_out[k] = first;
	cnt_out[k] = lo - first;
}

__global__ void __launch_bounds__(256) k_ivrows_write(const IvIterArgs a, const long long *lo, const long long *off, int64_t nrows, int32_t *ochrom,
													   int64_t *ostart, int64_t *oend, int64_t *oscope)
{
	const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= nrows)
		return;
	int64_t klo = 0, khi = a.nscope; // last k with off[k] <= r
	while (khi - klo > 1) {
		const int64_t mid = (klo + khi) >> 1;
		if (__ldg(off + mid) <= r)
			klo = mid;
		else
			khi = mid;
	}
	const int64_t k = klo, j = lo[k] + (r - off[k]);
	const int64_t s = a.sstart[k], e = a.send[k], is = a.istart[j], ie = a.iend[j];
	ochrom[r] = a.schrom[k];
	ostart[r] = s > is ? s : is;
	oend[r] = e < ie ? e : ie;
	oscope[r] = k;
}

extern "C" int mg_rows_intervals(mg_ctx *ctx, int64_t niter, const int32_t *h_ichrom, const int64_t *h_istart, const int64_t *h_iend,
								 int64_t nscope, const int32_t *h_schrom, const int64_t *h_sstart, const int64_t *h_send, void *stream,
								 mg_rows **out)
{
	MG_REQUIRE(ctx, niter >= 0 && nscope >= 0, "mg_rows_intervals: negative sizes");
	cudaStream_t st = mg_stream(stream);
	MG_CUDA(ctx, cudaSetDevice(ctx->device));
	const size_t ni = (size_t)(niter > 0 ? niter : 1), ns = (size_t)(nscope > 0 ? nscope : 1);
	char *tmp = nullptr;
	const size_t tmp_bytes = ni * 20 + ns * 20 + ns * 8 * 2 + (ns + 1) * 8;
	MG_CUDA(ctx, cudaMalloc(&tmp, tmp_bytes));
	IvIterArgs a;
	a.niter = niter;
	a.nscope = nscope;
	char *p = tmp;
	a.istart = (int64_t *)p; p += ni * 8;
	a.iend = (int64_t *)p; p += ni * 8;
	a.sstart = (int64_t *)p; p += ns * 8;
	a.send = (int64_t *)p; p += ns * 8;
	long long *lo = (long long *)p; p += ns * 8;
	long long *cnt = (long long *)p; p += ns * 8;
	long long *off = (long long *)p; p += (ns + 1) * 8;
	a.ichrom = (int32_t *)p; p += ni * 4;
	a.schrom = (int32_t *)p;
	if (niter) {
		MG_CUDA(ctx, cudaMemcpyAsync((void *)a.istart, h_istart, niter * 8, cudaMemcpyHostToDevice, st));
		MG_CUDA(ctx, cudaMemcpyAsync((void *)a.iend, h_iend, niter * 8, cudaMemcpyHostToDevice, st));
		MG_CUDA(ctx, cudaMemcpyAsync((void *)a.ichrom, h_ichrom, niter * 4, cudaMemcpyHostToDevice, st));
	}
	if (nscope) {
		MG_CUDA(ctx, cudaMemcpyAsync((void *)a.sstart, h_sstart, nscope * 8, cudaMemcpyHostToDevice, st));
		MG_CUDA(ctx, cudaMemcpyAsync((void *)a.send, h_send, nscope * 8, cudaMemcpyHostToDevice, st));
		MG_CUDA(ctx, cudaMemcpyAsync((void *)a.schrom, h_schrom, nscope * 4, cudaMemcpyHostToDevice, st));
	}
	long long nrows = 0;
	if (nscope && niter) {
		k_ivrows_range<<<(unsigned)mg_div_up(nscope, 256), 256, 0, st>>>(a, lo, cnt);
		MG_LAUNCH_CHECK(ctx);
		if (mg_excl_scan_i64(ctx, cnt, nscope, off, &nrows, st))
			return 1;
	}
	mg_rows *r = new (std::nothrow) mg_rows();
	MG_REQUIRE(ctx, r, "out of memory");
	const size_t nr = (size_t)(nrows > 0 ? nrows : 1);
	cudaError_t e = cudaMalloc(&r->d_block, nr * 28);
	if (e != cudaSuccess) {
		delete r;
		cudaFree(tmp);
		return mg_fail(ctx, "mg_rows_intervals: %s", cudaGetErrorString(e));
	}
	char *base = (char *)r->d_block;
	r->owns = true;
	r->src.kind = 0;
	r->src.n = nrows;
	r->src.start = (const int64_t *)base;
	r->src.end = (const int64_t *)(base + nr * 8);
	r->src.scope_idx = (const int64_t *)(base + nr * 16);
	r->src.chrom = (const int32_t *)(base + nr * 24);
	if (nrows) {
		k_ivrows_write<<<(unsigned)mg_div_up(nrows, 256), 256, 0, st>>>(a, lo, off, nrows, (int32_t *)r->src.chrom, (int64_t *)r->src.start,
																		 (int64_t *)r->src.end, (int64_t *)r->src.scope_idx);
		MG_LAUNCH_CHECK(ctx);
	}
	MG_CUDA(ctx, cudaStreamSynchronize(st));
	cudaFree(tmp);
	*out = r;
	return 0;
}

// ---- sparse tracks and interval sets -----------------------------------------------------------------------------
template <typename H>
static int mg_sparse_like_create(mg_ctx *ctx, H **out)
{
	H *t = new (std::nothrow) H();
	MG_REQUIRE(ctx, t, "out of memory");
	t->nchroms = ctx->nchroms;
	t->h_table.assign(ctx->nchroms, SparseChrom{nullptr, nullptr, nullptr, 0, nullptr, 0, 0});
	if (mg_alloc_tracked(ctx, t->allocs, t->bytes, ctx->nchroms, &t->d_table))
		return 1;
	MG_CUDA(ctx, cudaMemcpy(t->d_table, t->h_table.data(), sizeof(SparseChrom) * ctx->nchroms, cudaMemcpyHostToDevice));
	*out = t;
	return 0;
}

template <typename H>
static int mg_sparse_like_stage(mg_ctx *ctx, H *t, int chromid, const int64_t *h_start, const int64_t *h_end, const float *h_val, int64_t n,
								const char *what)
{
	MG_REQUIRE(ctx, chromid >= 0 && chromid < t->nchroms, "%s: chromid %d out of range", what, chromid);
	MG_REQUIRE(ctx, n >= 0, "%s: negative record count", what);
	MG_REQUIRE(ctx, t->h_table[chromid].start == nullptr, "%s: chromosome %d already staged", what, chromid);
	// the reference's load-time validation (src/GenomeTrackSparse.cpp:209-222)
	for (int64_t i = 0; i < n; ++i) {
		MG_REQUIRE(ctx, h_start[i] >= 0, "%s: interval %lld has negative start (%lld)", what, (long long)i, (long long)h_start[i]);
		MG_REQUIRE(ctx, h_start[i] < h_end[i], "%s: interval %lld has start (%lld) >= end (%lld)", what, (long long)i, (long long)h_start[i],
				   (long long)h_end[i]);
		MG_REQUIRE(ctx, i == 0 || h_start[i] >= h_end[i - 1], "%s: interval %lld [%lld, %lld) overlaps with previous interval [%lld, %lld)",
				   what, (long long)i, (long long)h_start[i], (long long)h_end[i], (long long)h_start[i - 1], (long long)h_end[i - 1]);
	}
	MG_CUDA(ctx, cudaSetDevice(ctx->device));
	SparseChrom sc{nullptr, nullptr, nullptr, n, nullptr, 0, 0};
	int64_t *ds = nullptr, *de = nullptr;
	float *dv = nullptr;
	if (mg_alloc_tracked(ctx, t->allocs, t->bytes, n, &ds) || mg_alloc_tracked(ctx, t->allocs, t->bytes, n, &de))
		return 1;
	if (n) {
		MG_CUDA(ctx, cudaMemcpy(ds, h_start, sizeof(int64_t) * n, cudaMemcpyHostToDevice));
		MG_CUDA(ctx, cudaMemcpy(de, h_end, sizeof(int64_t) * n, cudaMemcpyHostToDevice));
	}
	if (h_val) {
		if (mg_alloc_tracked(ctx, t->allocs, t->bytes, n, &dv))
			return 1;
		std::vector<float> clean(h_val, h_val + n);
		for (auto &v : clean)
			if (std::isinf(v))
				v = std::numeric_limits<float>::quiet_NaN(); // src/GenomeTrackSparse.cpp:205-206
		if (n)
			MG_CUDA(ctx, cudaMemcpy(dv, clean.data(), sizeof(float) * n, cudaMemcpyHostToDevice));
	}
	sc.start = ds;
	sc.end = de;
	sc.val = dv;
	{
		// coarse position index: cells of 2^shift bp holding ~8 records each; coarse[c] = first record with end > c << shift
		const int64_t csize = ctx->chrom_sizes[chromid];
		int shift = 8;
		while (shift < 40 && ((csize >> shift) + 2) > std::max<int64_t>(n / 8, 16))
			++shift;
		const int64_t ncoarse = (csize >> shift) + 2;
		std::vector<uint32_t> coarse((size_t)ncoarse);
		int64_t r = 0;
		for (int64_t c = 0; c < ncoarse; ++c) {
			const int64_t pos = c << shift;
			while (r < n && h_end[r] <= pos)
				++r;
			coarse[(size_t)c] = (uint32_t)r;
		}
		uint32_t *dc = nullptr;
		if (mg_alloc_tracked(ctx, t->allocs, t->bytes, ncoarse, &dc))
			return 1;
		MG_CUDA(ctx, cudaMemcpy(dc, coarse.data(), sizeof(uint32_t) * (size_t)ncoarse, cudaMemcpyHostToDevice));
		sc.coarse = dc;
		sc.ncoarse = ncoarse;
		sc.coarse_shift = shift;
	}
	t->h_table[chromid] = sc;
	MG_CUDA(ctx, cudaMemcpy(t->d_table + chromid, &sc, sizeof(SparseChrom), cudaMemcpyHostToDevice));
	return 0;
}

extern "C" int mg_sparse_create(mg_ctx *ctx, mg_sparse **out) { return mg_sparse_like_create(ctx, out); }
extern "C" int mg_sparse_stage(mg_ctx *ctx, mg_sparse *t, int chromid, const int64_t *h_start, const int64_t *h_end, const float *h_val, int64_t n)
{
	MG_REQUIRE(ctx, h_val || n == 0, "mg_sparse_stage: values are NULL");
	static const float dummy = 0;
	return mg_sparse_like_stage(ctx, t, chromid, h_start, h_end, h_val ? h_val : &dummy, n, "Invalid sparse track");
}
extern "C" void mg_sparse_free(mg_ctx *ctx, mg_sparse *t)
{
	if (!t)
		return;
	for (void *p : t->allocs)
		cudaFree(p);
	delete t;
}
extern "C" int mg_ivset_create(mg_ctx *ctx, mg_ivset **out) { return mg_sparse_like_create(ctx, out); }
extern "C" int mg_ivset_stage(mg_ctx *ctx, mg_ivset *s, int chromid, const int64_t *h_start, const int64_t *h_end, int64_t n)
{
	return mg_sparse_like_stage(ctx, s, chromid, h_start, h_end, nullptr, n, "Invalid interval set");
}
extern "C" void mg_ivset_free(mg_ctx *ctx, mg_ivset *s)
{
	if (!s)
		return;
	for (void *p : s->allocs)
		cudaFree(p);
	delete s;
}

// Rows per thread of the merge kernels (a thread block takes 256 x that many rows). Fixed-bin rows whose blocks are expected to
// hold few records -- the track's mean record density times the block's hull -- take 16: the per-block set-up latency is
// amortised over twice the rows (32 measured slower: nearest 0.44 ms against 0.39). Explicit rows (their windows live in
// per-thread register arrays) and dense records stay at 8.
static int mg_merge_rows_per_thread(const mg_ctx *ctx, const mg_rows *rows, int64_t nrec)
{
	if (rows->src.kind != 1 || ctx->merge_rows == 8)
		return 8;
	double genome = 0;
	for (int64_t sz : ctx->chrom_sizes)
		genome += (double)sz;
	const double per_block = genome > 0 ? (double)nrec / genome * (double)rows->src.binsize * MG_SP_THREADS * 16 : 0;
	return per_block <= MG_SF_CAP / 2 ? 16 : 8;
}

// The regular-grid blocks of a merge-kernel launch (GridTable, sparse.cuh): the predicate of mg_sf_grid_block evaluated on the
// host. Every condition is a lower or an upper bound on the block index, so per scope interval the qualifying blocks form one range.
static void mg_make_grid_table(const mg_ctx *ctx, const mg_rows *rows, int64_t first, int64_t count, int64_t sshift, int64_t eshift, int rpt,
							   GridTable &gt)
{
	memset(&gt, 0, sizeof(gt));
	gt.n = -1;
	const RowSrc &rs = rows->src;
	if (rs.kind != 1 || rs.nscope > MG_SCOPE_INLINE)
		return;
	gt.n = 0;
	const int64_t R = (int64_t)MG_SP_THREADS * rpt, binsize = rs.binsize, len = binsize + eshift - sshift;
	if (len <= 0 || len > MG_SF_RANGE || count <= 0)
		return;
	gt.len = (int)len;
	gt.step = R * binsize;
	const int64_t nblocks = mg_div_up(count, R);
	for (int64_t k = 0; k < rs.nscope && gt.n < MG_SCOPE_INLINE; ++k) {
		const int64_t row0 = rows->h_srow0[k], row1 = rows->h_srow0[k + 1], sbin0 = rows->h_sstart[k] / binsize;
		const int64_t csize = ctx->chrom_sizes[rows->h_schrom[k]];
		auto grid_block = [&](int64_t b) {
			const int64_t nrows = std::min(R, count - b * R), r0 = first + b * R, r1 = r0 + nrows - 1;
			if (!(r0 > row0 && r1 < row1 - 1))
				return false;
			const int64_t w0 = (sbin0 + (r0 - row0)) * binsize + sshift, span = (nrows - 1) * binsize + len;
			return !(w0 < 0 || span > MG_SF_RANGE || w0 + span > csize);
		};
		// blocks whose first row lies in this scope interval
		int64_t bf = row0 >= first ? mg_div_up(row0 - first, R) : 0, bl = row1 - 1 >= first ? (row1 - 1 - first) / R : -1;
		bl = std::min(bl, nblocks - 1);
		while (bf <= bl && !grid_block(bf))
			++bf;
		while (bl > bf && !grid_block(bl))
			--bl;
		if (bf > bl || bl >= (1ll << 32))
			continue;
		gt.bfirst[gt.n] = (unsigned)bf;
		gt.blast[gt.n] = (unsigned)bl;
		gt.origin0[gt.n] = (sbin0 + (first + bf * R - row0)) * binsize + sshift;
		gt.chrom[gt.n] = rows->h_schrom[k];
		gt.n++;
	}
}

extern "C" int mg_sparse_window(mg_ctx *ctx, const mg_sparse *t, const mg_rows *rows, int64_t first, int64_t count, int64_t sshift,
								int64_t eshift, int nfuncs, const int *funcs, const double *params, double *const *d_out, void *stream)
{
	MG_REQUIRE(ctx, first >= 0 && count >= 0 && first + count <= rows->src.n, "row range out of bounds");
	MG_REQUIRE(ctx, nfuncs > 0, "no functions requested");
	SparseQuery q;
	memset(&q, 0, sizeof(q));
	q.rows = rows->src;
	q.first = first;
	q.count = count;
	q.sshift = sshift;
	q.eshift = eshift;
	q.table = t->d_table;
	q.chrom_sizes = ctx->d_chrom_sizes;
	for (int k = 0; k < nfuncs; ++k) {
		const int f = funcs[k]; // the sparse reader has one path: MG_FUNC_GENERIC is a dense-track flag and an unknown id here
		MG_REQUIRE(ctx, f >= 0 && f < MG_NUM_FUNCS, "unknown function id %d", f);
		MG_REQUIRE(ctx, d_out[k], "output column %d is NULL", k);
		if (f == MG_QUANTILE) {
			MG_REQUIRE(ctx, q.nq < MG_MAX_Q && params, "bad quantile request");
			q.q_pct[q.nq] = params[k];
			q.q_out[q.nq++] = d_out[k];
		} else {
			MG_REQUIRE(ctx, !q.out[f], "function %d requested twice in one call", f);
			q.out[f] = d_out[k];
			if (f >= MG_FIRST_POS && f <= MG_MAX_POS && params && params[k] != 0.)
				q.pos_rel |= 1u << f;
		}
	}
	if (!count)
		return 0;
	MG_REQUIRE(ctx, count < (1ll << 31) * (MG_SP_THREADS * 8), "row range too long for one launch");
	if (!q.out[MG_STDDEV] && q.nq == 0 && !ctx->force_sweep && !mg_any_out(q.out, MG_LSE, MG_MAX_POS)) {
		cudaStream_t st = mg_stream(stream);
		int64_t nrec = 0;
		for (const SparseChrom &c : t->h_table)
			nrec += c.n;
		const int rpt = mg_merge_rows_per_thread(ctx, rows, nrec);
		const unsigned grid = (unsigned)mg_div_up(count, MG_SP_THREADS * rpt);
		mg_make_grid_table(ctx, rows, first, count, sshift, eshift, rpt, q.grid);
#define MG_MERGE_CASE(F)                                                                                                 \
	if (rpt == 16)                                                                                                      \
		k_sparse_merge<F, 16><<<grid, MG_SP_THREADS, 0, st>>>(q);                                                        \
	else                                                                                                                \
		k_sparse_merge<F, 8><<<grid, MG_SP_THREADS, 0, st>>>(q);                                                         \
	break;
		switch (nfuncs == 1 ? funcs[0] : -1) { // one requested column: the kernel specialised on it
		case MG_AVG: MG_MERGE_CASE(MG_AVG)
		case MG_SUM: MG_MERGE_CASE(MG_SUM)
		case MG_MIN: MG_MERGE_CASE(MG_MIN)
		case MG_MAX: MG_MERGE_CASE(MG_MAX)
		case MG_NEAREST: MG_MERGE_CASE(MG_NEAREST)
		case MG_SIZE: MG_MERGE_CASE(MG_SIZE)
		case MG_EXISTS: MG_MERGE_CASE(MG_EXISTS)
		default: MG_MERGE_CASE(-1)
		}
#undef MG_MERGE_CASE
	} else
		k_sparse_window<<<(unsigned)mg_div_up(count, MG_SP_THREADS * MG_SP_ROWS), MG_SP_THREADS, 0, mg_stream(stream)>>>(q);
	MG_LAUNCH_CHECK(ctx);
	return 0;
}

extern "C" int mg_coverage(mg_ctx *ctx, const mg_ivset *s, const mg_rows *rows, int64_t first, int64_t count, int64_t sshift, int64_t eshift,
						   double *d_out, void *stream)
{
	MG_REQUIRE(ctx, first >= 0 && count >= 0 && first + count <= rows->src.n, "row range out of bounds");
	MG_REQUIRE(ctx, d_out, "output column is NULL");
	if (!count)
		return 0;
	CoverageQuery q;
	q.rows = rows->src;
	q.first = first;
	q.count = count;
	q.sshift = sshift;
	q.eshift = eshift;
	q.table = s->d_table;
	q.chrom_sizes = ctx->d_chrom_sizes;
	q.out = d_out;
	q.raster = ctx->cov_raster && rows->src.kind == 1 && rows->src.binsize < (1 << 20) ? 1 : 0;
	q.div = mg_covdiv_make(rows->src.kind == 1 ? rows->src.binsize : 1);
	if (ctx->force_sweep)
		k_coverage<<<(unsigned)mg_div_up(count, MG_SP_THREADS * MG_SP_ROWS), MG_SP_THREADS, 0, mg_stream(stream)>>>(q);
	else
	{
		int64_t nrec = 0;
		for (const SparseChrom &c : s->h_table)
			nrec += c.n;
		const int rpt = mg_merge_rows_per_thread(ctx, rows, nrec);
		mg_make_grid_table(ctx, rows, first, count, sshift, eshift, rpt, q.grid);
		if (rpt == 16)
			k_coverage_merge<16><<<(unsigned)mg_div_up(count, MG_SP_THREADS * 16), MG_SP_THREADS, 0, mg_stream(stream)>>>(q);
		else
			k_coverage_merge<8><<<(unsigned)mg_div_up(count, MG_SP_THREADS * 8), MG_SP_THREADS, 0, mg_stream(stream)>>>(q);
	}
	MG_LAUNCH_CHECK(ctx);
	return 0;
}

// ---- sequence + PWM -----------------------------------------------------------------------------------------------
extern "C" int mg_seq_create(mg_ctx *ctx, mg_seq **out)
{
	mg_seq *s = new (std::nothrow) mg_seq();
	MG_REQUIRE(ctx, s, "out of memory");
	s->nchroms = ctx->nchroms;
	s->h_table.assign(ctx->nchroms, SeqChrom{nullptr, nullptr, 0});
	if (mg_alloc_tracked(ctx, s->allocs, s->bytes, ctx->nchroms, &s->d_table))
		return 1;
	MG_CUDA(ctx, cudaMemcpy(s->d_table, s->h_table.data(), sizeof(SeqChrom) * ctx->nchroms, cudaMemcpyHostToDevice));
	*out = s;
	return 0;
}

extern "C" int mg_seq_stage(mg_ctx *ctx, mg_seq *s, int chromid, const char *h_ascii, int64_t len)
{
	MG_REQUIRE(ctx, chromid >= 0 && chromid < s->nchroms, "mg_seq_stage: chromid %d out of range", chromid);
	MG_REQUIRE(ctx, len == ctx->chrom_sizes[chromid], "mg_seq_stage: chromosome %d has %lld bases, expected %lld", chromid, (long long)len,
			   (long long)ctx->chrom_sizes[chromid]);
	MG_REQUIRE(ctx, s->h_table[chromid].codes == nullptr, "mg_seq_stage: chromosome %d already staged", chromid);
	MG_CUDA(ctx, cudaSetDevice(ctx->device));
	const int64_t nthreads = mg_div_up(len, 32);
	uint32_t *codes = nullptr, *inval = nullptr, *lower = nullptr;
	if (mg_alloc_tracked(ctx, s->allocs, s->bytes, nthreads * 2 + 8, &codes) || mg_alloc_tracked(ctx, s->allocs, s->bytes, nthreads + 8, &inval) ||
		mg_alloc_tracked(ctx, s->allocs, s->bytes, nthreads + 8, &lower))
		return 1;
	MG_CUDA(ctx, cudaMemset(lower, 0, sizeof(uint32_t) * (nthreads + 8)));
	MG_CUDA(ctx, cudaMemset(codes, 0, sizeof(uint32_t) * (nthreads * 2 + 8)));
	MG_CUDA(ctx, cudaMemset(inval, 0xff, sizeof(uint32_t) * (nthreads + 8)));
	uint8_t *d_ascii = nullptr;
	MG_CUDA(ctx, cudaMalloc(&d_ascii, (size_t)(len > 0 ? len : 1)));
	if (len) {
		MG_CUDA(ctx, cudaMemcpy(d_ascii, h_ascii, (size_t)len, cudaMemcpyHostToDevice));
		int *d_other = nullptr, h_other = 0;
		MG_CUDA(ctx, cudaMalloc(&d_other, sizeof(int)));
		MG_CUDA(ctx, cudaMemset(d_other, 0, sizeof(int)));
		k_seq_pack<<<(unsigned)mg_div_up(nthreads, 256), 256>>>(d_ascii, len, codes, inval, d_other);
		MG_LAUNCH_CHECK(ctx);
		k_seq_lower<<<(unsigned)mg_div_up(nthreads, 256), 256>>>(d_ascii, len, lower);
		MG_LAUNCH_CHECK(ctx);
		MG_CUDA(ctx, cudaMemcpy(&h_other, d_other, sizeof(int), cudaMemcpyDeviceToHost));
		cudaFree(d_other);
		if (h_other)
			s->has_other = true;
	}
	cudaFree(d_ascii);
	SeqChrom sc{codes, inval, len, lower};
	s->h_table[chromid] = sc;
	MG_CUDA(ctx, cudaMemcpy(s->d_table + chromid, &sc, sizeof(SeqChrom), cudaMemcpyHostToDevice));
	return 0;
}

extern "C" void mg_seq_free(mg_ctx *ctx, mg_seq *s)
{
	if (!s)
		return;
	for (void *p : s->allocs)
		cudaFree(p);
	delete s;
}

// src/PWMScorer.cpp:1452-1459 (double -> float), src/DnaPSSM.cpp:77-106 (normalize), :703-719 (prior)
extern "C" int mg_pssm_logp(const double *h_pssm, int L, double prior, float *h_logp)
{
	if (!h_pssm || !h_logp || L <= 0)
		return 1;
	for (int i = 0; i < L; ++i) {
		float p[4];
		for (int c = 0; c < 4; ++c)
			p[c] = (float)h_pssm[i * 4 + c];
		for (int pass = 0; pass < 2; ++pass) {
			const float sum = p[0] + p[1] + p[2] + p[3];
			for (int c = 0; c < 4; ++c) {
				p[c] /= sum;
				h_logp[i * 4 + c] = p[c] != 0 ? logf(p[c]) : -INFINITY;
			}
			if (pass == 1 || prior <= 0)
				break;
			for (int c = 0; c < 4; ++c)
				p[c] = p[c] + (float)prior;
		}
	}
	return 0;
}

extern "C" int mg_kmer(mg_ctx *ctx, const mg_seq *s, const char *kmer, int mode, int extend, int strand, const mg_rows *rows, int64_t first,
					   int64_t count, int64_t sshift, int64_t eshift, double *d_out, void *stream)
{
	MG_REQUIRE(ctx, first >= 0 && count >= 0 && first + count <= rows->src.n, "row range out of bounds");
	MG_REQUIRE(ctx, kmer && kmer[0], "Kmer string cannot be empty"); // src/KmerCounter.cpp:12-15
	MG_REQUIRE(ctx, mode == MG_KMER_COUNT || mode == MG_KMER_FRAC, "unsupported kmer mode %d", mode);
	MG_REQUIRE(ctx, strand == 0 || strand == 1 || strand == -1, "strand must be 0, 1 or -1");
	const size_t k = strlen(kmer);
	MG_REQUIRE(ctx, k <= 32, "kmer of %d bases: the device path compares at most 32", (int)k);
	KmerQuery q;
	q.fwd = q.rev = 0;
	for (size_t j = 0; j < k; ++j) {
		const char c = (char)toupper((unsigned char)kmer[j]);
		const char *at = strchr("ACGT", c);
		MG_REQUIRE(ctx, at != nullptr, "kmer \"%s\": the device path matches A, C, G, T only", kmer);
		const uint64_t code = (uint64_t)(at - "ACGT");
		q.fwd |= code << (2 * j);
		q.rev |= (3 - code) << (2 * (k - 1 - j));
	}
	if (!count)
		return 0;
	q.rows = rows->src;
	q.first = first;
	q.count = count;
	q.sshift = sshift;
	q.eshift = eshift;
	q.table = s->d_table;
	q.chrom_sizes = ctx->d_chrom_sizes;
	q.k = (int)k;
	q.mode = mode;
	q.extend = extend;
	q.strand = strand;
	q.out = d_out;
	const int64_t blocks_needed = mg_div_up(count, 8);
	const int64_t cap = (int64_t)ctx->sm_count * 16;
	k_kmer<<<(unsigned)(blocks_needed < cap ? blocks_needed : cap), 256, 0, mg_stream(stream)>>>(q);
	MG_LAUNCH_CHECK(ctx);
	return 0;
}

extern "C" int mg_masked(mg_ctx *ctx, const mg_seq *s, int mode, const mg_rows *rows, int64_t first, int64_t count, int64_t sshift, int64_t eshift,
						 double *d_out, void *stream)
{
	MG_REQUIRE(ctx, first >= 0 && count >= 0 && first + count <= rows->src.n, "row range out of bounds");
	MG_REQUIRE(ctx, mode == MG_MASKED_COUNT || mode == MG_MASKED_FRAC, "unsupported masked mode %d", mode);
	if (!count)
		return 0;
	MaskedQuery q;
	q.rows = rows->src;
	q.first = first;
	q.count = count;
	q.sshift = sshift;
	q.eshift = eshift;
	q.table = s->d_table;
	q.chrom_sizes = ctx->d_chrom_sizes;
	q.mode = mode;
	q.out = d_out;
	const int64_t blocks_needed = mg_div_up(count, 8);
	const int64_t cap = (int64_t)ctx->sm_count * 16;
	k_masked<<<(unsigned)(blocks_needed < cap ? blocks_needed : cap), 256, 0, mg_stream(stream)>>>(q);
	MG_LAUNCH_CHECK(ctx);
	return 0;
}

extern "C" int mg_pwm_spatial(mg_ctx *ctx, const mg_seq *s, const float *h_logp, int L, int bidirect, int extend, int mode, int strand,
							  float score_thresh, const float *h_spat_factor, int nspat, int spat_bin, const mg_rows *rows, int64_t first,
							  int64_t count, int64_t sshift, int64_t eshift, double *d_out, void *stream);
extern "C" int mg_pwm(mg_ctx *ctx, const mg_seq *s, const float *h_logp, int L, int bidirect, int extend, int mode, int strand,
					  float score_thresh, const mg_rows *rows, int64_t first, int64_t count, int64_t sshift, int64_t eshift, double *d_out,
					  void *stream)
{
	return mg_pwm_spatial(ctx, s, h_logp, L, bidirect, extend, mode, strand, score_thresh, nullptr, 0, 1, rows, first, count, sshift, eshift, d_out,
						  stream);
}

extern "C" int mg_pwm_spatial(mg_ctx *ctx, const mg_seq *s, const float *h_logp, int L, int bidirect, int extend, int mode, int strand,
							  float score_thresh, const float *h_spat_factor, int nspat, int spat_bin, const mg_rows *rows, int64_t first,
							  int64_t count, int64_t sshift, int64_t eshift, double *d_out, void *stream)
{
	MG_REQUIRE(ctx, nspat >= 0 && nspat <= MG_PWM_MAXSPAT && (nspat == 0 || h_spat_factor), "spatial weights: 0..%d factors", MG_PWM_MAXSPAT);
	MG_REQUIRE(ctx, nspat == 0 || mode != MG_PWM_MAX_POS, "pwm.max.pos under spatial weights is not offered (the reference's fresh and sliding "
			   "evaluations disagree on it)");
	MG_REQUIRE(ctx, first >= 0 && count >= 0 && first + count <= rows->src.n, "row range out of bounds");
	MG_REQUIRE(ctx, L >= 1 && L <= MG_PWM_MAXL, "PSSM length %d not in [1, %d]", L, MG_PWM_MAXL);
	MG_REQUIRE(ctx, mode == MG_PWM_TOTAL || mode == MG_PWM_MAX || mode == MG_PWM_COUNT || mode == MG_PWM_MAX_POS, "unsupported pwm mode %d", mode);
	MG_REQUIRE(ctx, mode != MG_PWM_MAX_POS || !s->has_other,
			   "pwm.max.pos: the staged sequence holds characters other than ACGT / N / *, which the packed form cannot tell from N");
	MG_REQUIRE(ctx, strand == 1 || strand == -1, "strand must be 1 or -1");
	if (!count)
		return 0;
	cudaStream_t st = mg_stream(stream);
	MgScratch mg_scratch(st);
	std::vector<float2> tab((size_t)L * 5);
	for (int j = 0; j < L; ++j) {
		for (int c = 0; c < 4; ++c)
			tab[j * 5 + c] = make_float2(h_logp[j * 4 + c], h_logp[(L - 1 - j) * 4 + (3 - c)]);
		tab[j * 5 + 4] = make_float2(-INFINITY, -INFINITY);
		if (mode == MG_PWM_MAX_POS) { // max_like_match scores N as the row's average log-probability (src/DnaPSSM.cpp:306-308, DnaPSSM.h:133-135)
			const float *a = h_logp + j * 4, *b = h_logp + (L - 1 - j) * 4;
			tab[j * 5 + 4] = make_float2((a[0] + a[1] + a[2] + a[3]) / 4, (b[0] + b[1] + b[2] + b[3]) / 4);
		}
	}
	float2 *d_tab = nullptr;
	MG_CUDA(ctx, mg_scratch.alloc(&d_tab, sizeof(float2) * tab.size()));
	MG_CUDA(ctx, cudaMemcpyAsync(d_tab, tab.data(), sizeof(float2) * tab.size(), cudaMemcpyHostToDevice, st));
	MG_CUDA(ctx, cudaStreamSynchronize(st)); // tab is a stack-lifetime host buffer
	PwmQuery q;
	q.rows = rows->src;
	q.first = first;
	q.count = count;
	q.sshift = sshift;
	q.eshift = eshift;
	q.table = s->d_table;
	q.chrom_sizes = ctx->d_chrom_sizes;
	q.tab = d_tab;
	q.L = L;
	q.bidirect = bidirect;
	q.extend = extend;
	q.mode = mode;
	q.strand = strand;
	q.thresh = score_thresh;
	q.out = d_out;
	q.spat_log = nullptr;
	q.nspat = nspat;
	q.spat_bin = spat_bin < 1 ? 1 : spat_bin; // std::max(1, spat_bin_size), src/PWMScorer.cpp:118
	if (nspat) {
		std::vector<float> sl((size_t)nspat);
		for (int k = 0; k < nspat; ++k) // src/PWMScorer.cpp:121-124, in float on the host like the reference
			sl[k] = std::log(std::max(1e-30f, h_spat_factor[k]));
		float *d_sl = nullptr;
		MG_CUDA(ctx, mg_scratch.alloc(&d_sl, sizeof(float) * (size_t)nspat));
		MG_CUDA(ctx, cudaMemcpyAsync(d_sl, sl.data(), sizeof(float) * (size_t)nspat, cudaMemcpyHostToDevice, st));
		MG_CUDA(ctx, cudaStreamSynchronize(st));
		q.spat_log = d_sl;
	}
	const int64_t blocks_needed = mg_div_up(count, MG_PWM_WARPS);
	const int64_t cap = (int64_t)ctx->sm_count * 16;
	const unsigned grid = (unsigned)(blocks_needed < cap ? blocks_needed : cap);
#define MG_PWM_LAUNCH(B, M)                                                                                              \
	do {                                                                                                                \
		if (ctx->pwm_strict)                                                                                            \
			k_pwm<B, M, true><<<grid, MG_PWM_WARPS * 32, 0, st>>>(q);                                                    \
		else                                                                                                            \
			k_pwm<B, M, false><<<grid, MG_PWM_WARPS * 32, 0, st>>>(q);                                                   \
	} while (0)
	if (mode == MG_PWM_MAX_POS) { // no transcendental in this mode: one instantiation per strand handling
		if (bidirect)
			k_pwm<true, MG_PWM_MAX_POS, false><<<grid, MG_PWM_WARPS * 32, 0, st>>>(q);
		else
			k_pwm<false, MG_PWM_MAX_POS, false><<<grid, MG_PWM_WARPS * 32, 0, st>>>(q);
	} else if (bidirect) {
		if (mode == MG_PWM_TOTAL) MG_PWM_LAUNCH(true, MG_PWM_TOTAL);
		else if (mode == MG_PWM_MAX) MG_PWM_LAUNCH(true, MG_PWM_MAX);
		else MG_PWM_LAUNCH(true, MG_PWM_COUNT);
	} else {
		if (mode == MG_PWM_TOTAL) MG_PWM_LAUNCH(false, MG_PWM_TOTAL);
		else if (mode == MG_PWM_MAX) MG_PWM_LAUNCH(false, MG_PWM_MAX);
		else MG_PWM_LAUNCH(false, MG_PWM_COUNT);
	}
#undef MG_PWM_LAUNCH
	MG_LAUNCH_CHECK(ctx);
	return 0;
}

// ---- host-buffer forms of the sparse / coverage / PWM calls (the same three-slot pipeline as mg_h_dense_window) -------------
extern "C" int mg_h_sparse_window(mg_ctx *ctx, const mg_sparse *t, int64_t n, const int32_t *h_chrom, const int64_t *h_start, const int64_t *h_end,
								  int64_t sshift, int64_t eshift, int nfuncs, const int *funcs, const double *params, double *const *h_out)
{
	return mg_h_pipeline(ctx, n, h_chrom, h_start, h_end, nfuncs, h_out, [&](const mg_rows &rows, int64_t m, double *const *d_cols, cudaStream_t st) {
		return mg_sparse_window(ctx, t, &rows, 0, m, sshift, eshift, nfuncs, funcs, params, d_cols, st);
	});
}

extern "C" int mg_h_coverage(mg_ctx *ctx, const mg_ivset *s, int64_t n, const int32_t *h_chrom, const int64_t *h_start, const int64_t *h_end,
							 int64_t sshift, int64_t eshift, double *h_out)
{
	double *cols[1] = {h_out};
	return mg_h_pipeline(ctx, n, h_chrom, h_start, h_end, 1, cols, [&](const mg_rows &rows, int64_t m, double *const *d_cols, cudaStream_t st) {
		return mg_coverage(ctx, s, &rows, 0, m, sshift, eshift, d_cols[0], st);
	});
}

extern "C" int mg_h_pwm(mg_ctx *ctx, const mg_seq *s, const float *h_logp, int L, int bidirect, int extend, int mode, int strand, float score_thresh,
						int64_t n, const int32_t *h_chrom, const int64_t *h_start, const int64_t *h_end, int64_t sshift, int64_t eshift, double *h_out)
{
	double *cols[1] = {h_out};
	return mg_h_pipeline(ctx, n, h_chrom, h_start, h_end, 1, cols, [&](const mg_rows &rows, int64_t m, double *const *d_cols, cudaStream_t st) {
		return mg_pwm(ctx, s, h_logp, L, bidirect, extend, mode, strand, score_thresh, &rows, 0, m, sshift, eshift, d_cols[0], st);
	});
}

// ---- gsummary ------------------------------------------------------------------------------------------------------
extern "C" int mg_summary_init(mg_ctx *ctx, double *d_partial, void *stream)
{
	k_summary_init<<<1, 1, 0, mg_stream(stream)>>>(d_partial);
	MG_LAUNCH_CHECK(ctx);
	return 0;
}

static int mg_summary_blocks(mg_ctx *ctx, int64_t n)
{
	int64_t b = mg_div_up(n, MG_SUM_THREADS);
	const int64_t cap = (int64_t)ctx->sm_count * 8;
	if (b > cap) b = cap;
	if (b > MG_SUM_MAX_BLOCKS) b = MG_SUM_MAX_BLOCKS;
	return (int)(b < 1 ? 1 : b);
}

static int mg_summary_impl(mg_ctx *ctx, const double *d_vals, int64_t n, double *d_partial, void *stream, bool narrow)
{
	if (n <= 0)
		return 0;
	cudaStream_t st = mg_stream(stream);
	MgScratch mg_scratch(st);
	const int nb = mg_summary_blocks(ctx, n);
	Moments *scratch = nullptr;
	MG_CUDA(ctx, mg_scratch.alloc(&scratch, sizeof(Moments) * nb));
	if (narrow)
		k_summary_values<true><<<nb, MG_SUM_THREADS, 0, st>>>(d_vals, n, scratch);
	else
		k_summary_values<false><<<nb, MG_SUM_THREADS, 0, st>>>(d_vals, n, scratch);
	MG_LAUNCH_CHECK(ctx);
	k_summary_final<<<1, MG_SUM_THREADS, 0, st>>>(scratch, nb, d_partial);
	MG_LAUNCH_CHECK(ctx);
	return 0;
}

extern "C" int mg_summary(mg_ctx *ctx, const double *d_vals, int64_t n, double *d_partial, void *stream)
{
	return mg_summary_impl(ctx, d_vals, n, d_partial, stream, false);
}

// Per-stream device scratch (a zeroed 256-byte header + MG_SUM_MAX_BLOCKS block partials), allocated on first use and kept:
// kernels that finish with a "last block reduces" step need it without a malloc / free pair per call.
static int mg_stream_scratch(mg_ctx *ctx, cudaStream_t st, void **out)
{
	auto it = ctx->stream_scratch.find((void *)st);
	if (it == ctx->stream_scratch.end()) {
		void *p = nullptr;
		const size_t bytes = 256 + sizeof(Moments) * MG_SUM_MAX_BLOCKS;
		MG_CUDA(ctx, cudaMalloc(&p, bytes));
		MG_CUDA(ctx, cudaMemset(p, 0, bytes));
		it = ctx->stream_scratch.emplace((void *)st, p).first;
	}
	*out = it->second;
	return 0;
}

static bool mg_is_prefix_func(int f) { return f == MG_AVG || f == MG_SUM || f == MG_NEAREST || f == MG_SIZE || f == MG_EXISTS; }

// implicit fixed-bin rows at the track's own resolution with no shifts: each row is exactly one bin
static bool mg_is_stream_case(const mg_dense *t, const mg_rows *rows, int64_t sshift, int64_t eshift, int func)
{
	return rows->src.kind == 1 && rows->src.binsize == (int64_t)t->bin_size && sshift == 0 && eshift == 0 &&
		   (func == MG_AVG || func == MG_NEAREST || func == MG_SUM || func == MG_MIN || func == MG_MAX);
}

// Stream case over few scope intervals: contiguous bin segments of rows [first, first + count)
struct StreamSeg {
	const float *p;
	int64_t n;
};
static void mg_make_segtable(const std::vector<StreamSeg> &segs, SegTable &st)
{
	memset(&st, 0, sizeof(st));
	st.nseg = (int)segs.size();
	int64_t tiles = 0;
	for (int i = 0; i < st.nseg; ++i) {
		const int64_t mis = (int64_t)((reinterpret_cast<uintptr_t>(segs[i].p) >> 2) & 3);
		st.p[i] = segs[i].p;
		st.n[i] = segs[i].n;
		st.tile0[i] = tiles;
		tiles += mg_div_up(mis + segs[i].n, MG_SEG_TILE);
	}
	for (int i = st.nseg; i <= MG_MAX_STREAM_SEGS; ++i)
		st.tile0[i] = tiles;
}
static bool mg_stream_segments(const mg_dense *t, const mg_rows *rows, int64_t first, int64_t count, std::vector<StreamSeg> &segs)
{
	segs.clear();
	const int64_t last = first + count;
	const size_t ns = rows->h_schrom.size();
	// first scope interval that holds row `first`
	size_t k = std::upper_bound(rows->h_srow0.begin(), rows->h_srow0.end(), first) - rows->h_srow0.begin();
	k = k ? k - 1 : 0;
	for (; k < ns && rows->h_srow0[k] < last; ++k) {
		const int64_t r0 = std::max(first, rows->h_srow0[k]), r1 = std::min(last, rows->h_srow0[k + 1]);
		if (r1 <= r0)
			continue;
		const DenseChrom &ch = t->h_table[rows->h_schrom[k]];
		if (!ch.vals)
			return false;
		const int64_t b0 = rows->h_sstart[k] / rows->src.binsize + (r0 - rows->h_srow0[k]);
		int64_t n = r1 - r0;
		if (b0 + n > ch.nbins) // rows past the last stored bin read as NaN in the generic path: keep that path
			return false;
		if (segs.size() >= MG_MAX_STREAM_SEGS)
			return false;
		segs.push_back(StreamSeg{ch.vals + b0, n});
	}
	return true;
}

extern "C" int mg_dense_summary(mg_ctx *ctx, const mg_dense *t, const mg_rows *rows, int64_t first, int64_t count, int64_t sshift,
								int64_t eshift, int func, double param, double *d_partial, void *stream)
{
	MG_REQUIRE(ctx, first >= 0 && count >= 0 && first + count <= rows->src.n, "row range out of bounds");
	if (!count)
		return 0;
	cudaStream_t st = mg_stream(stream);
	MgScratch mg_scratch(st);
	std::vector<StreamSeg> segs;
	if (mg_is_stream_case(t, rows, sshift, eshift, func) && mg_stream_segments(t, rows, first, count, segs)) {
		// one pure-streaming launch over every contiguous segment, then the fixed-order final reduction
		SegTable tab;
		mg_make_segtable(segs, tab);
		const int total_blocks = (int)std::max<int64_t>(1, std::min<int64_t>(tab.tile0[tab.nseg], std::min<int64_t>((int64_t)ctx->sm_count * 8, MG_SUM_MAX_BLOCKS)));
		void *scratch = nullptr;
		if (mg_stream_scratch(ctx, st, &scratch))
			return 1;
		k_segs_summary<<<total_blocks, MG_SUM_THREADS, 0, st>>>(tab, count, (Moments *)((char *)scratch + 256), (unsigned *)scratch, d_partial);
		MG_LAUNCH_CHECK(ctx);
		return 0;
	}
	if (mg_is_stream_case(t, rows, sshift, eshift, func) || mg_is_prefix_func(func)) {
		const int nb = mg_summary_blocks(ctx, count);
		Moments *scratch = nullptr;
		MG_CUDA(ctx, mg_scratch.alloc(&scratch, sizeof(Moments) * nb));
		if (mg_is_stream_case(t, rows, sshift, eshift, func)) {
			k_dense_stream_summary<<<nb, MG_SUM_THREADS, 0, st>>>(rows->src, t->d_table, first, count, scratch);
		} else {
			DenseQuery q;
			bool need_sweep;
			double *dummy = (double *)d_partial; // mg_fill_dense_query wants a non-null column; never written by k_dense_summary
			if (mg_fill_dense_query(ctx, t, rows, first, count, sshift, eshift, 1, &func, &param, &dummy, q, need_sweep))
				return 1;
			k_dense_summary<<<nb, MG_SUM_THREADS, 0, st>>>(q, func, scratch);
		}
		MG_LAUNCH_CHECK(ctx);
		k_summary_final<<<1, MG_SUM_THREADS, 0, st>>>(scratch, nb, d_partial);
		MG_LAUNCH_CHECK(ctx);
		return 0;
	}
	// any other function: materialise the column on the device, then reduce it (values never leave HBM)
	double *col = nullptr;
	MG_CUDA(ctx, mg_scratch.alloc(&col, sizeof(double) * count));
	if (mg_dense_window(ctx, t, rows, first, count, sshift, eshift, 1, &func, &param, &col, stream) || mg_summary(ctx, col, count, d_partial, stream))
		return 1;
	return 0;
}

// src/GenomeTrackSummary.cpp:148-155 with get_mean / get_stdev :30-40
extern "C" void mg_summary_finalize(const double p[6], double out[7])
{
	const double n = p[0], nn = p[1], total = p[2], mn = p[3], mx = p[4], ssq = p[5];
	const double nan = std::numeric_limits<double>::quiet_NaN();
	out[0] = n;
	out[1] = n - nn;
	out[2] = nn ? mn : nan;
	out[3] = nn ? mx : nan;
	out[4] = nn ? total : nan;
	out[5] = nn ? total / nn : nan;
	if (nn > 1) {
		const double mean = total / nn;
		const double var = ssq / (nn - 1) - (mean * mean) * (nn / (nn - 1));
		out[6] = var > 0 ? sqrt(var) : 0;
	} else
		out[6] = nan;
}

extern "C" void mg_summary_finalize_rows(const double *partials, int64_t n, double *out)
{
	for (int64_t i = 0; i < n; ++i)
		mg_summary_finalize(partials + 6 * i, out + 7 * i);
}

// ---- gdist -----------------------------------------------------------------------------------------------------------
static int mg_make_hist_spec(mg_ctx *ctx, MgScratch &mg_scratch, int ndim, const double *h_breaks, const int *nbreaks, int include_lowest, HistSpec &hs, double **d_breaks,
							 cudaStream_t st)
{
	MG_REQUIRE(ctx, ndim >= 1 && ndim <= MG_HIST_MAX_DIMS, "histogram dimension %d not in [1, %d]", ndim, MG_HIST_MAX_DIMS);
	memset(&hs, 0, sizeof(hs));
	hs.ndim = ndim;
	hs.include_lowest = include_lowest;
	long long total = 1;
	int off = 0;
	for (int d = 0; d < ndim; ++d) {
		// BinFinder::init (src/BinFinder.cpp:10-43)
		MG_REQUIRE(ctx, nbreaks[d] >= 2, "Invalid number of breaks %d", nbreaks[d]);
		const double *b = h_breaks + off;
		double binsize = b[1] - b[0];
		for (int i = 1; i < nbreaks[d]; ++i) {
			MG_REQUIRE(ctx, b[i] != b[i - 1], "Breaks are not unique (break[%d]=break[%d]=%g)", i - 1, i, b[i]);
			MG_REQUIRE(ctx, b[i] > b[i - 1], "Breaks are not sorted (break[%d]=%g, break[%d]=%g)", i - 1, b[i - 1], i, b[i]);
			if ((float)(b[i] - b[i - 1]) != (float)binsize)
				binsize = 0;
		}
		if (!std::isfinite(binsize))
			binsize = 0;
		hs.nbreaks[d] = nbreaks[d];
		hs.boff[d] = off;
		hs.mult[d] = total;
		hs.binsize[d] = binsize;
		total *= nbreaks[d] - 1;
		off += nbreaks[d];
	}
	hs.total = total;
	if (!d_breaks) // the caller only wants the validated description (breaks travel another way)
		return 0;
	MG_CUDA(ctx, mg_scratch.alloc(d_breaks, sizeof(double) * off));
	MG_CUDA(ctx, cudaMemcpyAsync(*d_breaks, h_breaks, sizeof(double) * off, cudaMemcpyHostToDevice, st));
	MG_CUDA(ctx, cudaStreamSynchronize(st));
	hs.breaks = *d_breaks;
	return 0;
}

// BinFinder::val2bin on the host (src/BinFinder.h:55-81), extended monotonically: -1 below the first break, nbins above the last
static int mg_host_bin_monotone(const double *br, int nbreaks, double binsize, int include_lowest, double val)
{
	const int numbins = nbreaks - 1;
	if (include_lowest && val == br[0])
		return 0;
	if (val <= br[0])
		return -1;
	if (val > br[nbreaks - 1])
		return numbins;
	if (binsize != 0.) {
		const int b = (int)ceil((val - br[0]) / binsize) - 1;
		return b < numbins - 1 ? b : numbins - 1;
	}
	unsigned s = 0, e = (unsigned)numbins;
	while (e - s > 1) {
		const unsigned mid = (s + e) / 2;
		if (val <= br[mid])
			e = mid;
		else
			s = mid;
	}
	return (int)s;
}

static inline uint32_t mg_host_fkey(float v)
{
	uint32_t b;
	memcpy(&b, &v, 4);
	return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}
static inline float mg_host_fkey_inv(uint32_t k)
{
	const uint32_t b = k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu);
	float v;
	memcpy(&v, &b, 4);
	return v;
}

// thr[j], j = 0..nb: the largest float whose (monotone) bin is < j. Exact: bisection over the float order with the
// double formula the reference uses.
static void mg_float_thresholds(const double *br, int nbreaks, double binsize, int include_lowest, std::vector<float> &thr)
{
	const int nb = nbreaks - 1;
	thr.resize(nb + 1);
	const uint32_t kmin = mg_host_fkey(-INFINITY), kmax = mg_host_fkey(INFINITY);
	for (int j = 0; j <= nb; ++j) {
		// smallest key in [kmin, kmax] whose bin >= j (kmax + 1 if none)
		uint64_t lo = kmin, hi = (uint64_t)kmax + 1;
		while (lo < hi) {
			const uint64_t mid = (lo + hi) >> 1;
			if (mg_host_bin_monotone(br, nbreaks, binsize, include_lowest, (double)mg_host_fkey_inv((uint32_t)mid)) >= j)
				hi = mid;
			else
				lo = mid + 1;
		}
		thr[j] = lo > kmax ? INFINITY : (lo == kmin ? -INFINITY : mg_host_fkey_inv((uint32_t)lo - 1));
	}
}

static size_t mg_hist_private_smem(const HistSpec &hs)
{
	const int nb = hs.nbreaks[0] - 1;
	return (size_t)(2 * hs.nbreaks[0]) * 4 + (size_t)(MG_HIST_THREADS / 32) * nb * 32 * 4;
}

static int mg_hist_blocks(mg_ctx *ctx, int64_t n)
{
	int64_t b = mg_div_up(n, MG_HIST_THREADS * 8);
	const int64_t cap = (int64_t)ctx->sm_count * 2;
	if (b > cap) b = cap;
	return (int)(b < 1 ? 1 : b);
}

extern "C" int mg_hist(mg_ctx *ctx, int ndim, const double *const *d_cols, int64_t n, const double *h_breaks, const int *nbreaks,
					   int include_lowest, uint64_t *d_counts, void *stream)
{
	cudaStream_t st = mg_stream(stream);
	MgScratch mg_scratch(st);
	HistSpec hs;
	double *d_breaks = nullptr;
	if (mg_make_hist_spec(ctx, mg_scratch, ndim, h_breaks, nbreaks, include_lowest, hs, &d_breaks, st))
		return 1;
	if (n > 0) {
		if (ndim == 1 && hs.nbreaks[0] - 1 <= MG_HIST_PRIVATE_MAX) {
			const size_t smem = mg_hist_private_smem(hs);
			MG_CUDA(ctx, cudaFuncSetAttribute(k_values_hist_private, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
			k_values_hist_private<<<mg_hist_blocks(ctx, n), MG_HIST_THREADS, smem, st>>>(d_cols[0], n, hs, (unsigned long long *)d_counts);
		} else {
			HistCols cols;
			for (int d = 0; d < ndim; ++d)
				cols.col[d] = d_cols[d];
			k_hist_nd<<<(unsigned)mg_div_up(n, MG_HIST_THREADS * 4), MG_HIST_THREADS, 0, st>>>(cols, n, hs, (unsigned long long *)d_counts);
		}
		MG_LAUNCH_CHECK(ctx);
	}
	return 0;
}

extern "C" int mg_dense_hist(mg_ctx *ctx, const mg_dense *t, const mg_rows *rows, int64_t first, int64_t count, int64_t sshift, int64_t eshift,
							 int func, double param, const double *h_breaks, int nbreaks, int include_lowest, uint64_t *d_counts, void *stream)
{
	MG_REQUIRE(ctx, first >= 0 && count >= 0 && first + count <= rows->src.n, "row range out of bounds");
	cudaStream_t st = mg_stream(stream);
	MgScratch mg_scratch(st);
	const bool stream_case = mg_is_stream_case(t, rows, sshift, eshift, func);
	if (!stream_case && !mg_is_prefix_func(func)) {
		if (!count)
			return 0;
		double *col = nullptr;
		MG_CUDA(ctx, mg_scratch.alloc(&col, sizeof(double) * count));
		const double *cols[1] = {col};
		if (mg_dense_window(ctx, t, rows, first, count, sshift, eshift, 1, &func, &param, &col, stream) ||
			mg_hist(ctx, 1, cols, count, h_breaks, &nbreaks, include_lowest, d_counts, stream))
			return 1;
		return 0;
	}
	HistSpec hs;
	double *d_breaks = nullptr;
	if (mg_make_hist_spec(ctx, mg_scratch, 1, h_breaks, &nbreaks, include_lowest, hs, nullptr, st))
		return 1;
	std::vector<StreamSeg> segs;
	if (count > 0 && stream_case && nbreaks - 1 <= MG_HIST_PRIVATE_MAX && mg_stream_segments(t, rows, first, count, segs)) {
		// pure streaming over the contiguous segments with exact float thresholds (cached per break set: the bisection
		// costs more host time than the kernel takes)
		const int nb = nbreaks - 1;
		if (ctx->thr_include_lowest != include_lowest || ctx->thr_breaks.size() != (size_t)nbreaks ||
			memcmp(ctx->thr_breaks.data(), h_breaks, sizeof(double) * nbreaks)) {
			mg_float_thresholds(h_breaks, nbreaks, hs.binsize[0], include_lowest, ctx->thr_values);
			ctx->thr_breaks.assign(h_breaks, h_breaks + nbreaks);
			ctx->thr_include_lowest = include_lowest;
		}
		const std::vector<float> &thr = ctx->thr_values;
		FloatThr ft;
		memset(&ft, 0, sizeof(ft));
		ft.nb = nb;
		ft.uniform = hs.binsize[0] != 0.;
		ft.b0 = (float)h_breaks[0];
		ft.inv_binsize = ft.uniform ? (float)(1. / hs.binsize[0]) : 0.f;
		ft.near = 2.f; // no value is "far from every break": the kernel always looks at the thresholds
		if (ft.uniform) {
			// how far (in bins) the kernel's float position t = (v - b0) * inv + 1 can lie from the value's true position among THESE breaks:
			// the roundings of b0 and inv, the breaks' own deviation from b0 + j * binsize, three float roundings of at most 2^-24 relative
			const double bs = hs.binsize[0], tmax = nb + 2.;
			double err = fabs(h_breaks[0] - (double)ft.b0) / bs + tmax * (fabs(1. / bs - (double)ft.inv_binsize) * bs + 3. * 5.960464477539063e-08);
			for (int j = 0; j < nbreaks; ++j)
				err = std::max(err, fabs(h_breaks[j] - (h_breaks[0] + j * bs)) / bs + tmax * 3. * 5.960464477539063e-08);
			const double near = 4. * err + 9.5367431640625e-07;
			if (near < 0.25)
				ft.near = (float)near;
		}
		memcpy(ft.thr, thr.data(), sizeof(float) * thr.size());
		SegTable tab;
		mg_make_segtable(segs, tab);
		// 16-bit private columns (half the shared memory, twice the resident warps) whenever no thread can count past 65535
		const size_t thr_bytes = (size_t)((nb + 3 + 3) & ~3) * 4, cols = (size_t)(MG_HIST_THREADS / 32) * nb * 32;
		const int64_t tiles = tab.tile0[tab.nseg];
		for (int pass = 0; pass < 2; ++pass) {
			const size_t smem = thr_bytes + cols * (pass == 0 ? 2 : 4);
			const int64_t per_sm = std::max<int64_t>(1, std::min<int64_t>(8, (int64_t)(220 * 1024) / (int64_t)(smem + 1024)));
			const int64_t blocks = std::max<int64_t>(1, std::min<int64_t>(tiles, (int64_t)ctx->sm_count * per_sm));
			if (pass == 0) {
				if (mg_div_up(tiles, blocks) * 16 > 65535)
					continue;
				if (ctx->hist_red) {
					MG_CUDA(ctx, cudaFuncSetAttribute(k_segs_hist<uint16_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
					k_segs_hist<uint16_t, true><<<(unsigned)blocks, MG_HIST_THREADS, smem, st>>>(tab, ft, (unsigned long long *)d_counts);
				} else {
					MG_CUDA(ctx, cudaFuncSetAttribute(k_segs_hist<uint16_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
					k_segs_hist<uint16_t, false><<<(unsigned)blocks, MG_HIST_THREADS, smem, st>>>(tab, ft, (unsigned long long *)d_counts);
				}
			} else {
				MG_CUDA(ctx, cudaFuncSetAttribute(k_segs_hist<uint32_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
				k_segs_hist<uint32_t, false><<<(unsigned)blocks, MG_HIST_THREADS, smem, st>>>(tab, ft, (unsigned long long *)d_counts);
			}
			break;
		}
		MG_LAUNCH_CHECK(ctx);
		return 0;
	}
	if (mg_make_hist_spec(ctx, mg_scratch, 1, h_breaks, &nbreaks, include_lowest, hs, &d_breaks, st))
		return 1;
	if (count > 0) {
		const bool priv = nbreaks - 1 <= MG_HIST_PRIVATE_MAX;
		DenseQuery q;
		bool need_sweep;
		double *dummy = (double *)d_counts;
		if (mg_fill_dense_query(ctx, t, rows, first, count, sshift, eshift, 1, &func, &param, &dummy, q, need_sweep))
			return 1;
		if (priv) {
			const size_t smem = mg_hist_private_smem(hs);
			const int nb = mg_hist_blocks(ctx, count);
			if (stream_case) {
				MG_CUDA(ctx, cudaFuncSetAttribute(k_dense_stream_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
				k_dense_stream_hist<<<nb, MG_HIST_THREADS, smem, st>>>(rows->src, t->d_table, first, count, hs, (unsigned long long *)d_counts);
			} else {
				MG_CUDA(ctx, cudaFuncSetAttribute(k_dense_hist_private, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
				k_dense_hist_private<<<nb, MG_HIST_THREADS, smem, st>>>(q, func, hs, (unsigned long long *)d_counts);
			}
		} else
			k_dense_hist_global<<<(unsigned)mg_div_up(count, MG_HIST_THREADS * 4), MG_HIST_THREADS, 0, st>>>(q, func, hs, (unsigned long long *)d_counts);
		MG_LAUNCH_CHECK(ctx);
	}
	return 0;
}

// ---- gquantiles --------------------------------------------------------------------------------------------------------
// The selection over a column that lies in pieces on several devices (one piece = the whole-device case): every counting pass runs
// on each piece on its own device, the digit histograms (<= 1 MB a pass) meet on the host, which steers the selection anyway.
// Uses the legacy default stream of every device.
static int mg_quantiles_multi(int nsrc, mg_ctx *const *ctxs, const double *const *d_vals, const int64_t *ns, int nq, const double *h_percentiles,
							  double *h_out, int64_t *h_nkept, const cudaStream_t *sts = nullptr)
{
	mg_ctx *ctx = ctxs[0];
	MG_REQUIRE(ctx, nsrc >= 1 && nq >= 1 && nq <= MG_QS_MAXSLOTS / 2 && h_percentiles && h_out, "mg_quantiles: bad arguments (1..%d quantiles)",
			   MG_QS_MAXSLOTS / 2);
	for (int k = 0; k < nq; ++k) // C_gquantiles rejects these before scanning (src/GenomeTrackQuantiles.cpp:140-143)
		MG_REQUIRE(ctx, h_percentiles[k] >= 0 && h_percentiles[k] <= 1, "Percentile (%g) is not in [0, 1] range\n", h_percentiles[k]);
	const double nan = std::numeric_limits<double>::quiet_NaN();
	const size_t hist_words = (size_t)MG_QS_MAXSLOTS * 2048 + 1;
	struct Piece {
		unsigned long long *d_hist = nullptr;
		std::vector<unsigned long long> h;
	};
	std::vector<Piece> pc(nsrc);
	struct Release {
		std::vector<Piece> &pc;
		mg_ctx *const *ctxs;
		const cudaStream_t *sts;
		~Release()
		{
			for (size_t s = 0; s < pc.size(); ++s)
				if (pc[s].d_hist) {
					cudaSetDevice(ctxs[s]->device);
					cudaFreeAsync(pc[s].d_hist, sts ? sts[s] : nullptr);
				}
		}
	} release{pc, ctxs, sts};
	int64_t n = 0;
	for (int s = 0; s < nsrc; ++s) {
		MG_REQUIRE(ctx, ns[s] >= 0, "mg_quantiles: negative row count");
		n += ns[s];
		MG_CUDA(ctx, cudaSetDevice(ctxs[s]->device));
		MG_CUDA(ctx, cudaMallocAsync(&pc[s].d_hist, hist_words * 8, sts ? sts[s] : nullptr));
		pc[s].h.resize(hist_words);
	}
	std::vector<unsigned long long> h(hist_words);
	// one counting pass on every piece, then the sum of the pieces' histograms in h[0 .. words)
	auto count_pass = [&](const QsPass &p, size_t words, bool first) -> int {
		for (int s = 0; s < nsrc; ++s) {
			cudaStream_t st = sts ? sts[s] : nullptr;
			MG_CUDA(ctx, cudaSetDevice(ctxs[s]->device));
			MG_CUDA(ctx, cudaMemsetAsync(pc[s].d_hist, 0, words * 8, st));
			if (ns[s]) {
				const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(mg_div_up(ns[s], 256 * 8), (int64_t)ctxs[s]->sm_count * 8));
				if (first)
					k_qs_count<<<grid, 256, 0, st>>>(d_vals[s], ns[s], p, pc[s].d_hist + 1, pc[s].d_hist);
				else
					k_qs_count<<<grid, 256, 0, st>>>(d_vals[s], ns[s], p, pc[s].d_hist, nullptr);
				MG_LAUNCH_CHECK(ctxs[s]);
			}
			MG_CUDA(ctx, cudaMemcpyAsync(pc[s].h.data(), pc[s].d_hist, words * 8, cudaMemcpyDeviceToHost, st));
		}
		std::fill(h.begin(), h.begin() + words, 0ull);
		for (int s = 0; s < nsrc; ++s) {
			MG_CUDA(ctx, cudaSetDevice(ctxs[s]->device));
			MG_CUDA(ctx, cudaStreamSynchronize(sts ? sts[s] : nullptr));
			for (size_t w = 0; w < words; ++w)
				h[w] += pc[s].h[w];
		}
		return 0;
	};
	// pass 0: top 11 bits of every key + the NaN count
	QsPass p;
	memset(&p, 0, sizeof(p));
	if (count_pass(p, 2048 + 1, true))
		return 1;
	const int64_t N = n - (int64_t)h[0];
	if (h_nkept)
		*h_nkept = N;
	if (N <= 0) {
		for (int k = 0; k < nq; ++k)
			h_out[k] = nan;
		return 0;
	}
	// target ranks: floor / ceil of (N - 1) q for every quantile, unique, ascending (src/GenomeTrackQuantiles.cpp:183-195)
	std::vector<uint64_t> targets;
	for (int k = 0; k < nq; ++k) {
		const double idx = (double)(N - 1) * h_percentiles[k];
		targets.push_back((uint64_t)floor(idx));
		targets.push_back((uint64_t)ceil(idx));
	}
	std::sort(targets.begin(), targets.end());
	targets.erase(std::unique(targets.begin(), targets.end()), targets.end());
	const size_t T = targets.size();
	std::vector<uint64_t> rank(targets);     // rank of the target inside its current prefix bucket
	std::vector<uint32_t> prefix(T, 0);      // bits of its key found so far
	auto descend = [&](const unsigned long long *hist, size_t t, int nbins) { // digit whose cumulative count passes rank[t]
		unsigned long long acc = 0;
		for (int d = 0; d < nbins; ++d) {
			if (acc + hist[d] > rank[t]) {
				rank[t] -= acc;
				return (uint32_t)d;
			}
			acc += hist[d];
		}
		return (uint32_t)(nbins - 1); // unreachable: rank < bucket population
	};
	for (size_t t = 0; t < T; ++t)
		prefix[t] = descend(h.data() + 1, t, 2048);
	for (int pass = 1; pass <= 2; ++pass) {
		const int nbins = pass == 1 ? 2048 : 1024;
		std::vector<uint32_t> slots(prefix);
		std::sort(slots.begin(), slots.end());
		slots.erase(std::unique(slots.begin(), slots.end()), slots.end());
		memset(&p, 0, sizeof(p));
		p.pass = pass;
		p.nslots = (int)slots.size();
		for (size_t j = 0; j < slots.size(); ++j)
			p.prefix[j] = slots[j];
		if (count_pass(p, slots.size() * (size_t)nbins, false))
			return 1;
		for (size_t t = 0; t < T; ++t) {
			const size_t slot = std::lower_bound(slots.begin(), slots.end(), prefix[t]) - slots.begin();
			const uint32_t d = descend(h.data() + slot * (size_t)nbins, t, nbins);
			prefix[t] = (prefix[t] << (pass == 1 ? 11 : 10)) | d;
		}
	}
	std::vector<double> value(T);
	for (size_t t = 0; t < T; ++t)
		value[t] = (double)mg_host_fkey_inv(prefix[t]);
	for (int k = 0; k < nq; ++k) { // src/GenomeTrackQuantiles.cpp:222-232
		const double idx = (double)(N - 1) * h_percentiles[k];
		const uint64_t i1 = (uint64_t)floor(idx), i2 = (uint64_t)ceil(idx);
		const double w = idx - (double)i1;
		const size_t p1 = std::lower_bound(targets.begin(), targets.end(), i1) - targets.begin();
		const size_t p2 = std::lower_bound(targets.begin(), targets.end(), i2) - targets.begin();
		h_out[k] = value[p1] * (1.0 - w) + value[p2] * w;
	}
	return 0;
}

extern "C" int mg_quantiles(mg_ctx *ctx, const double *d_vals, int64_t n, int nq, const double *h_percentiles, double *h_out, int64_t *h_nkept,
							void *stream)
{
	MG_REQUIRE(ctx, n >= 0, "mg_quantiles: negative row count");
	const cudaStream_t st = mg_stream(stream);
	return mg_quantiles_multi(1, &ctx, &d_vals, &n, nq, h_percentiles, h_out, h_nkept, &st);
}

// ---- gintervals.quantiles: segmented form ----------------------------------------------------------------------------------
extern "C" int mg_quantiles_segmented(mg_ctx *ctx, const double *d_vals, int64_t nseg, const int64_t *h_seg_first, int nq,
									  const double *h_percentiles, double *h_out, int64_t *h_nkept, void *stream)
{
	MG_REQUIRE(ctx, nseg >= 0 && nseg < (int64_t)INT32_MAX && h_seg_first && nq >= 1 && nq <= MG_QS_MAXSLOTS / 2 && h_percentiles && h_out,
			   "mg_quantiles_segmented: bad arguments (1..%d quantiles)", MG_QS_MAXSLOTS / 2);
	for (int k = 0; k < nq; ++k) // gintervals_quantiles rejects these before scanning (src/GenomeTrackQuantiles.cpp:632-635)
		MG_REQUIRE(ctx, h_percentiles[k] >= 0 && h_percentiles[k] <= 1, "Percentile (%g) is not in [0, 1] range\n", h_percentiles[k]);
	for (int64_t s = 0; s < nseg; ++s)
		MG_REQUIRE(ctx, h_seg_first[s] >= 0 && h_seg_first[s] <= h_seg_first[s + 1], "mg_quantiles_segmented: segment offsets must ascend (segment %lld)",
				   (long long)s);
	if (!nseg)
		return 0;
	MG_REQUIRE(ctx, d_vals || h_seg_first[nseg] == 0, "mg_quantiles_segmented: no value column");
	cudaStream_t st = mg_stream(stream);
	MgScratch mg_scratch(st);
	// 0: one warp per segment (sort), 1: one thread block per segment (sort), 2: one thread block per segment (radix select in global memory)
	std::vector<int32_t> list[4]; // 3: radix select with the counting spread over the device (k_qlong_*)
	std::vector<int64_t> large; // 2^31 rows and more: the whole-device counting passes of mg_quantiles
	for (int64_t s = 0; s < nseg; ++s) {
		const int64_t c = h_seg_first[s + 1] - h_seg_first[s];
		if (c <= MG_QSEG_WARP_CAP)
			list[0].push_back((int32_t)s);
		else if (c <= MG_QSEG_BLOCK_CAP)
			list[1].push_back((int32_t)s);
		else if (c <= MG_QLONG_MIN)
			list[2].push_back((int32_t)s);
		else if (c < ((int64_t)1 << 31))
			list[3].push_back((int32_t)s);
		else
			large.push_back(s);
	}
	char *d_buf = nullptr;
	const size_t off_first = 0, off_pct = off_first + (size_t)(nseg + 1) * 8, off_out = off_pct + (size_t)nq * 8,
				 off_kept = off_out + (size_t)nq * nseg * 8, off_list = off_kept + (size_t)nseg * 8, total = off_list + (size_t)nseg * 4;
	MG_CUDA(ctx, mg_scratch.alloc(&d_buf, total));
	MG_CUDA(ctx, cudaMemcpyAsync(d_buf + off_first, h_seg_first, (size_t)(nseg + 1) * 8, cudaMemcpyHostToDevice, st));
	MG_CUDA(ctx, cudaMemcpyAsync(d_buf + off_pct, h_percentiles, (size_t)nq * 8, cudaMemcpyHostToDevice, st));
	QsegArgs a;
	a.vals = d_vals;
	a.seg_first = (const int64_t *)(d_buf + off_first);
	a.nseg = nseg;
	a.nq = nq;
	a.pct = (const double *)(d_buf + off_pct);
	a.out = (double *)(d_buf + off_out);
	a.nkept = (long long *)(d_buf + off_kept);
	size_t list_pos = 0;
	for (int tier = 0; tier < 3; ++tier) {
		if (list[tier].empty())
			continue;
		int32_t *d_list = (int32_t *)(d_buf + off_list) + list_pos;
		list_pos += list[tier].size();
		MG_CUDA(ctx, cudaMemcpyAsync(d_list, list[tier].data(), list[tier].size() * 4, cudaMemcpyHostToDevice, st));
		a.list = d_list;
		a.nlist = (int64_t)list[tier].size();
		if (tier == 0) {
			const unsigned grid = (unsigned)std::min<int64_t>(mg_div_up(a.nlist, 8), (int64_t)ctx->sm_count * 16);
			k_qseg<true><<<grid, 256, 8 * MG_QSEG_WARP_CAP * 4, st>>>(a);
		} else if (tier == 1) {
			const int smem = MG_QSEG_BLOCK_CAP * 4;
			MG_CUDA(ctx, cudaFuncSetAttribute(k_qseg<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
			const unsigned grid = (unsigned)std::min<int64_t>(a.nlist, (int64_t)ctx->sm_count * 3);
			k_qseg<false><<<grid, MG_QSEG_BLOCK_THREADS, smem, st>>>(a);
		} else {
			const int smem = std::min(2 * nq, (int)MG_QSEL_MAXT) * 1024; // a 256-entry histogram per distinct prefix
			MG_CUDA(ctx, cudaFuncSetAttribute(k_qseg_select, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
			const unsigned grid = (unsigned)std::min<int64_t>(a.nlist, (int64_t)ctx->sm_count * (smem <= 32768 ? 4 : 1));
			k_qseg_select<<<grid, MG_QSEL_THREADS, smem, st>>>(a);
		}
		MG_LAUNCH_CHECK(ctx);
	}
	char *d_long = nullptr;
	if (!list[3].empty()) {
		const size_t nl = list[3].size();
		std::vector<int32_t> chunk_seg;
		std::vector<int64_t> chunk_off;
		for (size_t e = 0; e < nl; ++e) {
			const int64_t c = h_seg_first[list[3][e] + 1] - h_seg_first[list[3][e]];
			for (int64_t o = 0; o < c; o += MG_QLONG_CHUNK) {
				chunk_seg.push_back((int32_t)e);
				chunk_off.push_back(o);
			}
		}
		const size_t nch = chunk_seg.size();
		const int maxS = std::min(2 * nq, (int)MG_QSEL_MAXT);
		const size_t o_state = 0, o_hist = o_state + nl * sizeof(QlongState), o_off = (o_hist + nl * (size_t)maxS * 1024 + 7) & ~(size_t)7,
					 o_cseg = o_off + nch * 8, o_list = o_cseg + nch * 4, tot = o_list + nl * 4;
		MG_CUDA(ctx, mg_scratch.alloc(&d_long, tot));
		MG_CUDA(ctx, cudaMemsetAsync(d_long, 0, o_off, st)); // states and counters
		MG_CUDA(ctx, cudaMemcpyAsync(d_long + o_off, chunk_off.data(), nch * 8, cudaMemcpyHostToDevice, st));
		MG_CUDA(ctx, cudaMemcpyAsync(d_long + o_cseg, chunk_seg.data(), nch * 4, cudaMemcpyHostToDevice, st));
		MG_CUDA(ctx, cudaMemcpyAsync(d_long + o_list, list[3].data(), nl * 4, cudaMemcpyHostToDevice, st));
		QlongArgs ql;
		ql.a = a;
		ql.a.list = (const int32_t *)(d_long + o_list);
		ql.a.nlist = (int64_t)nl;
		ql.chunk_seg = (const int32_t *)(d_long + o_cseg);
		ql.chunk_off = (const int64_t *)(d_long + o_off);
		ql.state = (QlongState *)(d_long + o_state);
		ql.hist = (uint32_t *)(d_long + o_hist);
		ql.maxS = maxS;
		const int smem = maxS * 1024;
		MG_CUDA(ctx, cudaFuncSetAttribute(k_qlong_count, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
		for (int pass = 0; pass < 4; ++pass) {
			k_qlong_count<<<(unsigned)nch, MG_QSEL_THREADS, smem, st>>>(ql, pass);
			MG_LAUNCH_CHECK(ctx);
			k_qlong_update<<<(unsigned)nl, 256, 0, st>>>(ql, pass);
			MG_LAUNCH_CHECK(ctx);
		}
	}
	std::vector<long long> kept((size_t)nseg);
	MG_CUDA(ctx, cudaMemcpyAsync(h_out, d_buf + off_out, (size_t)nq * nseg * 8, cudaMemcpyDeviceToHost, st));
	MG_CUDA(ctx, cudaMemcpyAsync(kept.data(), d_buf + off_kept, (size_t)nseg * 8, cudaMemcpyDeviceToHost, st));
	MG_CUDA(ctx, cudaStreamSynchronize(st));
	std::vector<double> q((size_t)nq);
	for (int64_t s : large) {
		int64_t nk = 0;
		const int rc = mg_quantiles(ctx, d_vals + h_seg_first[s], h_seg_first[s + 1] - h_seg_first[s], nq, h_percentiles, q.data(), &nk, stream);
		if (rc)
			return rc;
		kept[(size_t)s] = nk;
		for (int k = 0; k < nq; ++k)
			h_out[(size_t)k * nseg + s] = q[(size_t)k];
	}
	if (h_nkept)
		for (int64_t s = 0; s < nseg; ++s)
			h_nkept[s] = kept[(size_t)s];
	return 0;
}

// ---- gintervals.summary --------------------------------------------------------------------------------------------------
extern "C" int mg_summary_segmented(mg_ctx *ctx, const double *d_vals, int64_t nseg, const int64_t *h_seg_first, double *h_partials, void *stream)
{
	MG_REQUIRE(ctx, nseg >= 0 && h_seg_first && h_partials, "mg_summary_segmented: bad arguments");
	for (int64_t s = 0; s < nseg; ++s)
		MG_REQUIRE(ctx, h_seg_first[s] >= 0 && h_seg_first[s] <= h_seg_first[s + 1], "mg_summary_segmented: segment offsets must ascend (segment %lld)",
				   (long long)s);
	if (!nseg)
		return 0;
	MG_REQUIRE(ctx, d_vals || h_seg_first[nseg] == 0, "mg_summary_segmented: no value column");
	cudaStream_t st = mg_stream(stream);
	MgScratch mg_scratch(st);
	const int64_t big = (int64_t)1 << 18; // rows one warp still sums itself
	char *d_buf = nullptr;
	const size_t off_part = (size_t)(nseg + 1) * 8;
	MG_CUDA(ctx, mg_scratch.alloc(&d_buf, off_part + (size_t)nseg * 48));
	MG_CUDA(ctx, cudaMemcpyAsync(d_buf, h_seg_first, (size_t)(nseg + 1) * 8, cudaMemcpyHostToDevice, st));
	double *d_part = (double *)(d_buf + off_part);
	const unsigned grid = (unsigned)std::min<int64_t>(mg_div_up(nseg, 8), (int64_t)ctx->sm_count * 16);
	k_summary_segs<<<grid, 256, 0, st>>>(d_vals, (const int64_t *)d_buf, nseg, big, d_part);
	MG_LAUNCH_CHECK(ctx);
	for (int64_t s = 0; s < nseg; ++s) {
		const int64_t c = h_seg_first[s + 1] - h_seg_first[s];
		if (c > big && mg_summary_impl(ctx, d_vals + h_seg_first[s], c, d_part + 6 * s, stream, true)) // narrowed to float like the segments k_summary_segs takes
			return 1;
	}
	MG_CUDA(ctx, cudaMemcpyAsync(h_partials, d_part, (size_t)nseg * 48, cudaMemcpyDeviceToHost, st));
	MG_CUDA(ctx, cudaStreamSynchronize(st));
	return 0;
}

// ---- gbins.summary / gbins.quantiles ------------------------------------------------------------------------------------------
#define MG_BINS_MAX_TOTAL ((long long)1 << 24)
extern "C" int mg_bins_summary(mg_ctx *ctx, const double *d_vals, int ndim, const double *const *d_bin_cols, int64_t n, const double *h_breaks,
							   const int *nbreaks, int include_lowest, double *h_partials, void *stream)
{
	MG_REQUIRE(ctx, n >= 0 && d_bin_cols && h_breaks && nbreaks && h_partials, "mg_bins_summary: bad arguments");
	cudaStream_t st = mg_stream(stream);
	MgScratch mg_scratch(st);
	HistSpec hs;
	double *d_breaks = nullptr;
	if (mg_make_hist_spec(ctx, mg_scratch, ndim, h_breaks, nbreaks, include_lowest, hs, nullptr, st)) // validation only: nothing allocated yet
		return 1;
	MG_REQUIRE(ctx, hs.total <= MG_BINS_MAX_TOTAL, "mg_bins_summary: %lld bins exceed the limit of %lld", hs.total, MG_BINS_MAX_TOTAL);
	if (mg_make_hist_spec(ctx, mg_scratch, ndim, h_breaks, nbreaks, include_lowest, hs, &d_breaks, st))
		return 1;
	unsigned long long *d_acc = nullptr;
	MG_CUDA(ctx, mg_scratch.alloc(&d_acc, (size_t)hs.total * 96));
	double *d_part = (double *)(d_acc + 6 * hs.total);
	k_bins_init<<<(unsigned)std::min<long long>(mg_div_up(hs.total, 256), 1024), 256, 0, st>>>(d_acc, hs.total);
	MG_LAUNCH_CHECK(ctx);
	if (n > 0) {
		HistCols cols;
		for (int d = 0; d < ndim; ++d)
			cols.col[d] = d_bin_cols[d];
		const unsigned grid = (unsigned)std::min<int64_t>(mg_div_up(n, 256 * 8), (int64_t)ctx->sm_count * 8);
		if (hs.total <= MG_BINS_SMEM_MAX) {
			int rep = 1;
			while (rep < 8 && hs.total * rep * 2 <= MG_BINS_SMEM_MAX)
				rep *= 2;
			k_bins_summary<true><<<grid, 256, (size_t)hs.total * rep * 48, st>>>(d_vals, cols, n, hs, rep, d_acc);
		} else
			k_bins_summary<false><<<grid, 256, 0, st>>>(d_vals, cols, n, hs, 1, d_acc);
		MG_LAUNCH_CHECK(ctx);
	}
	k_bins_finalize<<<(unsigned)std::min<long long>(mg_div_up(hs.total, 256), 1024), 256, 0, st>>>(d_acc, hs.total, d_part);
	MG_LAUNCH_CHECK(ctx);
	MG_CUDA(ctx, cudaMemcpyAsync(h_partials, d_part, (size_t)hs.total * 48, cudaMemcpyDeviceToHost, st));
	MG_CUDA(ctx, cudaStreamSynchronize(st));
	return 0;
}

extern "C" int mg_bins_quantiles(mg_ctx *ctx, const double *d_vals, int ndim, const double *const *d_bin_cols, int64_t n, const double *h_breaks,
								 const int *nbreaks, int include_lowest, int nq, const double *h_percentiles, double *h_out, int64_t *h_nkept,
								 void *stream)
{
	MG_REQUIRE(ctx, n >= 0 && n < (int64_t)1 << 40 && d_bin_cols && h_breaks && nbreaks && h_out && h_percentiles && nq >= 1,
			   "mg_bins_quantiles: bad arguments");
	for (int k = 0; k < nq; ++k) // src/GenomeTrackQuantiles.cpp:1223
		MG_REQUIRE(ctx, h_percentiles[k] >= 0 && h_percentiles[k] <= 1, "Percentile (%g) is not in [0, 1] range\n", h_percentiles[k]);
	cudaStream_t st = mg_stream(stream);
	MgScratch mg_scratch(st);
	HistSpec hs;
	double *d_breaks = nullptr;
	if (mg_make_hist_spec(ctx, mg_scratch, ndim, h_breaks, nbreaks, include_lowest, hs, nullptr, st)) // validation only: nothing allocated yet
		return 1;
	MG_REQUIRE(ctx, hs.total <= MG_BINS_MAX_TOTAL, "mg_bins_quantiles: %lld bins exceed the limit of %lld", hs.total, MG_BINS_MAX_TOTAL);
	if (mg_make_hist_spec(ctx, mg_scratch, ndim, h_breaks, nbreaks, include_lowest, hs, &d_breaks, st))
		return 1;
	const size_t T = (size_t)hs.total;
	std::vector<int64_t> seg_first(T + 1, 0);
	char *d_buf = nullptr;
	const size_t off_idx = T * 8, off_sorted = (off_idx + (size_t)n * 4 + 7) & ~(size_t)7;
	if (n > 0) {
		MG_CUDA(ctx, mg_scratch.alloc(&d_buf, off_sorted + (size_t)n * 8));
		unsigned long long *d_counts = (unsigned long long *)d_buf;
		int32_t *d_idx = (int32_t *)(d_buf + off_idx);
		double *d_sorted = (double *)(d_buf + off_sorted);
		MG_CUDA(ctx, cudaMemsetAsync(d_counts, 0, T * 8, st));
		HistCols cols;
		for (int d = 0; d < ndim; ++d)
			cols.col[d] = d_bin_cols[d];
		const unsigned grid = (unsigned)std::min<int64_t>(mg_div_up(n, 256 * 8), (int64_t)ctx->sm_count * 8);
		const bool smem = hs.total <= MG_BINS_SCATTER_SMEM_MAX;
		if (smem)
			k_bins_index<true><<<grid, 256, 0, st>>>(d_vals, cols, n, hs, d_idx, d_counts);
		else
			k_bins_index<false><<<grid, 256, 0, st>>>(d_vals, cols, n, hs, d_idx, d_counts);
		MG_LAUNCH_CHECK(ctx);
		std::vector<unsigned long long> counts(T);
		MG_CUDA(ctx, cudaMemcpyAsync(counts.data(), d_counts, T * 8, cudaMemcpyDeviceToHost, st));
		MG_CUDA(ctx, cudaStreamSynchronize(st));
		for (size_t b = 0; b < T; ++b)
			seg_first[b + 1] = seg_first[b] + (int64_t)counts[b];
		MG_CUDA(ctx, cudaMemcpyAsync(d_counts, seg_first.data(), T * 8, cudaMemcpyHostToDevice, st)); // the bins' write cursors
		if (smem)
			k_bins_scatter<true><<<grid, 256, 0, st>>>(d_vals, d_idx, n, hs.total, d_counts, d_sorted);
		else
			k_bins_scatter<false><<<grid, 256, 0, st>>>(d_vals, d_idx, n, hs.total, d_counts, d_sorted);
		MG_LAUNCH_CHECK(ctx);
		const int rc = mg_quantiles_segmented(ctx, d_sorted, (int64_t)T, seg_first.data(), nq, h_percentiles, h_out, h_nkept, stream);
		return rc;
	}
	for (size_t j = 0; j < T * (size_t)nq; ++j)
		h_out[j] = std::numeric_limits<double>::quiet_NaN();
	if (h_nkept)
		for (size_t b = 0; b < T; ++b)
			h_nkept[b] = 0;
	return 0;
}

// ---- global.percentile* ------------------------------------------------------------------------------------------------------
extern "C" int mg_percentile_lookup(mg_ctx *ctx, double *d_col, int64_t n, const double *h_breaks, const double *h_bins, int nbreaks, void *stream)
{
	MG_REQUIRE(ctx, n >= 0 && h_breaks && h_bins, "mg_percentile_lookup: bad arguments");
	cudaStream_t st = mg_stream(stream);
	MgScratch mg_scratch(st);
	HistSpec hs;
	double *d_breaks = nullptr, *d_bins = nullptr;
	if (mg_make_hist_spec(ctx, mg_scratch, 1, h_breaks, &nbreaks, 0, hs, &d_breaks, st)) // BinFinder::init's checks (src/BinFinder.cpp:10-43)
		return 1;
	MG_CUDA(ctx, mg_scratch.alloc(&d_bins, sizeof(double) * nbreaks));
	MG_CUDA(ctx, cudaMemcpyAsync(d_bins, h_bins, sizeof(double) * nbreaks, cudaMemcpyHostToDevice, st));
	MG_CUDA(ctx, cudaStreamSynchronize(st)); // h_bins may be pageable: staged before the call returns
	if (n > 0) {
		const unsigned grid = (unsigned)std::min<int64_t>(mg_div_up(n, 256 * 4), (int64_t)ctx->sm_count * 16);
		k_percentile_lookup<<<grid, 256, 0, st>>>(d_col, n, hs, d_bins);
		MG_LAUNCH_CHECK(ctx);
	}
	return 0;
}

// ---- gscreen -----------------------------------------------------------------------------------------------------------
extern "C" int mg_screen_compare(mg_ctx *ctx, int nterms, const double *const *d_cols, const int *ops, const double *thresholds, int combine_or,
								 int64_t n, int8_t *d_flag, void *stream)
{
	MG_REQUIRE(ctx, nterms >= 1 && nterms <= MG_PRED_MAX_TERMS, "mg_screen_compare: %d terms (1..%d supported)", nterms, MG_PRED_MAX_TERMS);
	MG_REQUIRE(ctx, n >= 0 && d_flag, "mg_screen_compare: bad arguments");
	PredSpec p;
	memset(&p, 0, sizeof(p));
	p.nterms = nterms;
	p.combine_or = combine_or != 0;
	for (int k = 0; k < nterms; ++k) {
		MG_REQUIRE(ctx, d_cols[k], "mg_screen_compare: column %d is NULL", k);
		MG_REQUIRE(ctx, ops[k] >= MG_CMP_GT && ops[k] <= MG_CMP_NE, "mg_screen_compare: unknown comparison %d", ops[k]);
		p.col[k] = d_cols[k];
		p.op[k] = ops[k];
		p.thr[k] = thresholds[k];
	}
	if (!n)
		return 0;
	k_pred_flags<<<(unsigned)mg_div_up(mg_div_up(n, 4), 256), 256, 0, mg_stream(stream)>>>(p, n, d_flag);
	MG_LAUNCH_CHECK(ctx);
	return 0;
}

// the fixed-bin iterator's implicit rows take the 16-rows-per-thread kernels (the flags are then read 16 bytes at a time: the array must
// be 16-byte aligned and readable up to the next multiple of 16, which every allocation of this library is)
static bool mg_screen_fb(const mg_rows *rows, const int8_t *d_flag) { return rows->src.kind == 1 && ((uintptr_t)d_flag & 15u) == 0; }
static int64_t mg_screen_nblocks(const mg_rows *rows, const int8_t *d_flag, int64_t count)
{
	return mg_screen_fb(rows, d_flag) ? mg_div_up(count, (int64_t)MG_SCB_THREADS * MG_SCB_ROWS) : mg_div_up(count, MG_SCR_THREADS);
}
static void mg_screen_launch_count(const mg_rows *rows, int64_t first, int64_t count, const int8_t *d_flag, int64_t nblocks, long long *heads, long long *tails,
								   long long *total, cudaStream_t st)
{
	if (mg_screen_fb(rows, d_flag))
		k_screen_count_fb<<<(unsigned)nblocks, MG_SCB_THREADS, 0, st>>>(rows->src, first, count, d_flag, heads, tails);
	else
		k_screen_count<<<(unsigned)nblocks, MG_SCR_THREADS, 0, st>>>(rows->src, first, count, d_flag, heads, tails);
	k_screen_scan<<<1, 1024, 0, st>>>(heads, tails, nblocks, total);
}
static void mg_screen_launch_write(const mg_rows *rows, int64_t first, int64_t count, const int8_t *d_flag, int64_t nblocks, const long long *heads,
								   const long long *tails, int32_t *d_chrom, int64_t *d_start, int64_t *d_end, cudaStream_t st)
{
	if (mg_screen_fb(rows, d_flag))
		k_screen_write_fb<<<(unsigned)nblocks, MG_SCB_THREADS, 0, st>>>(rows->src, first, count, d_flag, heads, tails, d_chrom, d_start, d_end);
	else
		k_screen_write<<<(unsigned)nblocks, MG_SCR_THREADS, 0, st>>>(rows->src, first, count, d_flag, heads, tails, d_chrom, d_start, d_end);
}

extern "C" int mg_screen_merge(mg_ctx *ctx, const mg_rows *rows, int64_t first, int64_t count, const int8_t *d_flag, int32_t *d_chrom,
							   int64_t *d_start, int64_t *d_end, int64_t *h_nout, void *stream)
{
	MG_REQUIRE(ctx, first >= 0 && count >= 0 && first + count <= rows->src.n, "row range out of bounds");
	*h_nout = 0;
	if (!count)
		return 0;
	cudaStream_t st = mg_stream(stream);
	MgScratch mg_scratch(st);
	const int64_t nblocks = mg_screen_nblocks(rows, d_flag, count);
	long long *blk = nullptr;
	MG_CUDA(ctx, mg_scratch.alloc(&blk, sizeof(long long) * (2 * nblocks + 1)));
	long long *heads = blk, *tails = blk + nblocks, *total = blk + 2 * nblocks;
	mg_screen_launch_count(rows, first, count, d_flag, nblocks, heads, tails, total, st);
	MG_LAUNCH_CHECK(ctx);
	mg_screen_launch_write(rows, first, count, d_flag, nblocks, heads, tails, d_chrom, d_start, d_end, st);
	MG_LAUNCH_CHECK(ctx);
	long long n = 0;
	MG_CUDA(ctx, cudaMemcpyAsync(&n, total, sizeof(long long), cudaMemcpyDeviceToHost, st));
	MG_CUDA(ctx, cudaStreamSynchronize(st));
	*h_nout = n;
	return 0;
}

// The same with the intervals delivered to the host: the runs are counted first, so the device side of the result is allocated at its
// size (a 20 bp iterator over hg38 has 154 M rows and a few million runs), written, and copied into the caller's arrays.
extern "C" int mg_h_screen_merge(mg_ctx *ctx, const mg_rows *rows, int64_t first, int64_t count, const int8_t *d_flag, int32_t *h_chrom,
								 int64_t *h_start, int64_t *h_end, int64_t capacity, int64_t *h_nout, void *stream)
{
	MG_REQUIRE(ctx, first >= 0 && count >= 0 && first + count <= rows->src.n, "row range out of bounds");
	MG_REQUIRE(ctx, capacity >= 0, "mg_h_screen_merge: negative capacity");
	*h_nout = 0;
	if (!count)
		return 0;
	cudaStream_t st = mg_stream(stream);
	MgScratch mg_scratch(st);
	const int64_t nblocks = mg_screen_nblocks(rows, d_flag, count);
	long long *blk = nullptr;
	MG_CUDA(ctx, mg_scratch.alloc(&blk, sizeof(long long) * (2 * nblocks + 1)));
	long long *heads = blk, *tails = blk + nblocks, *total = blk + 2 * nblocks;
	mg_screen_launch_count(rows, first, count, d_flag, nblocks, heads, tails, total, st);
	MG_LAUNCH_CHECK(ctx);
	long long n = 0;
	MG_CUDA(ctx, cudaMemcpyAsync(&n, total, sizeof(long long), cudaMemcpyDeviceToHost, st));
	MG_CUDA(ctx, cudaStreamSynchronize(st));
	*h_nout = n;
	if (n == 0 || n > capacity) // too many for the caller's arrays: the count is the answer, nothing is copied
		return 0;
	int32_t *d_chrom = nullptr;
	int64_t *d_start = nullptr, *d_end = nullptr;
	MG_CUDA(ctx, mg_scratch.alloc(&d_chrom, sizeof(int32_t) * (size_t)n));
	MG_CUDA(ctx, mg_scratch.alloc(&d_start, sizeof(int64_t) * (size_t)n));
	MG_CUDA(ctx, mg_scratch.alloc(&d_end, sizeof(int64_t) * (size_t)n));
	mg_screen_launch_write(rows, first, count, d_flag, nblocks, heads, tails, d_chrom, d_start, d_end, st);
	MG_LAUNCH_CHECK(ctx);
	MG_CUDA(ctx, cudaMemcpyAsync(h_chrom, d_chrom, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, st));
	MG_CUDA(ctx, cudaMemcpyAsync(h_start, d_start, sizeof(int64_t) * (size_t)n, cudaMemcpyDeviceToHost, st));
	MG_CUDA(ctx, cudaMemcpyAsync(h_end, d_end, sizeof(int64_t) * (size_t)n, cudaMemcpyDeviceToHost, st));
	MG_CUDA(ctx, cudaStreamSynchronize(st));
	return 0;
}

// ---- several devices behind one host thread ------------------------------------------------------------------------------
#include "group.cuh"

#ifdef MG_DBG_BLOCKTIME
extern "C" int mg_dbg_blocktimes(unsigned long long *h_out, int n, int reset)
{
	if (h_out && cudaMemcpyFromSymbol(h_out, g_dbg_blocktime, sizeof(unsigned long long) * (size_t)n) != cudaSuccess)
		return 1;
	if (reset) {
		void *p = nullptr;
		if (cudaGetSymbolAddress(&p, g_dbg_blocktime) != cudaSuccess || cudaMemset(p, 0, sizeof(unsigned long long) << 16) != cudaSuccess)
			return 1;
	}
	return 0;
}
#endif
