// mishagpu.cu -- the C ABI declared in include/mishagpu.h: context, staging, row sources and kernel dispatch.
// One translation unit; kernels live in the .cuh files next to it. Built by misha_b200/build.py with
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 --shared ...  ->  misha_b200/libmishagpu.so
#include "common.cuh"
#include "rows.cuh"
#include "dense.cuh"
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
	for (int i = 0; i < 3; ++i) {
		if (ctx->pipe_dev[i])
			cudaFree(ctx->pipe_dev[i]);
		if (ctx->pipe_stream[i])
			cudaStreamDestroy(ctx->pipe_stream[i]);
	}
	cudaFree(ctx->d_chrom_sizes);
	delete ctx;
}

extern "C" const char *mg_last_error(const mg_ctx *ctx) { return ctx ? ctx->err.c_str() : g_mg_init_err.c_str(); }

extern "C" int mg_sync(mg_ctx *ctx, void *stream)
{
	MG_CUDA(ctx, cudaStreamSynchronize(mg_stream(stream)));
	return 0;
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

template <typename T>
static int mg_alloc_tracked(mg_ctx *ctx, std::vector<void *> &allocs, int64_t &bytes, int64_t count, T **out)
{
	void *p = nullptr;
	size_t sz = sizeof(T) * (size_t)(count > 0 ? count : 1);
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
	t->h_table.assign(ctx->nchroms, DenseChrom{nullptr, nullptr, 0, 0.});
	if (mg_alloc_tracked(ctx, t->allocs, t->bytes, ctx->nchroms, &t->d_table)) {
		delete t;
		return 1;
	}
	cudaError_t e = cudaMemcpy(t->d_table, t->h_table.data(), sizeof(DenseChrom) * ctx->nchroms, cudaMemcpyHostToDevice);
	if (e != cudaSuccess) {
		mg_dense_free(ctx, t);
		return mg_fail(ctx, "mg_dense_create: %s", cudaGetErrorString(e));
	}
	*out = t;
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
	float *vals = nullptr;
	Pref *pref = nullptr;
	const int64_t padded = (nbins + 3) / 4 * 4 + 4; // float4-safe tail, NaN filled
	if (mg_alloc_tracked(ctx, t->allocs, t->bytes, padded, &vals) || mg_alloc_tracked(ctx, t->allocs, t->bytes, nbins + 1, &pref))
		return 1;
	MG_CUDA(ctx, cudaMemset(vals, 0xff, sizeof(float) * padded)); // 0xffffffff is a NaN
	if (nbins)
		MG_CUDA(ctx, cudaMemcpy(vals, h_bins, sizeof(float) * nbins, cudaMemcpyHostToDevice));
	const int64_t ntiles = mg_div_up(nbins + 1, MG_ST_TILE); // +1: the closing entry pref[nbins]
	double *tile_sum = nullptr, *tile_abs = nullptr, *abs_total = nullptr;
	long long *tile_cnt = nullptr;
	MG_CUDA(ctx, cudaMalloc(&tile_sum, sizeof(double) * ntiles));
	MG_CUDA(ctx, cudaMalloc(&tile_abs, sizeof(double) * ntiles));
	MG_CUDA(ctx, cudaMalloc(&tile_cnt, sizeof(long long) * ntiles));
	MG_CUDA(ctx, cudaMalloc(&abs_total, sizeof(double)));
	k_dense_clean_reduce<<<(unsigned)ntiles, MG_ST_THREADS>>>(vals, nbins, tile_sum, tile_cnt, tile_abs);
	MG_LAUNCH_CHECK(ctx);
	k_dense_tile_scan<<<1, 1024>>>(tile_sum, tile_cnt, tile_abs, ntiles, abs_total);
	MG_LAUNCH_CHECK(ctx);
	k_dense_prefix_write<<<(unsigned)ntiles, MG_ST_THREADS>>>(vals, nbins, tile_sum, tile_cnt, pref);
	MG_LAUNCH_CHECK(ctx);
	DenseChrom dc{vals, pref, nbins, 0.};
	MG_CUDA(ctx, cudaMemcpy(&dc.abs_sum, abs_total, sizeof(double), cudaMemcpyDeviceToHost));
	t->h_table[chromid] = dc;
	MG_CUDA(ctx, cudaMemcpy(t->d_table + chromid, &dc, sizeof(DenseChrom), cudaMemcpyHostToDevice));
	cudaFree(tile_sum);
	cudaFree(tile_abs);
	cudaFree(tile_cnt);
	cudaFree(abs_total);
	return 0;
}

extern "C" void mg_dense_free(mg_ctx *ctx, mg_dense *t)
{
	if (!t)
		return;
	for (void *p : t->allocs)
		cudaFree(p);
	delete t;
}

extern "C" int64_t mg_dense_bytes(const mg_dense *t) { return t->bytes; }

// ---- rows ------------------------------------------------------------------------------------------------
extern "C" int mg_rows_upload(mg_ctx *ctx, int64_t n, const int32_t *h_chrom, const int64_t *h_start, const int64_t *h_end, void *stream,
							  mg_rows **out)
{
	MG_REQUIRE(ctx, n >= 0, "mg_rows_upload: negative row count");
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
	const size_t total = ns * 8 * 2 + (ns + 1) * 8 + ns * 4;
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
	r->src.schrom = (const int32_t *)(base + ns * 16 + (ns + 1) * 8);
	if (nscope) {
		if ((e = cudaMemcpy((void *)r->src.sstart, h_start, (size_t)nscope * 8, cudaMemcpyHostToDevice)) != cudaSuccess ||
			(e = cudaMemcpy((void *)r->src.send, h_end, (size_t)nscope * 8, cudaMemcpyHostToDevice)) != cudaSuccess ||
			(e = cudaMemcpy((void *)r->src.schrom, h_chrom, (size_t)nscope * 4, cudaMemcpyHostToDevice)) != cudaSuccess) {
			mg_rows_free(ctx, r);
			return mg_fail(ctx, "mg_rows_fixedbin: %s", cudaGetErrorString(e));
		}
	}
	if ((e = cudaMemcpy((void *)r->src.srow0, row0.data(), ((size_t)nscope + 1) * 8, cudaMemcpyHostToDevice)) != cudaSuccess) {
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
	need_sweep = false;
	for (int k = 0; k < nfuncs; ++k) {
		const int f = funcs[k];
		MG_REQUIRE(ctx, f >= 0 && f < MG_NUM_FUNCS, "unknown function id %d", f);
		MG_REQUIRE(ctx, d_out[k], "output column %d is NULL", k);
		if (f == MG_QUANTILE) {
			MG_REQUIRE(ctx, q.nq < MG_MAX_Q, "at most %d quantile columns per call", MG_MAX_Q);
			MG_REQUIRE(ctx, params, "quantile needs a percentile parameter");
			q.q_pct[q.nq] = params[k];
			q.q_out[q.nq] = d_out[k];
			q.nq++;
			need_sweep = true;
		} else {
			MG_REQUIRE(ctx, !q.out[f], "function %d requested twice in one call", f);
			q.out[f] = d_out[k];
			if (f == MG_MIN || f == MG_MAX || f == MG_STDDEV)
				need_sweep = true;
		}
	}
	return 0;
}

static int mg_launch_dense(mg_ctx *ctx, const DenseQuery &q, bool need_sweep, cudaStream_t st)
{
	if (!q.count)
		return 0;
	if (!need_sweep) {
		k_dense_prefix<<<(unsigned)mg_div_up(q.count, 256), 256, 0, st>>>(q);
	} else {
		const int64_t blocks_needed = mg_div_up(q.count, MG_SW_WARPS);
		const int64_t cap = (int64_t)ctx->sm_count * 32; // grid-stride beyond that
		k_dense_sweep<<<(unsigned)(blocks_needed < cap ? blocks_needed : cap), MG_SW_WARPS * 32, 0, st>>>(q);
	}
	MG_LAUNCH_CHECK(ctx);
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
	for (int i = 0; i < 3; ++i) {
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

extern "C" int mg_h_dense_window(mg_ctx *ctx, const mg_dense *t, int64_t n, const int32_t *h_chrom, const int64_t *h_start,
								 const int64_t *h_end, int64_t sshift, int64_t eshift, int nfuncs, const int *funcs, const double *params,
								 double *const *h_out)
{
	MG_REQUIRE(ctx, n >= 0 && nfuncs > 0 && nfuncs <= 32, "mg_h_dense_window: bad arguments");
	const int64_t CHUNK = 1 << 20; // rows per pipeline slot
	const int64_t slot_bytes = CHUNK * (20 + 8 * (int64_t)nfuncs);
	if (mg_pipe_prepare(ctx, slot_bytes))
		return 1;
	int64_t done = 0;
	for (int64_t c = 0; done < n; ++c, done += CHUNK) {
		const int slot = (int)(c % 3);
		cudaStream_t st = ctx->pipe_stream[slot];
		const int64_t m = n - done < CHUNK ? n - done : CHUNK;
		char *base = (char *)ctx->pipe_dev[slot];
		int64_t *d_start = (int64_t *)base, *d_end = (int64_t *)(base + CHUNK * 8);
		int32_t *d_chrom = (int32_t *)(base + CHUNK * 16);
		double *d_cols[32];
		for (int k = 0; k < nfuncs; ++k)
			d_cols[k] = (double *)(base + CHUNK * 20 + (int64_t)k * CHUNK * 8);
		MG_CUDA(ctx, cudaMemcpyAsync(d_start, h_start + done, (size_t)m * 8, cudaMemcpyHostToDevice, st));
		MG_CUDA(ctx, cudaMemcpyAsync(d_end, h_end + done, (size_t)m * 8, cudaMemcpyHostToDevice, st));
		MG_CUDA(ctx, cudaMemcpyAsync(d_chrom, h_chrom + done, (size_t)m * 4, cudaMemcpyHostToDevice, st));
		mg_rows rows;
		rows.src.kind = 0;
		rows.src.n = m;
		rows.src.chrom = d_chrom;
		rows.src.start = d_start;
		rows.src.end = d_end;
		DenseQuery q;
		bool need_sweep;
		if (mg_fill_dense_query(ctx, t, &rows, 0, m, sshift, eshift, nfuncs, funcs, params, d_cols, q, need_sweep) ||
			mg_launch_dense(ctx, q, need_sweep, st))
			return 1;
		for (int k = 0; k < nfuncs; ++k)
			MG_CUDA(ctx, cudaMemcpyAsync(h_out[k] + done, d_cols[k], (size_t)m * 8, cudaMemcpyDeviceToHost, st));
	}
	for (int i = 0; i < 3; ++i)
		MG_CUDA(ctx, cudaStreamSynchronize(ctx->pipe_stream[i]));
	return 0;
}
