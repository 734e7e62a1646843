// common.cuh -- context, error plumbing and small device helpers shared by every kernel file.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <map>
#include <new>
#include <algorithm>
#include <limits>
#include <chrono>
#include <cstdlib>

#include "../../include/mishagpu.h"

#define MG_SM_COUNT_B200 148

#define MG_PIPE_MAX_SLOTS 8
struct mg_ctx {
	int device = 0;
	int sm_count = MG_SM_COUNT_B200;
	int cc = 0;
	int64_t hbm_bytes = 0;
	int nchroms = 0;
	std::vector<int64_t> chrom_sizes;
	int64_t *d_chrom_sizes = nullptr;
	int64_t launches = 0;
	int64_t pipe_chunk_rows = 3 << 17; // rows per slot of the host-buffer pipeline (C3 end to end on four B200 boxes: 5.3 - 6.5 ms at 2^19, 5.5 - 5.7 at 3 * 2^17, 5.6 - 5.8 at 2^18, 5.7 at 2^20, 7.6 at 2^17; the copies alone 4.7 - 4.8)
	bool zero_copy = false; // host-buffer entry points: kernels read rows / write columns straight in pinned host memory.
	                        // Measured SLOWER than the 3-slot DMA pipeline on B200 (5.9 vs 5.3 ms for C3), hence off by default.
	bool pwm_strict = false; // PWM strand / anchor log-sum-exp with double-precision exp / log (correctly rounded float results)
	bool prefetch = true;    // row-query kernels prefetch their block record / window lines (off: ablation)
	int merge_rows = 0;      // 8: merge kernels at 8 rows per thread whatever the row source (0: chosen per launch)
	bool hist_red = true;    // gdist over a dense track: 16-bit counters of warp pairs packed in 32-bit words, one red.shared.add per value
	bool cov_raster = true;  // coverage over a regular grid of rows: records rasterised into the block's rows (off: per-row cursors; ablation)
	bool stage_records = true; // k_dense_rowquery stages block records in shared memory (option "stage_slots" = 0 turns it off)
	bool block_wm = true;    // quantiles of windows <= MG_BW_S bins walk the block-local wavelet matrices (off: ablation / tests of the
	                         // chromosome-wide matrix)
	int tile_kernel = 3;     // dense window functions: 0 the per-row index kernels (k_dense_prefix / k_dense_rowquery) only; 1 k_dense_tile
	                         // (thread block per tile of rows, everything built in shared memory, sums to the reference's bits) for
	                         // min / max / quantile sets, 2 also for avg / sum alone; 3 (default) k_dense_thin (staged bins and block
	                         // records, sums / extrema from the chromosome's structures) when two or more function groups share a
	                         // scan or a quantile is asked for, the per-row kernels for sums alone or extrema alone (measured, C3 shape:
	                         // avg 0.190 vs 0.186 ms, max 0.160 vs 0.187, quantile 0.322 vs 0.292); 4 k_dense_thin for everything
	int64_t hbm_budget = 0;  // bytes the dense tracks' index structures + cores may hold on this device (0: 85 % of the device's memory)
	int64_t use_tick = 0;    // advances with every query: least-recently-used index groups go first when the budget is short
	std::vector<struct mg_dense *> dense_tracks; // every live dense track of this context (budget accounting)
	bool inline_table = true; // the chromosome table travels in the kernel parameters of the dense tile kernels (dense.cuh: DenseQuery::inl_*)
	uint32_t pipe_attr = 0;  // k_dense_pipe instantiations whose shared-memory attributes have been set
	uint32_t tile_attr = 0;  // k_dense_tile instantiations whose dynamic shared-memory limit has been raised on this device
	bool force_sweep = false; // tests / ablation: route min/max/quantile through the warp-per-row sweep kernel
	std::string err;
	// exact float thresholds of the last 1-D break set (mg_dense_hist streaming path)
	std::vector<double> thr_breaks;
	std::vector<float> thr_values;
	int thr_include_lowest = -1;
	std::map<void *, void *> stream_scratch; // per-stream reduction scratch (mishagpu.cu: mg_stream_scratch)
	// internal streams + staging buffers of the host-buffer entry points (lazily created)
	cudaStream_t pipe_stream[MG_PIPE_MAX_SLOTS] = {};
	void *pipe_dev[MG_PIPE_MAX_SLOTS] = {};
	int64_t pipe_dev_bytes[MG_PIPE_MAX_SLOTS] = {};
	int pipe_slots = 3; // device slots (one stream each) the chunks of a host batch rotate through (option pipe_slots)
	unsigned long long *d_first_bad = nullptr; // first row of a host batch whose chromosome id was out of range (k_rows_sanitize); ~0: none
	uint8_t *d_served = nullptr; // member of a group: served[chromid] != 0 for the chromosomes this device owns (null: every chromosome)
	int64_t pipe_next = 0; // chunks enqueued so far: the slot of the next one is pipe_next % pipe_slots
};

extern thread_local std::string g_mg_init_err;

static inline int mg_fail(mg_ctx *ctx, const char *fmt, ...)
{
	char buf[1024];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	if (ctx)
		ctx->err = buf;
	else
		g_mg_init_err = buf;
	return 1;
}

#define MG_CUDA(ctx, call)                                                                                             \
	do {                                                                                                               \
		cudaError_t e__ = (call);                                                                                      \
		if (e__ != cudaSuccess)                                                                                        \
			return mg_fail(ctx, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__);          \
	} while (0)

#define MG_LAUNCH_CHECK(ctx)                                                                                           \
	do {                                                                                                               \
		(ctx)->launches++;                                                                                             \
		cudaError_t e__ = cudaGetLastError();                                                                          \
		if (e__ != cudaSuccess)                                                                                        \
			return mg_fail(ctx, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__);      \
	} while (0)

#define MG_REQUIRE(ctx, cond, ...)                                                                                     \
	do {                                                                                                               \
		if (!(cond))                                                                                                   \
			return mg_fail(ctx, __VA_ARGS__);                                                                          \
	} while (0)

static inline cudaStream_t mg_stream(void *s) { return (cudaStream_t)s; }

static inline int64_t mg_div_up(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device helpers ------------------------------------------------------------------------------------------
__device__ __forceinline__ double mg_nan() { return __longlong_as_double(0x7ff8000000000000LL); }
__device__ __forceinline__ float mg_nanf() { return __int_as_float(0x7fc00000); }

// float member widened to double, as the reference's last_*() accessors feed var[idx] (src/TrackVarProcessor.cpp:113-173)
__device__ __forceinline__ double mg_f2d(double x) { return (double)__double2float_rn(x); }

// streaming loads/stores: data touched once should not displace the reused working set in L1
__device__ __forceinline__ float mg_ldg_f(const float *p) { return __ldg(p); }

__device__ __forceinline__ void mg_st_stream(double *p, double v) { __stcs(p, v); }

__device__ __forceinline__ void mg_prefetch(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void mg_prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Window transform, src/TrackExpressionVars.h:395-402. Returns false when out of range.
__device__ __forceinline__ bool mg_window(int64_t start, int64_t end, int64_t sshift, int64_t eshift, int64_t chromsize, int64_t &ws,
										  int64_t &we)
{
	ws = start + sshift;
	we = end + eshift;
	if (ws < 0)
		ws = 0;
	if (we > chromsize)
		we = chromsize;
	return ws < we;
}

// floor(a / b), ceil(a / b) for 0 <= a, 0 < b < 2^32 with a 32-bit fast path (genomic coordinates fit in 31 bits
// for every genome misha ships; the 64-bit path keeps it general). Identical to the reference's
// (int64_t)(a / b) and (int64_t)ceil(a / (double)b) for a < 2^53.
__device__ __forceinline__ int64_t mg_floordiv(int64_t a, uint32_t b)
{
	if ((uint64_t)a <= 0xffffffffull)
		return (int64_t)((uint32_t)a / b);
	return a / (int64_t)b;
}
__device__ __forceinline__ int64_t mg_ceildiv(int64_t a, uint32_t b)
{
	if ((uint64_t)a <= 0xffffffffull) {
		uint32_t q = (uint32_t)a / b;
		return (int64_t)q + ((uint32_t)a - q * b != 0);
	}
	int64_t q = a / (int64_t)b;
	return q + (a - q * (int64_t)b != 0);
}

__device__ __forceinline__ double mg_shfl_xor_d(double v, int m)
{
	return __shfl_xor_sync(0xffffffffu, v, m);
}
__device__ __forceinline__ double mg_warp_sum_d(double v)
{
#pragma unroll
	for (int m = 16; m > 0; m >>= 1)
		v += __shfl_xor_sync(0xffffffffu, v, m);
	return v;
}
