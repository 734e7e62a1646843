// stats.cuh -- the consumers of a scan: gsummary moments, gdist histograms, gscreen run merging.
//
// gsummary : IntervalSummary's six doubles {num_bins, num_non_nan, sum, min, max, sum_sq}
//            (src/GenomeTrackSummary.cpp:22-69). Fixed-shape two-stage reduction (grid partials -> one block), so a
//            given input always reduces in the same order; results are accumulated into a 6-double device record
//            that the host reduces across GPUs (NCCL sum / min / max) and finalises with mg_summary_finalize.
// gdist    : BinFinder::val2bin (src/BinFinder.h:55-108) in double, BinsManager::vals2idx (src/BinsManager.h:53-72).
//            Small histograms use per-lane private columns in shared memory (no atomics, no bank conflicts);
//            larger ones shared-memory or global atomics.
// gscreen  : run-length merge of consecutive TRUE rows (src/GenomeTrackScreener.cpp:195-220).
#pragma once
#include "common.cuh"
#include "rows.cuh"
#include "dense.cuh"

struct Moments {
	double n, nn, sum, mn, mx, ssq;
};

__device__ __forceinline__ void mom_init(Moments &m)
{
	m.n = 0;
	m.nn = 0;
	m.sum = 0;
	m.mn = 1.7976931348623157e308;
	m.mx = -1.7976931348623157e308;
	m.ssq = 0;
}
__device__ __forceinline__ void mom_add(Moments &m, double v)
{
	m.n += 1;
	if (!isnan(v)) {
		m.nn += 1;
		m.sum += v;
		m.mn = fmin(m.mn, v);
		m.mx = fmax(m.mx, v);
		m.ssq += __dmul_rn(v, v);
	}
}
__device__ __forceinline__ void mom_merge(Moments &a, const Moments &b)
{
	a.n += b.n;
	a.nn += b.nn;
	a.sum += b.sum;
	a.mn = fmin(a.mn, b.mn);
	a.mx = fmax(a.mx, b.mx);
	a.ssq += b.ssq;
}
__device__ __forceinline__ Moments mom_shfl_xor(const Moments &m, int mask)
{
	Moments o;
	o.n = __shfl_xor_sync(0xffffffffu, m.n, mask);
	o.nn = __shfl_xor_sync(0xffffffffu, m.nn, mask);
	o.sum = __shfl_xor_sync(0xffffffffu, m.sum, mask);
	o.mn = __shfl_xor_sync(0xffffffffu, m.mn, mask);
	o.mx = __shfl_xor_sync(0xffffffffu, m.mx, mask);
	o.ssq = __shfl_xor_sync(0xffffffffu, m.ssq, mask);
	return o;
}

#define MG_SUM_THREADS 256
#define MG_SUM_MAX_BLOCKS 1184 // 148 SMs x 8

// block-level merge; result valid in thread 0
__device__ __forceinline__ void mom_block_reduce(Moments &m, Moments *smem /* >= 8 */)
{
#pragma unroll
	for (int k = 16; k > 0; k >>= 1) {
		Moments o = mom_shfl_xor(m, k);
		mom_merge(m, o);
	}
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (lane == 0)
		smem[warp] = m;
	__syncthreads();
	if (threadIdx.x == 0) {
		for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
			mom_merge(m, smem[w]);
	}
}

// NARROW: every value goes through float first, as the reference's per-interval summary reads it (`float val`,
// src/GenomeTrackSummary.cpp:314) -- what k_summary_segs does for the segments it handles itself
template <bool NARROW>
__global__ void __launch_bounds__(MG_SUM_THREADS) k_summary_values(const double *__restrict__ vals, int64_t n, Moments *block_out)
{
	__shared__ Moments sm[MG_SUM_THREADS / 32];
	Moments m;
	mom_init(m);
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
		mom_add(m, NARROW ? (double)(float)__ldcs(vals + i) : __ldcs(vals + i));
	mom_block_reduce(m, sm);
	if (threadIdx.x == 0)
		block_out[blockIdx.x] = m;
}

// one block: merge the per-block partials in index order and accumulate into the caller's 6-double record
__global__ void __launch_bounds__(MG_SUM_THREADS) k_summary_final(const Moments *block_in, int nblocks, double *partial)
{
	__shared__ Moments sm[MG_SUM_THREADS / 32];
	Moments m;
	mom_init(m);
	for (int i = threadIdx.x; i < nblocks; i += blockDim.x)
		mom_merge(m, block_in[i]);
	mom_block_reduce(m, sm);
	if (threadIdx.x == 0) {
		partial[0] += m.n;
		partial[1] += m.nn;
		partial[2] += m.sum;
		partial[3] = fmin(partial[3], m.mn);
		partial[4] = fmax(partial[4], m.mx);
		partial[5] += m.ssq;
	}
}

__global__ void k_summary_init(double *partial)
{
	partial[0] = 0;
	partial[1] = 0;
	partial[2] = 0;
	partial[3] = 1.7976931348623157e308;
	partial[4] = -1.7976931348623157e308;
	partial[5] = 0;
}

// ---- values of one prefix-able dense function for a row (avg / sum / nearest / size / exists), as k_dense_prefix ------
__device__ __forceinline__ double mg_dense_row_value(const DenseQuery &q, int64_t row, int func)
{
	DenseWin w;
	if (!mg_dense_win(q, row, w))
		return mg_nan();
	RowSum r;
	r.avg = r.sum = mg_nan();
	r.size = 0;
	r.exists = 0;
	if (w.ebin > w.sbin) {
		WinBasic wb;
		mg_win_basic<true, false>(w.ch, w.sbin, w.ebin, wb);
		mg_row_sum(w.ch, wb, w.sbin, w.ebin, r);
	}
	if (func == MG_SIZE)
		return r.size;
	if (func == MG_EXISTS)
		return r.exists;
	return func == MG_SUM ? r.sum : r.avg;
}

__global__ void __launch_bounds__(MG_SUM_THREADS) k_dense_summary(const DenseQuery q, int func, Moments *block_out)
{
	__shared__ Moments sm[MG_SUM_THREADS / 32];
	Moments m;
	mom_init(m);
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < q.count; i += stride)
		mom_add(m, mg_dense_row_value(q, q.first + i, func));
	mom_block_reduce(m, sm);
	if (threadIdx.x == 0)
		block_out[blockIdx.x] = m;
}

// ---- streaming fast path: fixed-bin iterator at the track's own bin size, no shifts => row value == bin value ----------
// Grid-stride over the implicit rows: a warp reads 128 contiguous bytes of the track per request. Each thread keeps
// the scope interval it is in and only steps forward (rows grow monotonically along its stride).
template <typename F>
__device__ __forceinline__ void mg_stream_rows(const RowSrc &rs, const DenseChrom *table, int64_t first, int64_t count, F &&f)
{
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	int64_t k = -1, row0 = 0, row1 = 0, bin0 = 0, nbins = 0;
	const float *vals = nullptr;
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
		const int64_t r = first + i;
		if (k < 0 || r >= row1) {
			k = k < 0 ? mg_scope_of_row(rs, r) : k + 1;
			while (r >= __ldg(rs.srow0 + k + 1))
				++k;
			row0 = __ldg(rs.srow0 + k);
			row1 = __ldg(rs.srow0 + k + 1);
			const DenseChrom *ch = table + __ldg(rs.schrom + k);
			vals = ch->vals;
			nbins = ch->nbins;
			bin0 = __ldg(rs.sbin0 + k);
		}
		const int64_t b = bin0 + (r - row0);
		f((double)(b < nbins ? __ldcs(vals + b) : mg_nanf()));
	}
}

__global__ void __launch_bounds__(MG_SUM_THREADS) k_dense_stream_summary(const RowSrc rs, const DenseChrom *table, int64_t first,
																		  int64_t count, Moments *block_out)
{
	__shared__ Moments sm[MG_SUM_THREADS / 32];
	Moments m;
	mom_init(m);
	mg_stream_rows(rs, table, first, count, [&](double v) { mom_add(m, v); });
	mom_block_reduce(m, sm);
	if (threadIdx.x == 0)
		block_out[blockIdx.x] = m;
}

// ---- contiguous-segment streaming kernels (fixed-bin iterator at track resolution over few scope intervals) ---------------
// One launch per scope interval: a pure stream over vals[b0, b0 + n) with aligned 16-byte loads, 4 independent loads in
// flight per thread. FP64 work per bin is one DADD (sum) and one DFMA (sum of squares: the product of two floats is exact
// in double, so fma(v, v, acc) rounds exactly like acc + v * v); counts are integers, min / max stay in float.
struct MomF {
	unsigned long long nn;
	double sum, ssq;
	float mn, mx;
};
__device__ __forceinline__ void momf_add(MomF &m, float v)
{
	if (!isnan(v)) {
		const double d = (double)v;
		m.nn++;
		m.sum += d;
		m.ssq = fma(d, d, m.ssq);
		m.mn = fminf(m.mn, v);
		m.mx = fmaxf(m.mx, v);
	}
}

// All segments of a scan in ONE launch: the segments are cut into tiles of MG_SEG_TILE floats, aligned to 16 bytes
// relative to each segment's own aligned-down base, and the tiles are dealt round-robin to a persistent grid. Interior
// tiles are four unpredicated 16-byte loads per thread; only a segment's first / last tile masks elements.
#define MG_MAX_STREAM_SEGS 64
#define MG_SEG_TILE (MG_SUM_THREADS * 16)
struct SegTable {
	int nseg;
	const float *p[MG_MAX_STREAM_SEGS]; // first element of the segment
	int64_t n[MG_MAX_STREAM_SEGS];
	int64_t tile0[MG_MAX_STREAM_SEGS + 1]; // exclusive prefix of tiles per segment
};

template <typename F>
__device__ __forceinline__ void mg_stream_segtable(const SegTable &st, F &&f)
{
	const int64_t ntiles = st.tile0[st.nseg];
	int s = 0;
	for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
		while (tile >= st.tile0[s + 1])
			++s;
		const float *p = st.p[s];
		const int64_t mis = (int64_t)((reinterpret_cast<uintptr_t>(p) >> 2) & 3); // floats between the aligned-down base and p
		const int64_t lo = mis, hi = mis + st.n[s];                                 // valid element range relative to that base
		const int64_t e0 = (tile - st.tile0[s]) * MG_SEG_TILE;
		const float4 *q = reinterpret_cast<const float4 *>(p - mis + e0) + threadIdx.x;
		if (e0 >= lo && e0 + MG_SEG_TILE <= hi) {
			const float4 a = __ldcs(q), b = __ldcs(q + MG_SUM_THREADS), c = __ldcs(q + 2 * MG_SUM_THREADS), d = __ldcs(q + 3 * MG_SUM_THREADS);
			f(a.x); f(a.y); f(a.z); f(a.w);
			f(b.x); f(b.y); f(b.z); f(b.w);
			f(c.x); f(c.y); f(c.z); f(c.w);
			f(d.x); f(d.y); f(d.z); f(d.w);
		} else {
#pragma unroll
			for (int k = 0; k < 4; ++k) {
				const int64_t i0 = e0 + ((int64_t)threadIdx.x + k * MG_SUM_THREADS) * 4;
				if (i0 + 3 < lo || i0 >= hi)
					continue;
				const float4 a = __ldcs(q + k * MG_SUM_THREADS); // stays inside the staged array: it is padded past nbins
				if (i0 >= lo && i0 < hi) f(a.x);
				if (i0 + 1 >= lo && i0 + 1 < hi) f(a.y);
				if (i0 + 2 >= lo && i0 + 2 < hi) f(a.z);
				if (i0 + 3 >= lo && i0 + 3 < hi) f(a.w);
			}
		}
	}
}

// The last block to finish (ticket counter) merges the per-block partials in index order and accumulates them into the
// caller's 6-double record: one launch per gsummary call, and the reduction order is fixed whatever the block schedule.
__global__ void __launch_bounds__(MG_SUM_THREADS) k_segs_summary(const SegTable st, int64_t nrows, Moments *block_out, unsigned *ticket,
																  double *partial)
{
	__shared__ Moments sm[MG_SUM_THREADS / 32];
	__shared__ bool is_last;
	MomF mf{0ull, 0., 0., __int_as_float(0x7f800000), __int_as_float(0xff800000)};
	mg_stream_segtable(st, [&](float v) { momf_add(mf, v); });
	Moments m;
	m.n = blockIdx.x == 0 && threadIdx.x == 0 ? (double)nrows : 0.;
	m.nn = (double)mf.nn;
	m.sum = mf.sum;
	m.ssq = mf.ssq;
	m.mn = mf.nn ? (double)mf.mn : 1.7976931348623157e308;
	m.mx = mf.nn ? (double)mf.mx : -1.7976931348623157e308;
	mom_block_reduce(m, sm);
	if (threadIdx.x == 0) {
		block_out[blockIdx.x] = m;
		__threadfence();
		is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
	}
	__syncthreads();
	if (!is_last)
		return;
	__threadfence();
	mom_init(m);
	for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x)
	{
		const double *p = reinterpret_cast<const double *>(block_out + i); // written by other blocks: read through L2
		Moments o;
		o.n = __ldcg(p);
		o.nn = __ldcg(p + 1);
		o.sum = __ldcg(p + 2);
		o.mn = __ldcg(p + 3);
		o.mx = __ldcg(p + 4);
		o.ssq = __ldcg(p + 5);
		mom_merge(m, o);
	}
	mom_block_reduce(m, sm);
	if (threadIdx.x == 0) {
		partial[0] += m.n;
		partial[1] += m.nn;
		partial[2] += m.sum;
		partial[3] = fmin(partial[3], m.mn);
		partial[4] = fmax(partial[4], m.mx);
		partial[5] += m.ssq;
		*ticket = 0; // ready for the next launch on this stream
	}
}

// ---- histogram -----------------------------------------------------------------------------------------------------------
#define MG_HIST_MAX_DIMS 8
#define MG_HIST_MAX_BREAKS 1024 // per call, all dimensions together (constant-bank resident)

struct HistSpec {
	int ndim;
	int include_lowest;
	int nbreaks[MG_HIST_MAX_DIMS];
	int boff[MG_HIST_MAX_DIMS];      // offset of the dimension's breaks in `breaks`
	long long mult[MG_HIST_MAX_DIMS]; // first dimension fastest
	double binsize[MG_HIST_MAX_DIMS]; // BinFinder::m_binsize (0 => binary search)
	long long total;
	const double *breaks; // device copy, concatenated
};

// BinFinder::val2bin, right-closed intervals (src/BinFinder.h:55-81)
__device__ __forceinline__ int mg_val2bin(const double *br, int nbreaks, double binsize, int include_lowest, double val)
{
	const double b0 = br[0];
	if (include_lowest && val == b0)
		return 0;
	if (isnan(val) || val <= b0 || val > br[nbreaks - 1])
		return -1;
	const int numbins = nbreaks - 1;
	if (binsize != 0.) {
		const int b = (int)ceil(__ddiv_rn(__dsub_rn(val, b0), binsize)) - 1;
		return b < numbins - 1 ? b : numbins - 1;
	}
	unsigned s = 0, e = (unsigned)numbins;
	while (e - s > 1) {
		const unsigned mid = (s + e) >> 1;
		if (val <= br[mid])
			e = mid;
		else
			s = mid;
	}
	return (int)s;
}

#define MG_HIST_THREADS 256
#define MG_HIST_PRIVATE_MAX 160 // bins: 160 x 32 lanes x 4 B = 20 KB per warp, 8 warps = 160 KB dynamic smem

// 1-D histogram with per-lane private columns: hist[warp][bin][lane]; no atomics in the hot loop.
template <typename Gen>
__device__ __forceinline__ void mg_hist1d_private(const HistSpec &hs, uint32_t *smem, unsigned long long *counts, Gen &&gen)
{
	const int nb = hs.nbreaks[0] - 1;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
	double *sbr = reinterpret_cast<double *>(smem);                       // breaks first (8-byte aligned)
	uint32_t *h = smem + 2 * hs.nbreaks[0] + warp * nb * 32;              // then the per-warp columns
	for (int t = threadIdx.x; t < hs.nbreaks[0]; t += blockDim.x)
		sbr[t] = hs.breaks[t];
	for (int t = lane; t < nb * 32; t += 32)
		h[t] = 0;
	__syncthreads();
	const double binsize = hs.binsize[0];
	gen([&](double v) {
		const int b = mg_val2bin(sbr, nb + 1, binsize, hs.include_lowest, v);
		if (b >= 0)
			h[b * 32 + lane]++;
	});
	__syncthreads();
	uint32_t *hall = smem + 2 * hs.nbreaks[0];
	for (int b = warp; b < nb; b += nwarps) { // one warp sums the 32 x nwarps replicas of a bin
		unsigned long long c = 0;
		for (int w = 0; w < nwarps; ++w)
			c += hall[(w * nb + b) * 32 + lane];
#pragma unroll
		for (int k = 16; k > 0; k >>= 1)
			c += __shfl_xor_sync(0xffffffffu, c, k);
		if (lane == 0 && c)
			atomicAdd(counts + b, c);
	}
}

__global__ void __launch_bounds__(MG_HIST_THREADS) k_dense_stream_hist(const RowSrc rs, const DenseChrom *table, int64_t first, int64_t count,
																		const HistSpec hs, unsigned long long *counts)
{
	extern __shared__ __align__(16) uint32_t hsmem[];
	mg_hist1d_private(hs, hsmem, counts, [&](auto &&emit) { mg_stream_rows(rs, table, first, count, emit); });
}

__global__ void __launch_bounds__(MG_HIST_THREADS) k_dense_hist_private(const DenseQuery q, int func, const HistSpec hs, unsigned long long *counts)
{
	extern __shared__ __align__(16) uint32_t hsmem[];
	mg_hist1d_private(hs, hsmem, counts, [&](auto &&emit) {
		const int64_t stride = (int64_t)gridDim.x * blockDim.x;
		for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < q.count; i += stride)
			emit(mg_dense_row_value(q, q.first + i, func));
	});
}

__global__ void __launch_bounds__(MG_HIST_THREADS) k_values_hist_private(const double *__restrict__ vals, int64_t n, const HistSpec hs,
																		  unsigned long long *counts)
{
	extern __shared__ __align__(16) uint32_t hsmem[];
	mg_hist1d_private(hs, hsmem, counts, [&](auto &&emit) {
		const int64_t stride = (int64_t)gridDim.x * blockDim.x;
		for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
			emit(__ldcs(vals + i));
	});
}

// Float inputs (dense-track values are float32 widened to double): val2bin is monotone in the value, so the bin of a
// float is the number of float thresholds below it. thr[j], j = 0..nb, = the largest float whose bin is < j (bin -1 =
// below the first break, nb = above the last); computed on the host by bisection with the double formula, hence exact.
__device__ __forceinline__ int mg_float_bin(const float *__restrict__ thr, int nb, float v)
{
	if (isnan(v))
		return -1;
	int lo = 0, hi = nb + 1; // count of thresholds < v lies in [lo, hi]
	while (lo < hi) {
		const int mid = (lo + hi) >> 1;
		if (thr[mid] < v)
			lo = mid + 1;
		else
			hi = mid;
	}
	return lo == nb + 1 ? -1 : lo - 1; // lo thresholds are below v: bin = lo - 1 (-1: too low); all nb+1 below: too high
}

// Thresholds travel as a kernel parameter (no upload, no synchronisation). With equal-width breaks the bin is guessed with
// one float multiply and checked against the exact thresholds on both sides (two shared-memory loads); a wrong guess walks.
struct FloatThr {
	int nb;
	int uniform;
	float b0, inv_binsize;
	float near; // a position whose fraction keeps this far from 0 and 1 has the count floor(position): see k_segs_hist (>= 1: never)
	float thr[MG_HIST_PRIVATE_MAX + 1];
};

// shared-memory accesses by 32-bit shared address: keeps the base in a register (the compiler otherwise re-derives the
// shared window base for every access of an extern __shared__ array reached through a pointer)
__device__ __forceinline__ float mg_lds_f(uint32_t a)
{
	float v;
	asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
	return v;
}
template <typename C>
__device__ __forceinline__ void mg_smem_inc(uint32_t a);
template <>
__device__ __forceinline__ void mg_smem_inc<uint32_t>(uint32_t a)
{
	uint32_t c;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(c) : "r"(a) : "memory");
	asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(c + 1u) : "memory");
}
template <>
__device__ __forceinline__ void mg_smem_inc<uint16_t>(uint32_t a)
{
	uint16_t c;
	asm volatile("ld.shared.u16 %0, [%1];" : "=h"(c) : "r"(a) : "memory");
	asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((uint16_t)(c + 1)) : "memory");
}

// C = counter type of the private columns: uint16_t halves the shared memory (twice the resident warps) and is used whenever
// no thread can see more than 65535 values (the host checks tiles per block x 16).
// PAIR (with C = uint16_t): the 16-bit counters of warps w and w + nwarps / 2 share 32-bit words (low / high half), and a count is
// ONE shared-memory reduction `red.shared.add.u32` of 1 or 65536 -- no load whose latency the thread has to wait out before
// its store, half the shared-memory instructions -- instead of a 16-bit load / add / store; same bytes of shared memory.
template <typename C, bool PAIR>
__global__ void __launch_bounds__(MG_HIST_THREADS) k_segs_hist(const SegTable st, const FloatThr ft, unsigned long long *counts)
{
	extern __shared__ __align__(16) uint32_t hsmem[];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
	const int nb = ft.nb;
	// P[0] = -inf, P[j] = thr[j-1] (j = 1..nb+1), P[nb+2] = +inf: c = #thresholds below v is right iff P[c] < v <= P[c+1]
	float *P = reinterpret_cast<float *>(hsmem);
	const int thr_words = (nb + 3 + 3) & ~3;
	C *hall = reinterpret_cast<C *>(hsmem + thr_words), *h = hall + warp * nb * 32;
	uint32_t *hw = hsmem + thr_words; // PAIR: [nwarps / 2][nb][32] words
	const int npairs = nwarps >> 1, pair = warp % npairs, half = warp / npairs;
	for (int t = threadIdx.x; t <= nb + 2; t += blockDim.x)
		P[t] = t == 0 ? __int_as_float(0xff800000) : (t == nb + 2 ? __int_as_float(0x7f800000) : ft.thr[t - 1]);
	if (PAIR) {
		for (int t = threadIdx.x; t < npairs * nb * 32; t += blockDim.x)
			hw[t] = 0;
	} else {
		for (int t = lane; t < nb * 32; t += 32)
			h[t] = 0;
	}
	__syncthreads();
	constexpr uint32_t COL = PAIR ? 128u : 32u * sizeof(C); // bytes per bin column
	const uint32_t pbase = (uint32_t)__cvta_generic_to_shared(P);
	// column of bin c - 1 at hbase + COL c
	const uint32_t hbase = PAIR ? (uint32_t)__cvta_generic_to_shared(hw + pair * nb * 32) + 4u * lane - COL
								: (uint32_t)__cvta_generic_to_shared(h) + (uint32_t)sizeof(C) * lane - COL;
	const uint32_t one = 1u << (16 * half);
	auto count_bin = [&](uint32_t addr) {
		if (PAIR)
			asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(one) : "memory");
		else
			mg_smem_inc<C>(addr);
	};
	if (ft.uniform) {
		const float b0 = ft.b0, inv = ft.inv_binsize, top = (float)(nb + 1);
		// t = the value's position in units of bins (+ 1). The host bounds its distance from the true position among these breaks (ft.near
		// is four times that bound): a t whose fraction keeps that far from both ends has the count floor(t) without a look at the
		// thresholds; the few values near a break -- and everything negative, huge or NaN -- take the exact path
		const float NEAR = ft.near;
		mg_stream_segtable(st, [&](float v) {
			const float t = (v - b0) * inv + 1.f;
			int c = __float2int_rz(t);
			const float fr = t - (float)c;
			if (!(fr >= NEAR && fr <= 1.f - NEAR)) {
				if (v != v)
					return;
				c = (int)fminf(fmaxf(t, 0.f), top);
				const float t0 = mg_lds_f(pbase + 4u * c), t1 = mg_lds_f(pbase + 4u * c + 4u);
				if (!(t0 < v) || t1 < v) { // the float guess missed: walk to the right count (sentinels bound both walks)
					while (mg_lds_f(pbase + 4u * c + 4u) < v)
						++c;
					while (!(mg_lds_f(pbase + 4u * c) < v))
						--c;
				}
			}
			if ((unsigned)(c - 1) < (unsigned)nb)
				count_bin(hbase + COL * c);
		});
	} else {
		mg_stream_segtable(st, [&](float v) {
			const int b = mg_float_bin(P + 1, nb, v);
			if (b >= 0)
				count_bin(hbase + COL * (b + 1));
		});
	}
	__syncthreads();
	for (int b = warp; b < nb; b += nwarps) {
		unsigned long long c = 0;
		if (PAIR) {
			for (int p = 0; p < npairs; ++p) {
				const uint32_t w = hw[(p * nb + b) * 32 + lane];
				c += (w & 0xffffu) + (w >> 16);
			}
		} else {
			for (int w = 0; w < nwarps; ++w)
				c += hall[(w * nb + b) * 32 + lane];
		}
#pragma unroll
		for (int k = 16; k > 0; k >>= 1)
			c += __shfl_xor_sync(0xffffffffu, c, k);
		if (lane == 0 && c)
			atomicAdd(counts + b, c);
	}
}

// general N-D histogram over device columns: global atomics (one per surviving row)
struct HistCols {
	const double *col[MG_HIST_MAX_DIMS];
};

__global__ void __launch_bounds__(MG_HIST_THREADS) k_hist_nd(const HistCols cols, int64_t n, const HistSpec hs, unsigned long long *counts)
{
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		long long idx = 0;
		bool ok = true;
		for (int d = 0; d < hs.ndim && ok; ++d) {
			const double v = __ldcs(cols.col[d] + i);
			const int b = mg_val2bin(hs.breaks + hs.boff[d], hs.nbreaks[d], hs.binsize[d], hs.include_lowest, v);
			if (b < 0)
				ok = false;
			else
				idx += b * hs.mult[d];
		}
		if (ok)
			atomicAdd(counts + idx, 1ull);
	}
}

// general 1-D histogram of a dense function that is too wide for private columns
__global__ void __launch_bounds__(MG_HIST_THREADS) k_dense_hist_global(const DenseQuery q, int func, const HistSpec hs, unsigned long long *counts)
{
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < q.count; i += stride) {
		const int b = mg_val2bin(hs.breaks, hs.nbreaks[0], hs.binsize[0], hs.include_lowest, mg_dense_row_value(q, q.first + i, func));
		if (b >= 0)
			atomicAdd(counts + b, 1ull);
	}
}

// ---- gquantiles: exact order statistics of a whole value column ------------------------------------------------------------
// Three counting passes over the column (11 + 11 + 10 bits of the order-preserving float key) find every requested rank at
// once: pass 0 histograms the top digit of all keys; passes 1 and 2 only count keys whose prefix is one a target rank fell
// into (a 2048-entry table in shared memory maps the top digit to its prefix slots). No sort, no temporary copy of the data.
#define MG_QS_MAXSLOTS 128
struct QsPass {
	int pass;   // 0, 1, 2
	int nslots; // distinct prefixes still followed (passes 1, 2)
	uint32_t prefix[MG_QS_MAXSLOTS]; // pass 1: top 11 bits; pass 2: top 22 bits; ascending
};

__device__ __forceinline__ uint32_t mg_qs_key(double v)
{
	const float f = (float)v; // the reference narrows every row value to float first (src/GenomeTrackQuantiles.cpp:168)
	if (f != f)
		return 0xffffffffu;
	const uint32_t b = __float_as_uint(f);
	return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}

__global__ void __launch_bounds__(256) k_qs_count(const double *__restrict__ vals, int64_t n, const QsPass p, unsigned long long *hist,
												   unsigned long long *nan_count)
{
	__shared__ uint32_t sh[2048];     // pass 0: the histogram; passes 1-2: first slot of every top digit (or 0xffffffff)
	for (int t = threadIdx.x; t < 2048; t += blockDim.x)
		sh[t] = p.pass == 0 ? 0u : 0xffffffffu;
	__syncthreads();
	if (p.pass > 0) {
		for (int t = threadIdx.x; t < p.nslots; t += blockDim.x) {
			const uint32_t top = p.pass == 1 ? p.prefix[t] : p.prefix[t] >> 11;
			atomicMin(&sh[top], (uint32_t)t); // prefixes ascend: the slots of one top digit are contiguous
		}
		__syncthreads();
	}
	unsigned long long nans = 0;
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		const uint32_t key = mg_qs_key(__ldcs(vals + i));
		if (key == 0xffffffffu) {
			++nans;
			continue;
		}
		if (p.pass == 0) {
			atomicAdd(&sh[key >> 21], 1u);
			continue;
		}
		uint32_t slot = sh[key >> 21];
		if (slot == 0xffffffffu)
			continue;
		if (p.pass == 1) {
			atomicAdd(hist + (size_t)slot * 2048 + ((key >> 10) & 2047u), 1ull);
		} else {
			const uint32_t pre = key >> 10;
			for (; slot < (uint32_t)p.nslots && (p.prefix[slot] >> 11) == (pre >> 11); ++slot)
				if (p.prefix[slot] == pre) {
					atomicAdd(hist + (size_t)slot * 1024 + (key & 1023u), 1ull);
					break;
				}
		}
	}
	if (p.pass == 0) {
		__syncthreads();
		for (int t = threadIdx.x; t < 2048; t += blockDim.x)
			if (sh[t])
				atomicAdd(hist + t, (unsigned long long)sh[t]);
		if (nans)
			atomicAdd(nan_count, nans);
	}
}

// ---- gintervals.quantiles: the quantiles of many contiguous segments of a value column in one launch ------------------------
// The reference keeps one StreamPercentiler<double> and resets it at every scope interval (src/GenomeTrackQuantiles.cpp:726-741);
// a segment is the run of rows of one scope interval. A segment's keys (mg_qs_key: float-narrowed, NaN = 0xffffffff sorts last,
// like the padding) are sorted in shared memory -- by one warp (WARP, up to MG_QSEG_WARP_CAP rows) or one thread block (up to
// MG_QSEG_BLOCK_CAP rows) -- and each quantile is interpolated in double as src/StreamPercentiler.h:210-216 does.
#define MG_QSEG_WARP_CAP 256
#define MG_QSEG_BLOCK_CAP 16384
#define MG_QSEG_BLOCK_THREADS 512
struct QsegArgs {
	const double *vals;
	const int64_t *seg_first; // nseg + 1 ascending row offsets
	const int32_t *list;      // the segments of this launch
	int64_t nlist, nseg;
	int nq;
	const double *pct; // nq percentiles in [0, 1]
	double *out;       // out[k * nseg + s]
	long long *nkept;  // non-NaN values of segment s
};

template <bool WARP> __global__ void __launch_bounds__(WARP ? 256 : MG_QSEG_BLOCK_THREADS) k_qseg(const QsegArgs a)
{
	extern __shared__ uint32_t qseg_sh[];
	const int T = WARP ? 32 : (int)blockDim.x;
	const int tid = WARP ? (int)(threadIdx.x & 31) : (int)threadIdx.x;
	uint32_t *s = WARP ? qseg_sh + (threadIdx.x >> 5) * MG_QSEG_WARP_CAP : qseg_sh;
	const int64_t worker = WARP ? (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5) : (int64_t)blockIdx.x;
	const int64_t nworkers = WARP ? (int64_t)gridDim.x * (blockDim.x >> 5) : (int64_t)gridDim.x;
	for (int64_t e = worker; e < a.nlist; e += nworkers) {
		const int64_t seg = a.list[e];
		const int64_t first = a.seg_first[seg];
		const int c = (int)(a.seg_first[seg + 1] - first);
		int P = 32;
		while (P < c)
			P <<= 1;
		for (int t = tid; t < P; t += T)
			s[t] = t < c ? mg_qs_key(__ldcs(a.vals + first + t)) : 0xffffffffu;
		if (WARP) __syncwarp(); else __syncthreads();
		for (int k = 2; k <= P; k <<= 1) {
			for (int j = k >> 1; j > 0; j >>= 1) {
				for (int t = tid; t < (P >> 1); t += T) {
					const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
					const int ixj = i | j;
					const uint32_t x = s[i], y = s[ixj];
					if ((x > y) == ((i & k) == 0)) {
						s[i] = y;
						s[ixj] = x;
					}
				}
				if (WARP) __syncwarp(); else __syncthreads();
			}
		}
		// N = number of keys below the NaN key (every non-NaN float's key is <= 0xff800000)
		int lo = 0, hi = c;
		while (lo < hi) {
			const int mid = (lo + hi) >> 1;
			if (s[mid] == 0xffffffffu)
				hi = mid;
			else
				lo = mid + 1;
		}
		const int N = lo;
		if (tid == 0)
			a.nkept[seg] = N;
		for (int k = tid; k < a.nq; k += T) {
			double r = mg_nan();
			if (N > 0) {
				const double idx = __dmul_rn((double)(N - 1), a.pct[k]);
				const double f1 = floor(idx);
				const int i1 = (int)f1, i2 = (int)ceil(idx);
				const double w = __dsub_rn(idx, f1);
				r = __dadd_rn(__dmul_rn((double)mg_fkey_inv(s[i1]), __dsub_rn(1.0, w)), __dmul_rn((double)mg_fkey_inv(s[i2]), w));
			}
			a.out[(int64_t)k * a.nseg + seg] = r;
		}
		if (WARP) __syncwarp(); else __syncthreads();
	}
}

// Segments too long for a thread block's shared memory: one thread block per segment selects the target ranks by MSD radix
// select over the segment in global memory -- four passes of 8 bits; a pass counts, per distinct key prefix a target rank fell
// into (a sorted list in shared memory, found by bisection), the next digit of the keys under it (mg_warp_count).
// Count `what` (0xffffffff = nothing) into hist, one shared-memory atomic per distinct value for the first two values a warp
// holds (the high digits of a track's keys are nearly constant across a warp), one per lane for the rest.
__device__ __forceinline__ void mg_warp_count(uint32_t *hist, uint32_t what, int lane)
{
	unsigned remaining = __ballot_sync(0xffffffffu, what != 0xffffffffu);
#pragma unroll
	for (int r = 0; r < 2; ++r) {
		if (!remaining)
			return;
		const int leader = __ffs(remaining) - 1;
		const uint32_t lv = __shfl_sync(0xffffffffu, what, leader);
		const unsigned m = __ballot_sync(0xffffffffu, what == lv) & remaining;
		if (lane == leader)
			atomicAdd(&hist[lv], (uint32_t)__popc(m));
		remaining &= ~m;
	}
	if ((remaining >> lane) & 1u)
		atomicAdd(&hist[what], 1u);
}

#define MG_QSEL_THREADS 1024
#define MG_QSEL_UNROLL 8 // keys a thread has in flight: one thread block streams its segment alone
#define MG_QSEL_MAXT (MG_QS_MAXSLOTS) // 2 x 64 quantiles
__global__ void __launch_bounds__(MG_QSEL_THREADS) k_qseg_select(const QsegArgs a)
{
	extern __shared__ uint32_t qsel_hist[]; // [slot][256]
	__shared__ uint32_t s_slot[MG_QSEL_MAXT];          // distinct prefixes of this pass, ascending
	__shared__ uint32_t s_prefix[MG_QSEL_MAXT];        // per target: the digits of its key found so far
	__shared__ unsigned long long s_rank[MG_QSEL_MAXT]; // per target: its rank among the keys under its prefix
	__shared__ unsigned long long s_target[MG_QSEL_MAXT];
	__shared__ int s_T, s_S;
	__shared__ unsigned long long s_nan;
	const int tid = threadIdx.x, lane = tid & 31;
	for (int64_t e = blockIdx.x; e < a.nlist; e += gridDim.x) {
		const int64_t seg = a.list[e];
		const int64_t first = a.seg_first[seg];
		const int64_t c = a.seg_first[seg + 1] - first, c32 = (c + 31) & ~(int64_t)31;
		const double *v = a.vals + first;
		// pass 0: top digit of every key, NaN rows
		for (int t = tid; t < 256; t += MG_QSEL_THREADS)
			qsel_hist[t] = 0;
		if (tid == 0)
			s_nan = 0;
		__syncthreads();
		unsigned long long nans = 0;
		for (int64_t base = 0; base < c32; base += MG_QSEL_UNROLL * MG_QSEL_THREADS) {
			uint32_t key[MG_QSEL_UNROLL];
			double raw[MG_QSEL_UNROLL]; // all loads first (clamped, unconditional): a thread block streams its segment alone
#pragma unroll
			for (int j = 0; j < MG_QSEL_UNROLL; ++j)
				raw[j] = __ldcs(v + min(base + j * MG_QSEL_THREADS + tid, c - 1));
#pragma unroll
			for (int j = 0; j < MG_QSEL_UNROLL; ++j)
				key[j] = base + j * MG_QSEL_THREADS + tid < c ? mg_qs_key(raw[j]) : 0xfffffffeu; // beyond the segment: neither a key nor a NaN row
#pragma unroll
			for (int j = 0; j < MG_QSEL_UNROLL; ++j) {
				nans += key[j] == 0xffffffffu;
				const uint32_t what = key[j] >= 0xfffffffeu ? 0xffffffffu : key[j] >> 24;
				mg_warp_count(qsel_hist, what, lane);
			}
		}
		nans = __reduce_add_sync(0xffffffffu, (unsigned)nans);
		if (lane == 0 && nans)
			atomicAdd(&s_nan, nans);
		__syncthreads();
		const long long N = c - (long long)s_nan;
		if (tid == 0) {
			a.nkept[seg] = N;
			int T = 0;
			if (N > 0) {
				for (int k = 0; k < a.nq; ++k) { // floor / ceil of (N - 1) q, sorted and unique (insertion: T <= 128)
					const double idx = __dmul_rn((double)(N - 1), a.pct[k]);
					const unsigned long long two[2] = {(unsigned long long)floor(idx), (unsigned long long)ceil(idx)};
					for (int h = 0; h < 2; ++h) {
						int j = 0;
						while (j < T && s_target[j] < two[h])
							++j;
						if (j < T && s_target[j] == two[h])
							continue;
						for (int m = T; m > j; --m)
							s_target[m] = s_target[m - 1];
						s_target[j] = two[h];
						++T;
					}
				}
			}
			s_T = T;
		}
		__syncthreads();
		const int T = s_T;
		if (T == 0) {
			for (int k = tid; k < a.nq; k += MG_QSEL_THREADS)
				a.out[(int64_t)k * a.nseg + seg] = mg_nan();
			__syncthreads();
			continue;
		}
		if (tid < T) { // digit 0 of every target
			unsigned long long r = s_target[tid], acc = 0;
			int d = 0;
			for (; d < 255; ++d) {
				if (acc + qsel_hist[d] > r)
					break;
				acc += qsel_hist[d];
			}
			s_rank[tid] = r - acc;
			s_prefix[tid] = (uint32_t)d;
		}
		__syncthreads();
		for (int pass = 1; pass < 4; ++pass) {
			const int shift = 32 - 8 * pass; // the key's bits above `shift` are the prefix, the 8 below the digit counted now
			if (tid == 0) {
				int S = 0;
				for (int t = 0; t < T; ++t)
					if (!S || s_slot[S - 1] != s_prefix[t]) // targets ascend, so do their prefixes
						s_slot[S++] = s_prefix[t];
				s_S = S;
			}
			__syncthreads();
			const int S = s_S;
			for (int t = tid; t < S * 256; t += MG_QSEL_THREADS)
				qsel_hist[t] = 0;
			__syncthreads();
			for (int64_t base = 0; base < c32; base += MG_QSEL_UNROLL * MG_QSEL_THREADS) {
				uint32_t keys[MG_QSEL_UNROLL];
				double raw[MG_QSEL_UNROLL];
#pragma unroll
				for (int j = 0; j < MG_QSEL_UNROLL; ++j)
					raw[j] = __ldcs(v + min(base + j * MG_QSEL_THREADS + tid, c - 1));
#pragma unroll
				for (int j = 0; j < MG_QSEL_UNROLL; ++j)
					keys[j] = base + j * MG_QSEL_THREADS + tid < c ? mg_qs_key(raw[j]) : 0xffffffffu;
#pragma unroll
				for (int j = 0; j < MG_QSEL_UNROLL; ++j) {
					uint32_t what = 0xffffffffu;
					{
						const uint32_t key = keys[j];
						if (key != 0xffffffffu) {
							const uint32_t pre = key >> shift;
							int lo = 0, hi = S;
							while (lo < hi) {
								const int mid = (lo + hi) >> 1;
								if (s_slot[mid] < pre)
									lo = mid + 1;
								else
									hi = mid;
							}
							if (lo < S && s_slot[lo] == pre)
								what = ((uint32_t)lo << 8) | ((key >> (shift - 8)) & 255u);
						}
					}
					mg_warp_count(qsel_hist, what, lane);
				}
			}
			__syncthreads();
			if (tid < T) {
				int j = 0;
				while (s_slot[j] != s_prefix[tid])
					++j;
				const uint32_t *h = qsel_hist + j * 256;
				unsigned long long r = s_rank[tid], acc = 0;
				int d = 0;
				for (; d < 255; ++d) {
					if (acc + h[d] > r)
						break;
					acc += h[d];
				}
				s_rank[tid] = r - acc;
				s_prefix[tid] = (s_prefix[tid] << 8) | (uint32_t)d;
			}
			__syncthreads();
		}
		for (int k = tid; k < a.nq; k += MG_QSEL_THREADS) {
			const double idx = __dmul_rn((double)(N - 1), a.pct[k]);
			const double f1 = floor(idx);
			const unsigned long long i1 = (unsigned long long)f1, i2 = (unsigned long long)ceil(idx);
			int p1 = 0, p2 = 0;
			while (s_target[p1] != i1)
				++p1;
			while (s_target[p2] != i2)
				++p2;
			const double w = __dsub_rn(idx, f1);
			a.out[(int64_t)k * a.nseg + seg] =
				__dadd_rn(__dmul_rn((double)mg_fkey_inv(s_prefix[p1]), __dsub_rn(1.0, w)), __dmul_rn((double)mg_fkey_inv(s_prefix[p2]), w));
		}
		__syncthreads();
	}
}

// Segments beyond MG_QLONG_MIN rows: the same radix select with the counting spread over the device -- k_qlong_count, one thread
// block per chunk of MG_QLONG_CHUNK rows (shared-memory histograms flushed into the segment's global ones), and k_qlong_update,
// one thread block per segment, which descends the target ranks by one digit and prepares the next pass; eight launches, no
// host round trip in between.
#define MG_QLONG_MIN ((int64_t)1 << 18)
#define MG_QLONG_CHUNK ((int64_t)1 << 16)
struct QlongState {
	unsigned long long nan;
	long long N;
	int T, S;
	unsigned long long target[MG_QSEL_MAXT], rank[MG_QSEL_MAXT];
	uint32_t prefix[MG_QSEL_MAXT], slot[MG_QSEL_MAXT];
};
struct QlongArgs {
	QsegArgs a;                // list = the long segments
	const int32_t *chunk_seg;  // per chunk: index into a.list
	const int64_t *chunk_off;  // per chunk: first row, relative to its segment
	QlongState *state;         // per long segment
	uint32_t *hist;            // per long segment: maxS x 256 counters, zero before every counting pass
	int maxS;
};

__global__ void __launch_bounds__(MG_QSEL_THREADS) k_qlong_count(const QlongArgs q, int pass)
{
	extern __shared__ uint32_t qsel_hist[];
	__shared__ uint32_t s_slot[MG_QSEL_MAXT];
	__shared__ unsigned long long s_nan;
	const int tid = threadIdx.x, lane = tid & 31;
	const int e = q.chunk_seg[blockIdx.x];
	const int64_t seg = q.a.list[e];
	const int64_t first = q.a.seg_first[seg] + q.chunk_off[blockIdx.x];
	const int64_t c = min(q.a.seg_first[seg + 1] - first, MG_QLONG_CHUNK);
	const double *v = q.a.vals + first;
	const QlongState *stt = q.state + e;
	const int S = pass == 0 ? 1 : stt->S;
	if (pass > 0 && stt->T == 0)
		return;
	for (int t = tid; t < S * 256; t += MG_QSEL_THREADS)
		qsel_hist[t] = 0;
	if (tid < S && pass > 0)
		s_slot[tid] = stt->slot[tid];
	if (tid == 0)
		s_nan = 0;
	__syncthreads();
	const int shift = 32 - 8 * pass;
	unsigned nans = 0;
	for (int64_t base = 0; base < c; base += MG_QSEL_UNROLL * MG_QSEL_THREADS) {
		double raw[MG_QSEL_UNROLL];
#pragma unroll
		for (int j = 0; j < MG_QSEL_UNROLL; ++j)
			raw[j] = __ldcs(v + min(base + j * MG_QSEL_THREADS + tid, c - 1));
#pragma unroll
		for (int j = 0; j < MG_QSEL_UNROLL; ++j) {
			const bool in = base + j * MG_QSEL_THREADS + tid < c;
			const uint32_t key = in ? mg_qs_key(raw[j]) : 0xffffffffu;
			uint32_t what = 0xffffffffu;
			if (pass == 0) {
				nans += in && key == 0xffffffffu;
				if (key != 0xffffffffu)
					what = key >> 24;
			} else if (key != 0xffffffffu) {
				const uint32_t pre = key >> shift;
				int lo = 0, hi = S;
				while (lo < hi) {
					const int mid = (lo + hi) >> 1;
					if (s_slot[mid] < pre)
						lo = mid + 1;
					else
						hi = mid;
				}
				if (lo < S && s_slot[lo] == pre)
					what = ((uint32_t)lo << 8) | ((key >> (shift - 8)) & 255u);
			}
			mg_warp_count(qsel_hist, what, lane);
		}
	}
	if (pass == 0) {
		nans = __reduce_add_sync(0xffffffffu, nans);
		if (lane == 0 && nans)
			atomicAdd(&s_nan, (unsigned long long)nans);
	}
	__syncthreads();
	uint32_t *g = q.hist + (size_t)e * q.maxS * 256;
	for (int t = tid; t < S * 256; t += MG_QSEL_THREADS)
		if (qsel_hist[t])
			atomicAdd(g + t, qsel_hist[t]);
	if (pass == 0 && tid == 0 && s_nan)
		atomicAdd(&q.state[e].nan, s_nan);
}

__global__ void __launch_bounds__(256) k_qlong_update(const QlongArgs q, int pass)
{
	const int e = blockIdx.x, tid = threadIdx.x;
	const int64_t seg = q.a.list[e];
	QlongState *stt = q.state + e;
	uint32_t *g = q.hist + (size_t)e * q.maxS * 256;
	if (pass == 0) {
		if (tid == 0) {
			const long long N = (q.a.seg_first[seg + 1] - q.a.seg_first[seg]) - (long long)stt->nan;
			stt->N = N;
			q.a.nkept[seg] = N;
			int T = 0;
			if (N > 0)
				for (int k = 0; k < q.a.nq; ++k) {
					const double idx = __dmul_rn((double)(N - 1), q.a.pct[k]);
					const unsigned long long two[2] = {(unsigned long long)floor(idx), (unsigned long long)ceil(idx)};
					for (int h = 0; h < 2; ++h) {
						int j = 0;
						while (j < T && stt->target[j] < two[h])
							++j;
						if (j < T && stt->target[j] == two[h])
							continue;
						for (int m = T; m > j; --m)
							stt->target[m] = stt->target[m - 1];
						stt->target[j] = two[h];
						++T;
					}
				}
			stt->T = T;
		}
		__syncthreads();
	}
	const int T = stt->T;
	if (T == 0) {
		if (pass == 3)
			for (int k = tid; k < q.a.nq; k += blockDim.x)
				q.a.out[(int64_t)k * q.a.nseg + seg] = mg_nan();
		return;
	}
	const int S_old = pass == 0 ? 1 : stt->S;
	if (tid < T) {
		int j = 0;
		if (pass > 0)
			while (stt->slot[j] != stt->prefix[tid])
				++j;
		const uint32_t *h = g + j * 256;
		unsigned long long r = pass == 0 ? stt->target[tid] : stt->rank[tid], acc = 0;
		int d = 0;
		for (; d < 255; ++d) {
			if (acc + h[d] > r)
				break;
			acc += h[d];
		}
		stt->rank[tid] = r - acc;
		stt->prefix[tid] = pass == 0 ? (uint32_t)d : (stt->prefix[tid] << 8) | (uint32_t)d;
	}
	__syncthreads();
	if (pass < 3) {
		for (int t = tid; t < S_old * 256; t += blockDim.x) // the counters of the pass just read
			g[t] = 0;
		if (tid == 0) {
			int S = 0;
			for (int t = 0; t < T; ++t)
				if (!S || stt->slot[S - 1] != stt->prefix[t])
					stt->slot[S++] = stt->prefix[t];
			stt->S = S;
		}
		return;
	}
	const long long N = stt->N;
	for (int k = tid; k < q.a.nq; k += blockDim.x) {
		const double idx = __dmul_rn((double)(N - 1), q.a.pct[k]);
		const double f1 = floor(idx);
		const unsigned long long i1 = (unsigned long long)f1, i2 = (unsigned long long)ceil(idx);
		int p1 = 0, p2 = 0;
		while (stt->target[p1] != i1)
			++p1;
		while (stt->target[p2] != i2)
			++p2;
		const double w = __dsub_rn(idx, f1);
		q.a.out[(int64_t)k * q.a.nseg + seg] =
			__dadd_rn(__dmul_rn((double)mg_fkey_inv(stt->prefix[p1]), __dsub_rn(1.0, w)), __dmul_rn((double)mg_fkey_inv(stt->prefix[p2]), w));
	}
}

// ---- gintervals.summary: one IntervalSummary per scope interval (src/GenomeTrackSummary.cpp:248-431) ---------------------------
// A warp per segment (the run of rows of one scope interval); the row value is narrowed to float first, as `float val =
// scanner.last_real(0)` does. Segments beyond `big` rows only get their record initialised: the host sends those through
// mg_summary's whole-device kernels.
__global__ void __launch_bounds__(256) k_summary_segs(const double *__restrict__ vals, const int64_t *__restrict__ seg_first, int64_t nseg,
													   int64_t big, double *partials)
{
	const int lane = threadIdx.x & 31;
	const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
	for (int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); s < nseg; s += nwarps) {
		const int64_t first = seg_first[s], c = seg_first[s + 1] - first;
		Moments m;
		mom_init(m);
		if (c <= big)
			for (int64_t i = lane; i < c; i += 32)
				mom_add(m, (double)(float)__ldcs(vals + first + i));
#pragma unroll
		for (int k = 16; k > 0; k >>= 1) {
			Moments o = mom_shfl_xor(m, k);
			mom_merge(m, o);
		}
		if (lane == 0) {
			double *p = partials + 6 * s;
			p[0] = m.n, p[1] = m.nn, p[2] = m.sum, p[3] = m.mn, p[4] = m.mx, p[5] = m.ssq;
		}
	}
}

// ---- gbins.summary / gbins.quantiles: rows grouped by the N-D bin of their bin columns -----------------------------------------
// (src/GenomeTrackSummary.cpp:433-506, src/GenomeTrackQuantiles.cpp:1199-1300; bin = BinsManager::vals2idx, src/BinsManager.h:53-72)
__device__ __forceinline__ long long mg_vals2idx(const HistCols &cols, const HistSpec &hs, int64_t i)
{
	long long idx = 0;
	for (int d = 0; d < hs.ndim; ++d) {
		const int b = mg_val2bin(hs.breaks + hs.boff[d], hs.nbreaks[d], hs.binsize[d], hs.include_lowest, __ldcs(cols.col[d] + i));
		if (b < 0)
			return -1;
		idx += b * hs.mult[d];
	}
	return idx;
}

// order-preserving 64-bit key of a double (no NaN reaches it): min / max become integer atomics
__device__ __forceinline__ unsigned long long mg_dkey(double v)
{
	const unsigned long long b = (unsigned long long)__double_as_longlong(v);
	return b ^ ((b >> 63) ? ~0ull : 0x8000000000000000ull);
}
__device__ __forceinline__ double mg_dkey_inv(unsigned long long k)
{
	const unsigned long long b = k ^ ((k >> 63) ? 0x8000000000000000ull : ~0ull);
	return __longlong_as_double((long long)b);
}

// accumulator record of one bin: {rows, non-NaN rows, sum, min key, max key, sum of squares}, 6 x 8 bytes
#define MG_BINS_SMEM_MAX 1024 // bins whose records fit a thread block's 48 KB of shared memory
__device__ __forceinline__ void mg_bins_rec_init(unsigned long long *r)
{
	r[0] = 0, r[1] = 0, r[2] = 0 /* 0.0 */, r[3] = ~0ull, r[4] = 0, r[5] = 0;
}

__global__ void __launch_bounds__(256) k_bins_init(unsigned long long *acc, long long total)
{
	for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < total; b += (long long)gridDim.x * blockDim.x)
		mg_bins_rec_init(acc + 6 * b);
}

// SMEM: `rep` (a power of two) copies of every bin's record in shared memory, a lane uses copy lane % rep -- the rows of a warp
// crowd into few bins, and a shared-memory double add is a compare-and-swap loop that same-address neighbours make retry.
template <bool SMEM> __global__ void __launch_bounds__(256) k_bins_summary(const double *__restrict__ vals, const HistCols cols, int64_t n,
																			 const HistSpec hs, int rep, unsigned long long *acc)
{
	extern __shared__ unsigned long long bins_sh[];
	unsigned long long *a = SMEM ? bins_sh : acc;
	const int nrec = SMEM ? (int)hs.total * rep : 0;
	const int mine = SMEM ? (int)(threadIdx.x & (rep - 1)) : 0;
	if (SMEM) {
		for (int b = threadIdx.x; b < nrec; b += blockDim.x)
			mg_bins_rec_init(a + 6 * b);
		__syncthreads();
	}
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		const long long idx = mg_vals2idx(cols, hs, i);
		if (idx < 0)
			continue;
		unsigned long long *r = a + 6 * (SMEM ? idx * rep + mine : idx);
		const double v = (double)(float)__ldcs(vals + i);
		atomicAdd(r, 1ull);
		if (!isnan(v)) {
			atomicAdd(r + 1, 1ull);
			atomicAdd((double *)(r + 2), v);
			atomicMin(r + 3, mg_dkey(v));
			atomicMax(r + 4, mg_dkey(v));
			atomicAdd((double *)(r + 5), __dmul_rn(v, v));
		}
	}
	if (SMEM) {
		__syncthreads();
		for (int b = threadIdx.x; b < nrec; b += blockDim.x) {
			const unsigned long long *r = a + 6 * b;
			unsigned long long *g = acc + 6 * (long long)(b / rep);
			if (!r[0])
				continue;
			atomicAdd(g, r[0]);
			if (r[1]) {
				atomicAdd(g + 1, r[1]);
				atomicAdd((double *)(g + 2), __longlong_as_double((long long)r[2]));
				atomicMin(g + 3, r[3]);
				atomicMax(g + 4, r[4]);
				atomicAdd((double *)(g + 5), __longlong_as_double((long long)r[5]));
			}
		}
	}
}

// accumulator records -> IntervalSummary's six doubles {num_bins, num_non_nan, sum, min, max, sum_sq}
__global__ void __launch_bounds__(256) k_bins_finalize(const unsigned long long *acc, long long total, double *partials)
{
	for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < total; b += (long long)gridDim.x * blockDim.x) {
		const unsigned long long *r = acc + 6 * b;
		double *p = partials + 6 * b;
		p[0] = (double)r[0];
		p[1] = (double)r[1];
		p[2] = __longlong_as_double((long long)r[2]);
		p[3] = r[1] ? mg_dkey_inv(r[3]) : 1.7976931348623157e308;
		p[4] = r[1] ? mg_dkey_inv(r[4]) : -1.7976931348623157e308;
		p[5] = __longlong_as_double((long long)r[5]);
	}
}

// gbins.quantiles, step 1: the bin of every row whose (float-narrowed) value is not NaN, -1 otherwise, and the bins' populations
// (SMEM: counted in a shared-memory histogram per thread block, flushed once; otherwise global atomics -- many bins, little contention).
#define MG_BINS_SCATTER_SMEM_MAX 2048
#define MG_BINS_SCATTER_ITEMS 8
template <bool SMEM> __global__ void __launch_bounds__(256) k_bins_index(const double *__restrict__ vals, const HistCols cols, int64_t n,
																		   const HistSpec hs, int32_t *idx_out, unsigned long long *counts)
{
	__shared__ uint32_t sh_cnt[SMEM ? MG_BINS_SCATTER_SMEM_MAX : 1];
	if (SMEM) {
		for (int b = threadIdx.x; b < (int)hs.total; b += blockDim.x)
			sh_cnt[b] = 0;
		__syncthreads();
	}
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		long long idx = -1;
		const float v = (float)__ldcs(vals + i);
		if (v == v)
			idx = mg_vals2idx(cols, hs, i);
		idx_out[i] = (int32_t)idx;
		if (idx >= 0) {
			if (SMEM)
				atomicAdd(&sh_cnt[idx], 1u); // a thread block sees fewer than 2^32 rows
			else
				atomicAdd(counts + idx, 1ull);
		}
	}
	if (SMEM) {
		__syncthreads();
		for (int b = threadIdx.x; b < (int)hs.total; b += blockDim.x)
			if (sh_cnt[b])
				atomicAdd(counts + b, (unsigned long long)sh_cnt[b]);
	}
}

// step 2: scatter the values next to their bin's others (order inside a bin is irrelevant: the segment gets selected / sorted).
// SMEM: a thread block takes tiles of 256 x 8 rows; a shared-memory atomic gives every row its rank among the tile's rows of its
// bin, one global atomic per (tile, bin) reserves the range, the rows then store. Otherwise one global atomic per row.
template <bool SMEM> __global__ void __launch_bounds__(256) k_bins_scatter(const double *__restrict__ vals, const int32_t *__restrict__ idx_in,
																			 int64_t n, long long total, unsigned long long *cursor, double *out)
{
	__shared__ uint32_t sh_cnt[SMEM ? MG_BINS_SCATTER_SMEM_MAX : 1];
	__shared__ unsigned long long sh_base[SMEM ? MG_BINS_SCATTER_SMEM_MAX : 1];
	if (!SMEM) {
		const int64_t stride = (int64_t)gridDim.x * blockDim.x;
		for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
			const int idx = idx_in[i];
			if (idx >= 0)
				out[atomicAdd(cursor + idx, 1ull)] = vals[i];
		}
		return;
	}
	const int64_t tile = (int64_t)blockDim.x * MG_BINS_SCATTER_ITEMS;
	for (int64_t base = (int64_t)blockIdx.x * tile; base < n; base += (int64_t)gridDim.x * tile) {
		for (int b = threadIdx.x; b < (int)total; b += blockDim.x)
			sh_cnt[b] = 0;
		__syncthreads();
		int idx[MG_BINS_SCATTER_ITEMS];
		uint32_t rank[MG_BINS_SCATTER_ITEMS];
		double v[MG_BINS_SCATTER_ITEMS];
#pragma unroll
		for (int j = 0; j < MG_BINS_SCATTER_ITEMS; ++j) {
			const int64_t i = base + (int64_t)j * blockDim.x + threadIdx.x;
			idx[j] = i < n ? idx_in[i] : -1;
			v[j] = i < n ? __ldcs(vals + i) : 0.;
		}
#pragma unroll
		for (int j = 0; j < MG_BINS_SCATTER_ITEMS; ++j)
			rank[j] = idx[j] >= 0 ? atomicAdd(&sh_cnt[idx[j]], 1u) : 0u;
		__syncthreads();
		for (int b = threadIdx.x; b < (int)total; b += blockDim.x)
			if (sh_cnt[b])
				sh_base[b] = atomicAdd(cursor + b, (unsigned long long)sh_cnt[b]);
		__syncthreads();
#pragma unroll
		for (int j = 0; j < MG_BINS_SCATTER_ITEMS; ++j)
			if (idx[j] >= 0)
				out[sh_base[idx[j]] + rank[j]] = v[j];
		__syncthreads();
	}
}

// ---- global.percentile / global.percentile.min / global.percentile.max ---------------------------------------------------------
// The vtrack's avg / min / max value is replaced by the track's precomputed percentile of it (src/TrackVarProcessor.cpp:98-110):
// bins[val2bin(v)]; a value at or below the first break gets bins[0], one above the last break 1; NaN stays NaN.
__global__ void __launch_bounds__(256) k_percentile_lookup(double *col, int64_t n, const HistSpec hs, const double *__restrict__ bins)
{
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		const double v = col[i];
		if (isnan(v))
			continue;
		const int b = mg_val2bin(hs.breaks, hs.nbreaks[0], hs.binsize[0], 0, v);
		col[i] = b >= 0 ? __ldg(bins + b) : (v <= hs.breaks[0] ? __ldg(bins) : 1.);
	}
}

// ---- gscreen comparison predicates on the device -----------------------------------------------------------------------------
struct PredSpec {
	int nterms, combine_or;
	const double *col[MG_PRED_MAX_TERMS];
	int op[MG_PRED_MAX_TERMS];
	double thr[MG_PRED_MAX_TERMS];
};

__device__ __forceinline__ bool mg_cmp(int op, double v, double t)
{
	switch (op) { // every comparison with a NaN operand is false (R: NA, which never counts as TRUE)
	case MG_CMP_GT: return v > t;
	case MG_CMP_GE: return v >= t;
	case MG_CMP_LT: return v < t;
	case MG_CMP_LE: return v <= t;
	case MG_CMP_EQ: return v == t;
	default: return v < t || v > t;
	}
}

// 4 rows per thread: two 16-byte loads per column, one 4-byte store of flags
__global__ void __launch_bounds__(256) k_pred_flags(const PredSpec p, int64_t n, int8_t *__restrict__ flag)
{
	const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
	if (i0 >= n)
		return;
	bool r[4];
#pragma unroll
	for (int j = 0; j < 4; ++j)
		r[j] = p.combine_or == 0;
	const bool full = i0 + 4 <= n;
	for (int k = 0; k < p.nterms; ++k) {
		double v[4];
		if (full && (reinterpret_cast<uintptr_t>(p.col[k]) & 15) == 0) {
			const double2 a = __ldcs(reinterpret_cast<const double2 *>(p.col[k] + i0)), b = __ldcs(reinterpret_cast<const double2 *>(p.col[k] + i0 + 2));
			v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
		} else {
#pragma unroll
			for (int j = 0; j < 4; ++j)
				v[j] = i0 + j < n ? __ldcs(p.col[k] + i0 + j) : mg_nan();
		}
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			const bool c = mg_cmp(p.op[k], v[j], p.thr[k]);
			r[j] = p.combine_or ? (r[j] || c) : (r[j] && c);
		}
	}
	if (full && (reinterpret_cast<uintptr_t>(flag) & 3) == 0)
		*reinterpret_cast<char4 *>(flag + i0) = make_char4(r[0], r[1], r[2], r[3]);
	else
		for (int j = 0; j < 4 && i0 + j < n; ++j)
			flag[i0 + j] = r[j];
}

// ---- gscreen run merge ---------------------------------------------------------------------------------------------------
// A TRUE row continues the current output interval iff the previous row is TRUE, on the same chromosome, and ends where
// this one starts (src/GenomeTrackScreener.cpp:203-211); otherwise it is a run HEAD. A TRUE row whose successor does not
// continue it is a run TAIL. The k-th head and the k-th tail delimit the k-th output interval, so the merge is two
// order-preserving compactions (start/chrom from heads, end from tails) over one block-scan.
#define MG_SCR_THREADS 256

__device__ __forceinline__ void mg_screen_classify(const RowSrc &rs, int64_t first, int64_t count, const int8_t *flag, int64_t i, bool &head,
												   bool &tail, int &c, int64_t &s, int64_t &e)
{
	int64_t k;
	head = tail = false;
	if (flag[i] != 1)
		return;
	mg_get_row(rs, first + i, c, s, e, k);
	head = true;
	if (i > 0 && flag[i - 1] == 1) {
		int pc;
		int64_t ps, pe;
		mg_get_row(rs, first + i - 1, pc, ps, pe, k);
		head = !(pc == c && pe == s);
	}
	tail = true;
	if (i + 1 < count && flag[i + 1] == 1) {
		int nc;
		int64_t ns, ne;
		mg_get_row(rs, first + i + 1, nc, ns, ne, k);
		tail = !(nc == c && ns == e);
	}
}

__global__ void __launch_bounds__(MG_SCR_THREADS) k_screen_count(const RowSrc rs, int64_t first, int64_t count, const int8_t *flag,
																  long long *block_heads, long long *block_tails)
{
	__shared__ long long sc[33];
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	bool head = false, tail = false;
	int c;
	int64_t s, e;
	if (i < count)
		mg_screen_classify(rs, first, count, flag, i, head, tail, c, s, e);
	long long th, tt;
	mg_block_excl_scan<long long>(head ? 1 : 0, sc, th);
	mg_block_excl_scan<long long>(tail ? 1 : 0, sc, tt);
	if (threadIdx.x == 0) {
		block_heads[blockIdx.x] = th;
		block_tails[blockIdx.x] = tt;
	}
}

// ---- the fixed-bin iterator's implicit rows: 16 rows per thread, heads and tails as bit masks -----------------------------------
// Consecutive implicit rows of one scope interval always touch (row r ends where row r + 1 starts), so inside a scope interval a
// head is a TRUE row after a row that is not TRUE and a tail a TRUE row before one that is not: two shifts and an AND over 16 flags
// loaded at once. Only the few 16-row chunks that hold a scope interval's first row look at coordinates (two scope intervals may
// touch, then the run goes on: src/GenomeTrackScreener.cpp:203-211 compares chromosome and coordinates, nothing else).
#define MG_SCB_ROWS 16
#define MG_SCB_THREADS 256

__device__ __forceinline__ uint32_t mg_true_bits4(uint32_t w) // bit j = byte j of w equals 1
{
	const uint32_t m = __vcmpeq4(w, 0x01010101u) & 0x01010101u;
	return (m * 0x01020408u) >> 24;
}

// heads / tails (bit j = row i0 + j) of the 16 rows from i0 (a multiple of 16); rc: the scope interval of row i0
__device__ __forceinline__ void mg_screen_bits(const RowSrc &rs, int64_t first, int64_t count, const int8_t *__restrict__ flag, int64_t i0, uint32_t &heads,
											   uint32_t &tails, RowCache &rc)
{
	const int64_t left = count - i0;
	uint32_t bits = 0;
	if (left >= 16) {
		const uint4 f = __ldg(reinterpret_cast<const uint4 *>(flag + i0));
		bits = mg_true_bits4(f.x) | (mg_true_bits4(f.y) << 4) | (mg_true_bits4(f.z) << 8) | (mg_true_bits4(f.w) << 12);
	} else { // the last rows: nothing is read past the array
		for (int j = 0; j < (int)left; ++j)
			bits |= (flag[i0 + j] == 1 ? 1u : 0u) << j;
	}
	const uint32_t prev = i0 > 0 && flag[i0 - 1] == 1 ? 1u : 0u, next = left > 16 && flag[i0 + 16] == 1 ? 1u : 0u;
	const uint32_t w = prev | (bits << 1) | (next << 17); // bit j + 1 = row i0 + j
	uint32_t brk = 0;                                      // bit j: rows i0 + j - 1 and i0 + j do not touch (j = 0 .. 16)
	if (bits) {
		const int64_t r0 = first + i0;
		mg_rowcache_load(rs, r0, rc);
		if (r0 == rc.row0 || r0 + 16 >= rc.row1) { // a scope interval starts inside rows i0 .. i0 + 16: look at the coordinates there
			RowCache c2;
			for (int j = 0; j <= 16; ++j) {
				const int64_t r = r0 + j;
				if (r < 1 || r >= rs.n || i0 + j < 1 || i0 + j >= count)
					continue;
				int ca, cb;
				int64_t sa, ea, sb, eb, k;
				mg_get_row_cached(rs, r - 1, c2, ca, sa, ea, k);
				mg_get_row_cached(rs, r, c2, cb, sb, eb, k);
				if (!(ca == cb && ea == sb))
					brk |= 1u << j;
			}
		}
	}
	heads = (w >> 1) & (~w | brk) & 0xffffu;
	tails = (w >> 1) & (~(w >> 2) | (brk >> 1)) & 0xffffu;
}

__global__ void __launch_bounds__(MG_SCB_THREADS) k_screen_count_fb(const RowSrc rs, int64_t first, int64_t count, const int8_t *flag, long long *block_heads,
																	 long long *block_tails)
{
	__shared__ int sc[33];
	const int64_t i0 = ((int64_t)blockIdx.x * MG_SCB_THREADS + threadIdx.x) * MG_SCB_ROWS;
	uint32_t heads = 0, tails = 0;
	RowCache rc;
	if (i0 < count)
		mg_screen_bits(rs, first, count, flag, i0, heads, tails, rc);
	int th;
	mg_block_excl_scan<int>(__popc(heads) | (__popc(tails) << 16), sc, th); // both counts in one scan: a block holds 4096 rows
	if (threadIdx.x == 0) {
		block_heads[blockIdx.x] = th & 0xffff;
		block_tails[blockIdx.x] = th >> 16;
	}
}

__global__ void __launch_bounds__(MG_SCB_THREADS) k_screen_write_fb(const RowSrc rs, int64_t first, int64_t count, const int8_t *flag,
																	 const long long *block_heads, const long long *block_tails, int32_t *ochrom,
																	 int64_t *ostart, int64_t *oend)
{
	__shared__ int sc[33];
	const int64_t i0 = ((int64_t)blockIdx.x * MG_SCB_THREADS + threadIdx.x) * MG_SCB_ROWS;
	uint32_t heads = 0, tails = 0;
	RowCache rc;
	if (i0 < count)
		mg_screen_bits(rs, first, count, flag, i0, heads, tails, rc);
	int tot;
	const int off = mg_block_excl_scan<int>(__popc(heads) | (__popc(tails) << 16), sc, tot);
	long long hoff = block_heads[blockIdx.x] + (off & 0xffff), toff = block_tails[blockIdx.x] + (off >> 16);
	for (uint32_t m = heads | tails; m; m &= m - 1u) {
		const int j = __ffs(m) - 1;
		int c;
		int64_t s, e, k;
		mg_get_row_cached(rs, first + i0 + j, rc, c, s, e, k);
		if ((heads >> j) & 1u) {
			ochrom[hoff] = c;
			ostart[hoff] = s;
			++hoff;
		}
		if ((tails >> j) & 1u)
			oend[toff++] = e;
	}
}

__global__ void __launch_bounds__(1024) k_screen_scan(long long *block_heads, long long *block_tails, int64_t nblocks, long long *total_out)
{
	__shared__ long long sc[33];
	const int64_t chunk = (nblocks + blockDim.x - 1) / blockDim.x;
	const int64_t lo = (int64_t)threadIdx.x * chunk, hi = min(lo + chunk, nblocks);
	long long sh = 0, st = 0;
	for (int64_t i = lo; i < hi; ++i) {
		sh += block_heads[i];
		st += block_tails[i];
	}
	long long total_h, total_t;
	long long h0 = mg_block_excl_scan<long long>(sh, sc, total_h);
	long long t0 = mg_block_excl_scan<long long>(st, sc, total_t);
	for (int64_t i = lo; i < hi; ++i) {
		const long long vh = block_heads[i], vt = block_tails[i];
		block_heads[i] = h0;
		block_tails[i] = t0;
		h0 += vh;
		t0 += vt;
	}
	if (threadIdx.x == 0)
		*total_out = total_h;
}

__global__ void __launch_bounds__(MG_SCR_THREADS) k_screen_write(const RowSrc rs, int64_t first, int64_t count, const int8_t *flag,
																  const long long *block_heads, const long long *block_tails, int32_t *ochrom,
																  int64_t *ostart, int64_t *oend)
{
	__shared__ long long sc[33];
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	bool head = false, tail = false;
	int c = 0;
	int64_t s = 0, e = 0;
	if (i < count)
		mg_screen_classify(rs, first, count, flag, i, head, tail, c, s, e);
	long long th, tt;
	const long long hoff = block_heads[blockIdx.x] + mg_block_excl_scan<long long>(head ? 1 : 0, sc, th);
	const long long toff = block_tails[blockIdx.x] + mg_block_excl_scan<long long>(tail ? 1 : 0, sc, tt);
	if (head) {
		ochrom[hoff] = c;
		ostart[hoff] = s;
	}
	if (tail)
		oend[toff] = e;
}
