// dense_blockwm.cuh -- block-local wavelet matrices: the quantile index for windows of up to MG_BW_S bins.
//
// The chromosome is cut into blocks of B = 2^LB bins that overlap by half (block b covers bins [b*S, b*S + B), S = B/2),
// so every window of <= S bins lies inside block floor(sbin / S). Inside a block the bins are ranked 0..B-1 by
// (value, position) with NaN (and the padding past the chromosome end) ranked last; the ranks are a permutation of
// [0, B), therefore
//   * a wavelet matrix over the ranks has LB levels and every level has exactly B/2 zeros -- no per-level constant;
//   * the k-th smallest non-NaN value of a window is the k-th smallest RANK of the window's positions (k < n, n = the
//     window's non-NaN count from the prefix entries) -- NaNs never need compacting;
//   * the (k+1)-th is the next larger rank whose position lies in the window: a short scan of pos_of_rank.
// A block record is LB levels of {uint64 bits[B/64], uint16 ones_before_word[B/64]} (10 bytes per 64 bins: 320 bytes a level
// for B = 2048), one more such level marking the NaN bins (a quantile-only query gets its non-NaN count from it: two
// lookups instead of the prefix entries and two sectors of bins), and pos_of_rank[B] (uint16).
// Rows sorted by position walk the same 3.8 KB of levels as their neighbours. The row-query kernel stages the levels of the
// records its 256 rows need in shared memory with one bulk copy each (cp.async.bulk + mbarrier), so the LB dependent steps
// of a walk are shared-memory reads; a row whose record is not among the staged ones walks the same bytes through L1. The
// value of a rank is read from the track itself: vals[b*S + pos_of_rank[rank]].
#pragma once
#include "common.cuh"
#include "dense.cuh"

#ifndef MG_BW_LB
#define MG_BW_LB 11
#endif
#define MG_BW_B (1 << MG_BW_LB)
#ifndef MG_BW_STRIDE
#define MG_BW_STRIDE (MG_BW_B / 2) // block b covers bins [b * STRIDE, b * STRIDE + B)
#endif
#define MG_BW_S (MG_BW_B - MG_BW_STRIDE) // longest window that is guaranteed to lie inside one block
#define MG_BW_HALF (MG_BW_B / 2)         // zeros of every level (the ranks are a permutation of [0, B))
#define MG_BW_UNITS (MG_BW_B / 32)                                    // 32-bit ballots per level (build only)
#define MG_BW_WORDS (MG_BW_B / 64)                                    // 64-bit words per level
#define MG_BW_CNT_OFFSET (MG_BW_WORDS * 8)                            // uint16 ones before each word, after the words
#define MG_BW_LEVEL_BYTES (MG_BW_WORDS * 10)
#define MG_BW_NAN_OFFSET (MG_BW_LB * MG_BW_LEVEL_BYTES)              // one more level-shaped bitmap: bit = the bin is NaN
#define MG_BW_POS_OFFSET (MG_BW_NAN_OFFSET + MG_BW_LEVEL_BYTES)        // pos_of_rank[B] (uint16)
#define MG_BW_STAGE_BYTES MG_BW_POS_OFFSET                            // what a staged record holds: the levels and the NaN level
static_assert(MG_BW_LEVEL_BYTES % 16 == 0 && MG_BW_B <= 65536, "levels must stay 16-byte multiples (bulk copies) with 16-bit counts");
#define MG_BW_RECORD_BYTES (MG_BW_POS_OFFSET + MG_BW_B * 2)
#define MG_BW_BUILD_THREADS 1024
#define MG_BW_PER_THREAD (MG_BW_B / MG_BW_BUILD_THREADS)

__device__ __forceinline__ uint32_t mg_bw_key(float v)
{
	if (isnan(v))
		return 0xffffffffu;
	const uint32_t b = __float_as_uint(v);
	return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}

// one CTA per block: rank the block's bins (bitonic sort of (key, position)), then build the LB levels in shared memory
__global__ void __launch_bounds__(MG_BW_BUILD_THREADS) k_bw_build(const float *__restrict__ vals, int64_t nbins, uint8_t *__restrict__ records)
{
	__shared__ unsigned long long skey[MG_BW_B];
	__shared__ uint32_t sbits[MG_BW_UNITS], scnt[MG_BW_UNITS];
	uint16_t *sa = reinterpret_cast<uint16_t *>(skey), *sb = sa + MG_BW_B; // reuse the sort buffer once the ranks are known
	const int t = threadIdx.x;
	const int64_t base = (int64_t)blockIdx.x * MG_BW_STRIDE;
	uint8_t *rec = records + (size_t)blockIdx.x * MG_BW_RECORD_BYTES;
#pragma unroll
	for (int j = 0; j < MG_BW_PER_THREAD; ++j) {
		const int e = t + j * MG_BW_BUILD_THREADS;
		const int64_t i = base + e;
		const float v = i < nbins ? vals[i] : mg_nanf();
		skey[e] = ((unsigned long long)mg_bw_key(v) << 32) | (unsigned)e;
		const uint32_t nanbits = __ballot_sync(0xffffffffu, v != v);
		if ((t & 31) == 0)
			sbits[e >> 5] = nanbits;
	}
	__syncthreads();
	if (t < MG_BW_WORDS) { // the NaN level: few words, a serial count per word is fine
		uint32_t before = 0;
		for (int u = 0; u < 2 * t; ++u)
			before += __popc(sbits[u]);
		reinterpret_cast<unsigned long long *>(rec + MG_BW_NAN_OFFSET)[t] = ((unsigned long long)sbits[2 * t + 1] << 32) | sbits[2 * t];
		reinterpret_cast<uint16_t *>(rec + MG_BW_NAN_OFFSET + MG_BW_CNT_OFFSET)[t] = (uint16_t)before;
	}
	__syncthreads();
	for (int k = 2; k <= MG_BW_B; k <<= 1) {
		for (int j = k >> 1; j > 0; j >>= 1) {
			for (int p = t; p < MG_BW_B / 2; p += MG_BW_BUILD_THREADS) {
				const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1)), ixj = i | j;
				const unsigned long long a = skey[i], b = skey[ixj];
				if ((a > b) == ((i & k) == 0)) {
					skey[i] = b;
					skey[ixj] = a;
				}
			}
			__syncthreads();
		}
	}
	uint16_t *pos_of_rank = reinterpret_cast<uint16_t *>(rec + MG_BW_POS_OFFSET);
	uint16_t *cur = sa, *nxt = sb;
	uint32_t mypos[MG_BW_PER_THREAD];
#pragma unroll
	for (int j = 0; j < MG_BW_PER_THREAD; ++j) {
		const int r = t + j * MG_BW_BUILD_THREADS;
		mypos[j] = (uint32_t)skey[r];
		pos_of_rank[r] = (uint16_t)mypos[j];
	}
	__syncthreads(); // every key has been read: the buffer becomes the two rank arrays
#pragma unroll
	for (int j = 0; j < MG_BW_PER_THREAD; ++j)
		cur[mypos[j]] = (uint16_t)(t + j * MG_BW_BUILD_THREADS);
	__syncthreads();
	uint8_t *level = rec;
	for (int lv = 0; lv < MG_BW_LB; ++lv, level += MG_BW_LEVEL_BYTES) {
		const int shift = MG_BW_LB - 1 - lv;
		uint32_t myword[MG_BW_PER_THREAD];
#pragma unroll
		for (int j = 0; j < MG_BW_PER_THREAD; ++j) {
			const int e = t + j * MG_BW_BUILD_THREADS;
			myword[j] = __ballot_sync(0xffffffffu, (cur[e] >> shift) & 1);
			if ((t & 31) == 0)
				sbits[e >> 5] = myword[j];
		}
		__syncthreads();
		if (t < 32) { // exclusive scan of the MG_BW_UNITS word popcounts by one warp
			uint32_t c[MG_BW_UNITS / 32], s = 0;
#pragma unroll
			for (int j = 0; j < MG_BW_UNITS / 32; ++j) {
				c[j] = __popc(sbits[t * (MG_BW_UNITS / 32) + j]);
				s += c[j];
			}
			uint32_t incl = s;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
				if (t >= d)
					incl += o;
			}
			uint32_t run = incl - s;
#pragma unroll
			for (int j = 0; j < MG_BW_UNITS / 32; ++j) {
				scnt[t * (MG_BW_UNITS / 32) + j] = run;
				run += c[j];
			}
		}
		__syncthreads();
#pragma unroll
		for (int j = 0; j < MG_BW_PER_THREAD; ++j) {
			const int e = t + j * MG_BW_BUILD_THREADS;
			const uint32_t bits = myword[j], lane = t & 31;
			const uint32_t ones_before = scnt[e >> 5] + __popc(bits & ((1u << lane) - 1u));
			const uint16_t v = cur[e];
			nxt[((bits >> lane) & 1u) ? MG_BW_HALF + ones_before : (uint32_t)e - ones_before] = v;
			if (lane == 0) { // little-endian: ballot 2w is the low half of word w
				reinterpret_cast<uint32_t *>(level)[e >> 5] = bits;
				if (!((e >> 5) & 1))
					reinterpret_cast<uint16_t *>(level + MG_BW_CNT_OFFSET)[e >> 6] = (uint16_t)scnt[e >> 5];
			}
		}
		__syncthreads();
		uint16_t *tmp = cur;
		cur = nxt;
		nxt = tmp;
	}
}

// G: the record is read from global memory through the read-only path (true) or from its shared-memory copy (false)
template <bool G, typename T>
__device__ __forceinline__ T mg_bw_ld(const T *p)
{
	return G ? __ldg(p) : *p;
}
// ones among positions [0, i) of a level, 0 <= i < B
template <bool G>
__device__ __forceinline__ uint32_t mg_bw_rank_lo(const uint8_t *__restrict__ lv, uint32_t i)
{
	const uint32_t w = i >> 6;
	const unsigned long long bits = mg_bw_ld<G>(reinterpret_cast<const unsigned long long *>(lv) + w);
	const uint32_t c = mg_bw_ld<G>(reinterpret_cast<const uint16_t *>(lv + MG_BW_CNT_OFFSET) + w);
	return c + __popcll(bits & ((1ull << (i & 63u)) - 1ull));
}
// ones among positions [0, i) of a level, 1 <= i <= B (looks at the word of position i - 1, so i == B needs no padding word)
template <bool G>
__device__ __forceinline__ uint32_t mg_bw_rank_hi(const uint8_t *__restrict__ lv, uint32_t i)
{
	const uint32_t j = i - 1u, w = j >> 6;
	const unsigned long long bits = mg_bw_ld<G>(reinterpret_cast<const unsigned long long *>(lv) + w);
	const uint32_t c = mg_bw_ld<G>(reinterpret_cast<const uint16_t *>(lv + MG_BW_CNT_OFFSET) + w);
	return c + __popcll(bits << (~j & 63u));
}

// non-NaN bins of the block-local window [l, r), r > l; lv = the record's levels (global or staged)
template <bool G>
__device__ __forceinline__ uint32_t mg_bw_count(const uint8_t *__restrict__ lv, uint32_t l, uint32_t r)
{
	lv += MG_BW_NAN_OFFSET;
	return (r - l) - (mg_bw_rank_hi<G>(lv, r) - mg_bw_rank_lo<G>(lv, l));
}

// k-th smallest (0-based) non-NaN value of the block-local window [l, r), k < n <= r - l; with want2 (k + 1 < n) also the
// (k+1)-th. lv = the record's levels (global or staged), rec = the record in global memory (pos_of_rank is read there),
// vals_block = the track at the block's first bin.
#ifndef MG_BW_SELECT_INLINE
#define MG_BW_SELECT_INLINE __forceinline__
#endif
#ifndef MG_BW_UNROLL
#define MG_BW_UNROLL MG_BW_LB
#endif
template <bool G>
__device__ MG_BW_SELECT_INLINE void mg_bw_select(const uint8_t *__restrict__ lv, const uint8_t *__restrict__ rec, const float *__restrict__ vals_block, uint32_t l,
											 uint32_t r, uint32_t k, bool want2, float &x1, float &x2)
{
	const uint32_t l0 = l, len = r - l;
	uint32_t rank = 0;
	constexpr int unroll = MG_BW_UNROLL;
#pragma unroll unroll
	for (int level = 0; level < MG_BW_LB; ++level, lv += MG_BW_LEVEL_BYTES) {
		const uint32_t ol = mg_bw_rank_lo<G>(lv, l), orr = mg_bw_rank_hi<G>(lv, r);
		const uint32_t nz = (r - l) - (orr - ol);
		const bool one = k >= nz;
		l = one ? MG_BW_HALF + ol : l - ol;
		r = one ? MG_BW_HALF + orr : r - orr;
		k = one ? k - nz : k;
		rank = (rank << 1) | (one ? 1u : 0u);
	}
	const uint16_t *pos_of_rank = reinterpret_cast<const uint16_t *>(rec + MG_BW_POS_OFFSET);
	x1 = __ldg(vals_block + __ldg(pos_of_rank + rank));
	x2 = x1;
	if (want2) {
		// the next larger rank whose position lies in the window; exists because k + 1 < n
		uint32_t p;
		do {
			++rank;
			p = __ldg(pos_of_rank + rank);
		} while (p - l0 >= len && rank < MG_BW_B - 1);
		x2 = __ldg(vals_block + p);
	}
}

// ---- staging a record's levels in shared memory: one bulk copy per record, completion on an mbarrier ---------------------
__device__ __forceinline__ uint32_t mg_smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mg_mbar_init(uint64_t *bar, uint32_t arrivals)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mg_smem_addr(bar)), "r"(arrivals) : "memory");
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mg_mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mg_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mg_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(mg_smem_addr(dst)), "l"(src), "r"(bytes),
				 "r"(mg_smem_addr(bar))
				 : "memory");
}
// a waiting thread is suspended by the hardware for up to this long per try instead of spinning on the barrier word
#ifndef MG_MBAR_SUSPEND_NS
#define MG_MBAR_SUSPEND_NS 100000u
#endif
__device__ __forceinline__ void mg_mbar_wait(uint64_t *bar, uint32_t parity)
{
	asm volatile("{\n"
				 ".reg .pred p;\n"
				 "MG_WAIT_%=:\n"
				 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
				 "@p bra MG_DONE_%=;\n"
				 "bra MG_WAIT_%=;\n"
				 "MG_DONE_%=:\n"
				 "}" ::"r"(mg_smem_addr(bar)),
				 "r"(parity), "r"(MG_MBAR_SUSPEND_NS)
				 : "memory");
}
