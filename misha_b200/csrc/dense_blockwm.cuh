// dense_blockwm.cuh -- block-local wavelet matrices: the quantile index for windows of up to MG_BW_S bins.
//
// The chromosome is cut into blocks of B = 2^LB bins that overlap by half (block b covers bins [b*S, b*S + B), S = B/2),
// so every window of <= S bins lies inside block floor(sbin / S). Inside a block the bins are ranked 0..B-1 by
// (value, position) with NaN (and the padding past the chromosome end) ranked last; the ranks are a permutation of
// [0, B), therefore
//   * a wavelet matrix over the ranks has every level split exactly in half (B/2 zeros) -- no per-level constant;
//   * the k-th smallest non-NaN value of a window is the k-th smallest RANK of the window's positions (k < n, n = the
//     window's non-NaN count) -- NaNs never need compacting.
// Only the TOP MG_BW_WL = LB - 4 levels are stored and walked: they resolve the k-th smallest rank to its bucket of 16 consecutive
// ranks and the number k' of window members of that bucket below it. The rest is one look at pos_of_rank: the positions of the
// bucket's 16 ranks and of the next bucket's are 64 contiguous bytes; a 32-bit membership mask (position inside the window?)
// gives the (k'+1)-th member = the k-th smallest, and the next set bit = the (k+1)-th smallest the interpolation of
// src/StreamPercentiler.h:210-217 needs. That replaces 4 more levels (8 dependent rank queries) and a separate successor search.
//
// A block record: WL levels + one level-shaped NaN bitmap, each {uint32 bits[B/32 + 1], uint16 ones_before_word[B/32 + 1]} padded
// to 16 bytes (the extra unit serves rank(B): a walk's upper end reaches B) -- a rank query is a 4-byte and a 2-byte load at
// offsets derived from one index, straight from the copy a thread block stages in shared memory (no re-layout pass) or from L1;
// then pos_of_rank[B + 16] (uint16; the tail holds B - 1, a position no window of <= S bins starting below S reaches). Inside every
// group of 16 ranks the entries are interleaved -- halfword 2j holds rank j, halfword 2j + 1 rank j + 8 (mg_bw_pos_slot) -- so that
// the membership test of two positions packed in one 32-bit word leaves its two result bits 16 apart in rank order.
// The value of a rank is read from the track itself (or its staged copy): vals[b*S + pos_of_rank[rank]].
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
#define MG_BW_BUCKET_BITS 4
#define MG_BW_BUCKET (1 << MG_BW_BUCKET_BITS)                            // ranks resolved by the pos_of_rank scan
#define MG_BW_WL (MG_BW_LB - MG_BW_BUCKET_BITS)                          // levels stored and walked
#define MG_BW_UNITS (MG_BW_B / 32)                                      // 32-bit words per level
#define MG_BW_CNT_OFFSET ((MG_BW_UNITS + 1) * 4)                         // uint16 ones before each word, after the words
#define MG_BW_LEVEL_BYTES ((((MG_BW_UNITS + 1) * 6) + 15) & ~15)
#define MG_BW_NAN_OFFSET (MG_BW_WL * MG_BW_LEVEL_BYTES)                 // one more level-shaped bitmap: bit = the bin is NaN
#define MG_BW_POS_OFFSET (MG_BW_NAN_OFFSET + MG_BW_LEVEL_BYTES)           // pos_of_rank[B + 16] (uint16)
#define MG_BW_STAGE_BYTES MG_BW_NAN_OFFSET                               // what a staged record holds: the walked levels
#define MG_BW_POS_PAD 16
static_assert(MG_BW_LEVEL_BYTES % 16 == 0 && MG_BW_LB <= 14, "levels must stay 16-byte multiples (bulk copies); packed position tests need B <= 2^14");
static_assert(MG_BW_STRIDE + MG_BW_S <= MG_BW_B && MG_BW_S <= MG_BW_B / 2, "window geometry");
#define MG_BW_RECORD_BYTES (MG_BW_POS_OFFSET + (MG_BW_B + MG_BW_POS_PAD) * 2)
static_assert(MG_BW_RECORD_BYTES % 16 == 0, "records must stay 16-byte aligned");
#define MG_BW_BUILD_THREADS 1024
#define MG_BW_PER_THREAD (MG_BW_B / MG_BW_BUILD_THREADS)

// where pos_of_rank keeps rank r
__device__ __forceinline__ uint32_t mg_bw_pos_slot(uint32_t r) { return (r & ~15u) | ((r & 7u) << 1) | ((r >> 3) & 1u); }

__device__ __forceinline__ uint32_t mg_bw_key(float v)
{
	if (isnan(v))
		return 0xffffffffu;
	const uint32_t b = __float_as_uint(v);
	return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}

// one level's words and counts from the ballots in sbits[] / their exclusive prefix in scnt[]
__device__ __forceinline__ void mg_bw_store_level(uint8_t *level, const uint32_t *sbits, const uint32_t *scnt, int t)
{
	if (t < MG_BW_UNITS) {
		reinterpret_cast<uint32_t *>(level)[t] = sbits[t];
		reinterpret_cast<uint16_t *>(level + MG_BW_CNT_OFFSET)[t] = (uint16_t)scnt[t];
	} else if (t == MG_BW_UNITS) { // the unit of rank(B): every bit below it is counted
		reinterpret_cast<uint32_t *>(level)[t] = 0u;
		reinterpret_cast<uint16_t *>(level + MG_BW_CNT_OFFSET)[t] = (uint16_t)(scnt[MG_BW_UNITS - 1] + __popc(sbits[MG_BW_UNITS - 1]));
	}
}

// exclusive scan of the MG_BW_UNITS word popcounts by one warp
__device__ __forceinline__ void mg_bw_scan_units(const uint32_t *sbits, uint32_t *scnt, int t)
{
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

// one CTA per block: rank the block's bins (bitonic sort of (key, position)), then build the walked levels in shared memory
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
	if (t < 32)
		mg_bw_scan_units(sbits, scnt, t);
	__syncthreads();
	mg_bw_store_level(rec + MG_BW_NAN_OFFSET, sbits, scnt, t);
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
		pos_of_rank[mg_bw_pos_slot((uint32_t)r)] = (uint16_t)mypos[j];
	}
	if (t < MG_BW_POS_PAD)
		pos_of_rank[MG_BW_B + t] = (uint16_t)(MG_BW_B - 1);
	__syncthreads(); // every key has been read: the buffer becomes the two rank arrays
#pragma unroll
	for (int j = 0; j < MG_BW_PER_THREAD; ++j)
		cur[mypos[j]] = (uint16_t)(t + j * MG_BW_BUILD_THREADS);
	__syncthreads();
	uint8_t *level = rec;
	for (int lv = 0; lv < MG_BW_WL; ++lv, level += MG_BW_LEVEL_BYTES) {
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
		if (t < 32)
			mg_bw_scan_units(sbits, scnt, t);
		__syncthreads();
		mg_bw_store_level(level, sbits, scnt, t);
#pragma unroll
		for (int j = 0; j < MG_BW_PER_THREAD; ++j) {
			const int e = t + j * MG_BW_BUILD_THREADS;
			const uint32_t bits = myword[j], lane = t & 31;
			const uint32_t ones_before = scnt[e >> 5] + __popc(bits & ((1u << lane) - 1u));
			nxt[((bits >> lane) & 1u) ? MG_BW_HALF + ones_before : (uint32_t)e - ones_before] = cur[e];
		}
		__syncthreads();
		uint16_t *tmp = cur;
		cur = nxt;
		nxt = tmp;
	}
}

// G: the levels are read from global memory through the read-only path (true) or from their shared-memory copy (false)
template <bool G, typename T>
__device__ __forceinline__ T mg_bw_ld(const T *p)
{
	return G ? __ldg(p) : *p;
}
// ones among positions [0, i) of a level, 0 <= i <= B
template <bool G>
__device__ __forceinline__ uint32_t mg_bw_rank(const uint8_t *__restrict__ lv, uint32_t i)
{
	const uint32_t w = i >> 5;
	const uint32_t bits = mg_bw_ld<G>(reinterpret_cast<const uint32_t *>(lv) + w);
	const uint32_t c = mg_bw_ld<G>(reinterpret_cast<const uint16_t *>(lv + MG_BW_CNT_OFFSET) + w);
	return c + __popc(bits & ((1u << (i & 31u)) - 1u));
}

// non-NaN bins of the block-local window [l, r), r > l; rec = the record (its NaN level is read in place)
__device__ __forceinline__ uint32_t mg_bw_count(const uint8_t *__restrict__ rec, uint32_t l, uint32_t r)
{
	const uint8_t *lv = rec + MG_BW_NAN_OFFSET;
	return (r - l) - (mg_bw_rank<true>(lv, r) - mg_bw_rank<true>(lv, l));
}

// the walk: the bucket of MG_BW_BUCKET consecutive ranks that holds the k-th smallest rank of the window (returned: its first rank) and how
// many of the window's members in that bucket lie below it (k on return)
template <bool G>
__device__ __forceinline__ uint32_t mg_bw_walk(const uint8_t *__restrict__ lv, uint32_t l, uint32_t r, uint32_t &k)
{
	uint32_t bucket = 0;
#pragma unroll
	for (int level = 0; level < MG_BW_WL; ++level, lv += MG_BW_LEVEL_BYTES) {
		const uint32_t ol = mg_bw_rank<G>(lv, l), orr = mg_bw_rank<G>(lv, r);
		const uint32_t nz = (r - l) - (orr - ol);
		const bool one = k >= nz;
		l = one ? MG_BW_HALF + ol : l - ol;
		r = one ? MG_BW_HALF + orr : r - orr;
		k = one ? k - nz : k;
		bucket = (bucket << 1) | (one ? 1u : 0u);
	}
	return bucket << MG_BW_BUCKET_BITS;
}

// which of the 32 ranks rank0 .. rank0 + 31 sit at positions inside the window [l0, l0 + len): bit j = rank rank0 + j. Membership of two
// positions at once: with p < B <= 2^14 in either half of a word, p + (2^15 - l0) has bit 15 set iff p >= l0 and p + (2^15 - l0 - len)
// has it clear iff p < l0 + len, neither addition carries into the other half.
__device__ __forceinline__ uint32_t mg_bw_members(const uint16_t *__restrict__ pos, uint32_t rank0, uint32_t l0, uint32_t len)
{
	const uint4 *p4 = reinterpret_cast<const uint4 *>(pos + rank0);
	const uint32_t addlo = (0x8000u - l0) * 0x10001u, addhi = (0x8000u - l0 - len) * 0x10001u;
	uint32_t m = 0;
#pragma unroll
	for (int g = 0; g < 4; ++g) {
		const uint4 qv = __ldg(p4 + g);
		const uint32_t w[4] = {qv.x, qv.y, qv.z, qv.w};
#pragma unroll
		for (int e = 0; e < 4; ++e) {
			const int j = (g & 1) * 4 + e; // word j of its group of 16 ranks: ranks j (low half) and j + 8 (high half)
			const uint32_t in = (w[e] + addlo) & ~(w[e] + addhi) & 0x80008000u;
			// first group: rank j -> bit j, rank j + 8 -> bit 16 + j; second group: bits 8 + j and 24 + j
			m |= g < 2 ? in >> (15 - j) : in >> (7 - j);
		}
	}
	return __byte_perm(m, 0, 0x3120); // bytes (ranks 0-7, 16-23, 8-15, 24-31) -> rank order
}

// the (k + 1)-th set bit among the low 16 of m (k < 16; it exists): a bisection by population counts -- the same four steps in every
// lane, where clearing k bits one by one runs as long as the warp's largest k
__device__ __forceinline__ uint32_t mg_bw_pick16(uint32_t m, uint32_t k)
{
	uint32_t c = __popc(m & 0xffu);
	bool up = k >= c;
	k = up ? k - c : k;
	uint32_t j = up ? 8u : 0u;
	c = __popc((m >> j) & 0xfu);
	up = k >= c;
	k = up ? k - c : k;
	j = up ? j + 4u : j;
	c = __popc((m >> j) & 0x3u);
	up = k >= c;
	k = up ? k - c : k;
	j = up ? j + 2u : j;
	c = (m >> j) & 1u;
	return k >= c ? j + 1u : j;
}

// rank of the k-th smallest of the window by a walk of its own: what the successor search falls back to when the next member is not
// among the 31 ranks behind the k-th -- windows with a handful of values in a block full of other values (a long run of NaN under the
// window) have their members hundreds of ranks apart, and looking at pos_of_rank entry by entry is one dependent load per rank
template <bool G>
__device__ __noinline__ uint32_t mg_bw_rank_of_kth(const uint8_t *lv, const uint16_t *pos, uint32_t l, uint32_t r, uint32_t k)
{
	const uint32_t rank0 = mg_bw_walk<G>(lv, l, r, k);
	return rank0 + mg_bw_pick16(mg_bw_members(pos, rank0, l, r - l), k);
}

// k-th smallest (0-based) non-NaN value of the block-local window [l, r), k < n <= r - l <= S; with want2 (k + 1 < n) also the
// (k+1)-th. lv = the record's walked levels (global or staged), pos = the record's pos_of_rank in global memory, vals_block =
// the track (or its staged copy) at the block's first bin.
// G: where the levels live (true: global memory), GV: where the bins live
template <bool G, bool GV = G>
__device__ __forceinline__ void mg_bw_select(const uint8_t *__restrict__ lv, const uint16_t *__restrict__ pos, const float *__restrict__ vals_block, uint32_t l,
											 uint32_t r, uint32_t k, bool want2, float &x1, float &x2)
{
	// the bucket's 16 ranks hold >= kb + 1 window members; the k-th smallest is the (kb + 1)-th of them in rank order, its successor the
	// next member, here or in the ranks that follow
	uint32_t kb = k;
	const uint32_t rank0 = mg_bw_walk<G>(lv, l, r, kb);
	uint32_t m = mg_bw_members(pos, rank0, l, r - l);
	const uint32_t j1 = mg_bw_pick16(m, kb);
	x1 = mg_bw_ld<GV>(vals_block + __ldg(pos + mg_bw_pos_slot(rank0 + j1)));
	x2 = x1;
	if (want2) {
		m &= 0xfffffffeu << j1; // the members above
		uint32_t rk = rank0 + (uint32_t)__ffs(m) - 1u;
		if (!m) // none among the 31 ranks that follow (it exists: k + 1 < n)
			rk = mg_bw_rank_of_kth<G>(lv, pos, l, r, k + 1u);
		x2 = mg_bw_ld<GV>(vals_block + __ldg(pos + mg_bw_pos_slot(rk)));
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
