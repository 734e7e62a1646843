// dense_index.cuh -- per-chromosome index structures built once at staging so that EVERY dense window function of an
// arbitrary row is O(1) or O(log) lookups from one thread (no per-row sweep of the window):
//
//   prefix sums (dense.cuh)      avg / sum / nearest / size / exists      2 x 16-byte entries
//   min / max pyramids           fan-out 8; level j entry i = extremum of level j-1 entries [8i, 8i+8)
//                                a window decomposes into <= 7 head + <= 7 tail entries per level: 2 sectors per level
//   wavelet matrix               over the DENSE RANKS of the non-NaN values (rank = position in the chromosome's sorted
//                                order, ties by position; compacted in bin order, a window maps to
//                                [cnt[sbin], cnt[ebin]) through the prefix counts). Ranks make every level split any
//                                range in half whatever the value distribution, so a window of n values is down to
//                                <= 8 candidates after log2(n/8) levels and finishes from a checkpoint. Level L holds
//                                bit (levels-1-L) of every rank after the stable partitions of levels 0..L-1, in 8-byte
//                                units {uint32 ones before the unit, uint32 bits}: a rank query = one 8-byte load + popc.
//                                Exact for any window size; value of rank r = wm_sorted[r].
#pragma once
#include "common.cuh"
#include "dense.cuh"
#include "dense_blockwm.cuh"

// ------------------------------------------------------------------------------------------------------------------
// pyramids
__global__ void __launch_bounds__(256) k_pyr_build(const float *__restrict__ child_max, const float *__restrict__ child_min, int64_t nchild,
													float *pmax, float *pmin, int64_t nparent_padded)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nparent_padded)
		return;
	float mx = __int_as_float(0xff800000), mn = __int_as_float(0x7f800000);
#pragma unroll
	for (int j = 0; j < 8; ++j) {
		const int64_t c = i * 8 + j;
		if (c < nchild) {
			mx = fmaxf(mx, child_max[c]); // fmaxf/fminf return the non-NaN operand: NaN bins are skipped
			mn = fminf(mn, child_min[c]);
		}
	}
	pmax[i] = mx;
	pmin[i] = mn;
}

// sparse-table level j from level j - 1: out[g] = ext(in[g], in[g + half]) (the second operand only where it exists)
__global__ void __launch_bounds__(256) k_st_build(const float *__restrict__ in_max, const float *__restrict__ in_min, int64_t n, int64_t half,
												   float *out_max, float *out_min)
{
	const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= n)
		return;
	float mx = in_max[g], mn = in_min[g];
	if (g + half < n) {
		mx = fmaxf(mx, in_max[g + half]);
		mn = fminf(mn, in_min[g + half]);
	}
	out_max[g] = mx;
	out_min[g] = mn;
}

// extremum of the full groups [gs, ge) folded into acc: two sparse-table lookups, or the pyramid for very long runs
template <bool IS_MAX>
__device__ __forceinline__ float mg_pyr_query_from(const DenseChrom *ch, int level, int64_t lo, int64_t hi, float acc);

template <bool IS_MAX>
__device__ __forceinline__ float mg_groups_ext(const DenseChrom *ch, int64_t gs, int64_t ge, float acc)
{
	const int64_t len = ge - gs;
	if (len <= 0)
		return acc;
	if (len >= (1ll << MG_ST_LEVELS))
		return mg_pyr_query_from<IS_MAX>(ch, 1, gs, ge, acc);
	const int j = 31 - __clz((int)len);
	const float *t = IS_MAX ? ch->smax[j] : ch->smin[j];
	const float a = __ldg(t + gs), b = __ldg(t + ge - (1ll << j));
	return IS_MAX ? fmaxf(acc, fmaxf(a, b)) : fminf(acc, fminf(a, b));
}

// extremum over entries [lo, hi) of one aligned group of 8 (lo, hi within [g, g+8]); two 16-byte loads = one sector
template <bool IS_MAX>
__device__ __forceinline__ float mg_group_ext(const float *__restrict__ p, int64_t g, int lo, int hi, float acc)
{
	const float4 a = __ldg(reinterpret_cast<const float4 *>(p + g));
	const float4 b = __ldg(reinterpret_cast<const float4 *>(p + g + 4));
	const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
	for (int j = 0; j < 8; ++j)
		if (j >= lo && j < hi)
			acc = IS_MAX ? fmaxf(acc, v[j]) : fminf(acc, v[j]);
	return acc;
}

// extremum of entries [lo, hi) of pyramid level `level` (0 = the bins themselves) folded into acc; -inf (max) / +inf (min)
// when every bin is NaN
template <bool IS_MAX>
__device__ __forceinline__ float mg_pyr_query_from(const DenseChrom *ch, int level, int64_t lo, int64_t hi, float acc)
{
	for (; lo < hi; ++level) {
		const float *p = level == 0 ? ch->vals : (IS_MAX ? ch->pmax[level - 1] : ch->pmin[level - 1]);
		const int64_t glo = lo & ~7ll, ghi = hi & ~7ll;
		if (glo == ghi || (glo + 8 == hi)) { // the whole range sits in one group
			acc = mg_group_ext<IS_MAX>(p, glo, (int)(lo - glo), (int)(hi - glo), acc);
			break;
		}
		int64_t nlo = lo;
		if (lo != glo) { // partial head group
			acc = mg_group_ext<IS_MAX>(p, glo, (int)(lo - glo), 8, acc);
			nlo = glo + 8;
		}
		if (hi != ghi) // partial tail group
			acc = mg_group_ext<IS_MAX>(p, ghi, 0, (int)(hi - ghi), acc);
		lo = nlo >> 3;
		hi = ghi >> 3;
	}
	return acc;
}

template <bool IS_MAX>
__device__ __forceinline__ float mg_pyr_query(const DenseChrom *ch, int64_t s, int64_t e)
{
	return mg_pyr_query_from<IS_MAX>(ch, 0, s, e, IS_MAX ? __int_as_float(0xff800000) : __int_as_float(0x7f800000));
}

// ------------------------------------------------------------------------------------------------------------------
// wavelet matrix
#define MG_WM_LEVELS 32 // upper bound; a chromosome uses ceil(log2(#values)) levels

// compact the non-NaN values' keys in bin order: keys[j] = fkey(vals[i]), pos[j] = j with j = cnt[i]
__global__ void __launch_bounds__(256) k_wm_compact(const float *__restrict__ vals, const Pref *__restrict__ pref, int64_t nbins, uint32_t *keys,
													 uint32_t *pos)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nbins)
		return;
	const float v = vals[i];
	if (!isnan(v)) {
		const long long j = pref[i].cnt;
		keys[j] = mg_fkey(v);
		pos[j] = (uint32_t)j;
	}
}

// A level is an array of 8-byte units {uint32 ones before the unit, uint32 bits of 32 positions}: a rank query is one
// 8-byte load, a shift and a popcount. One thread per unit builds the bit word.
__global__ void __launch_bounds__(256) k_wm_bits(const uint32_t *__restrict__ keys, int64_t m, int shift, uint2 *units, int64_t nunits)
{
	const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (u >= nunits)
		return;
	const int64_t p0 = u * 32;
	uint32_t bits = 0;
#pragma unroll 8
	for (int j = 0; j < 32; ++j) {
		const int64_t p = p0 + j;
		if (p < m)
			bits |= ((keys[p] >> shift) & 1u) << j;
	}
	units[u] = make_uint2(0u, bits);
}

// one block: exclusive scan of the units' popcounts into their headers; zeros[level] = m - total ones
__global__ void __launch_bounds__(1024) k_wm_scan(uint2 *units, int64_t nunits, int64_t m, uint32_t *zeros_out)
{
	__shared__ long long sc[33];
	const int64_t chunk = (nunits + blockDim.x - 1) / blockDim.x;
	const int64_t lo = (int64_t)threadIdx.x * chunk, hi = min(lo + chunk, nunits);
	long long s = 0;
	for (int64_t u = lo; u < hi; ++u)
		s += __popc(units[u].y);
	long long total;
	long long s0 = mg_block_excl_scan<long long>(s, sc, total);
	for (int64_t u = lo; u < hi; ++u) {
		const uint32_t bits = units[u].y;
		units[u].x = (uint32_t)s0;
		s0 += __popc(bits);
	}
	if (threadIdx.x == 0)
		*zeros_out = (uint32_t)(m - total);
}

// ones among positions [0, i) of a level; also the bit AT position i
__device__ __forceinline__ uint32_t mg_wm_rank1(const uint2 *__restrict__ level, uint32_t i, uint32_t &bit_at_i)
{
	const uint2 u = __ldg(level + (i >> 5));
	const uint32_t bi = i & 31u;
	bit_at_i = (u.y >> bi) & 1u;
	return u.x + __popc(u.y & ((1u << bi) - 1u));
}

// stable partition of the keys by the level's bit
__global__ void __launch_bounds__(256) k_wm_scatter(const uint32_t *__restrict__ keys, int64_t m, const uint2 *__restrict__ level, uint32_t zeros,
													 uint32_t *keys_out)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m)
		return;
	uint32_t bit;
	const uint32_t r1 = mg_wm_rank1(level, (uint32_t)i, bit);
	keys_out[bit ? zeros + r1 : (uint32_t)i - r1] = keys[i];
}

// one pass of the LSD bit sort that ranks the values: stable partition of (key, pos) pairs by the pass's bit
__global__ void __launch_bounds__(256) k_wm_sort_scatter(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ pos, int64_t m,
														  const uint2 *__restrict__ level, uint32_t zeros, uint32_t *keys_out, uint32_t *pos_out)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m)
		return;
	uint32_t bit;
	const uint32_t r1 = mg_wm_rank1(level, (uint32_t)i, bit);
	const uint32_t dst = bit ? zeros + r1 : (uint32_t)i - r1;
	keys_out[dst] = keys[i];
	pos_out[dst] = pos[i];
}

// after the sort: sorted[j] = value of rank j; rank_of[pos[j]] = j
__global__ void __launch_bounds__(256) k_wm_assign_ranks(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ pos, int64_t m, float *sorted,
														  uint32_t *rank_of)
{
	const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= m)
		return;
	sorted[j] = mg_fkey_inv(keys[j]);
	rank_of[pos[j]] = (uint32_t)j;
}

// checkpoint content: the VALUE key of every element in the order entering a level (ordering by value key == ordering by
// rank), so a finished walk needs no rank -> value lookup
__global__ void __launch_bounds__(256) k_wm_ckpt(const uint32_t *__restrict__ rank_keys, int64_t m, const float *__restrict__ sorted, uint32_t *ck)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < m)
		ck[i] = mg_fkey(sorted[rank_keys[i]]);
}

struct WmState {
	uint32_t l, r, k;
};

// advance one order statistic by one level; returns the bit it takes
__device__ __forceinline__ bool mg_wm_step(const uint2 *__restrict__ lv, uint32_t zeros, WmState &st)
{
	uint32_t bit_l, bit_r;
	const uint32_t ol = mg_wm_rank1(lv, st.l, bit_l);
	if (st.r - st.l == 1) { // singleton range: its own bit decides, one rank query is enough
		st.l = bit_l ? zeros + ol : st.l - ol;
		st.r = st.l + 1;
		st.k = 0;
		return bit_l != 0;
	}
	const uint32_t orr = mg_wm_rank1(lv, st.r, bit_r);
	const uint32_t nz = (st.r - st.l) - (orr - ol);
	if (st.k < nz) {
		st.l -= ol;
		st.r -= orr;
		return false;
	}
	st.k -= nz;
	st.l = zeros + ol;
	st.r = zeros + orr;
	return true;
}

// ---- finishing a walk from a checkpoint ------------------------------------------------------------------------------
// wm_ckpt[c] holds the full keys in the order entering level 4 (c + 1). Once a range holds <= MG_WM_FINISH keys, the
// order statistics are read off directly: three 16-byte loads cover any 8 consecutive keys, an 8-input sorting network
// (19 compare-exchanges) orders them.
#define MG_WM_FINISH 8

__device__ __forceinline__ void mg_cswap(uint32_t &a, uint32_t &b)
{
	const uint32_t lo = min(a, b), hi = max(a, b);
	a = lo;
	b = hi;
}

__device__ __forceinline__ uint32_t mg_pick8(const uint32_t (&v)[8], uint32_t k)
{
	uint32_t r = v[0];
#pragma unroll
	for (int j = 1; j < 8; ++j)
		r = k == (uint32_t)j ? v[j] : r;
	return r;
}

// k-th (and optionally (k+1)-th) smallest of ckpt[l, r), r - l <= 8
__device__ __forceinline__ void mg_wm_finish(const uint32_t *__restrict__ ckpt, uint32_t l, uint32_t r, uint32_t k, bool want2, uint32_t &key1,
											 uint32_t &key2)
{
	const uint32_t base = l & ~3u;
	const uint4 q0 = __ldg(reinterpret_cast<const uint4 *>(ckpt + base));
	const uint4 q1 = __ldg(reinterpret_cast<const uint4 *>(ckpt + base + 4));
	const uint4 q2 = __ldg(reinterpret_cast<const uint4 *>(ckpt + base + 8));
	const uint32_t c[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
	const uint32_t off = l - base, n = r - l;
	uint32_t v[8];
#pragma unroll
	for (int j = 0; j < 8; ++j) { // v[j] = c[off + j] for j < n, else +inf padding
		uint32_t x = c[j];
		x = off == 1 ? c[j + 1] : x;
		x = off == 2 ? c[j + 2] : x;
		x = off == 3 ? c[j + 3] : x;
		v[j] = (uint32_t)j < n ? x : 0xffffffffu;
	}
	mg_cswap(v[0], v[1]); mg_cswap(v[2], v[3]); mg_cswap(v[4], v[5]); mg_cswap(v[6], v[7]);
	mg_cswap(v[0], v[2]); mg_cswap(v[1], v[3]); mg_cswap(v[4], v[6]); mg_cswap(v[5], v[7]);
	mg_cswap(v[1], v[2]); mg_cswap(v[5], v[6]);
	mg_cswap(v[0], v[4]); mg_cswap(v[1], v[5]); mg_cswap(v[2], v[6]); mg_cswap(v[3], v[7]);
	mg_cswap(v[2], v[4]); mg_cswap(v[3], v[5]);
	mg_cswap(v[1], v[2]); mg_cswap(v[3], v[4]); mg_cswap(v[5], v[6]);
	key1 = mg_pick8(v, k);
	key2 = want2 ? mg_pick8(v, k + 1) : key1;
}

// RANK of the k-th smallest value (0-based) among compacted positions [l, r), r > l, k < r - l; with want2 (requires
// k + 1 < r - l) also of the (k+1)-th. The two order statistics share one walk until the level where they part; every walk stops at the
// first checkpoint level (4, 8, ..., 28) where its range holds <= 8 keys.
__device__ __noinline__ void mg_wm_select(const DenseChrom *ch, uint32_t l, uint32_t r, uint32_t k, bool want2, float &x1, float &x2)
{
	uint32_t key1, key2;
	const size_t stride = (size_t)ch->wm_blocks;
	const uint2 *lv = ch->wm;
	const uint32_t *zp = ch->wm_zeros;
	WmState a{l, r, k}, b{0, 0, 0};
	bool joint = want2, split = false, a_done = false, b_done = false;
	uint32_t v1 = 0, v2 = 0;
	const int nlevels = ch->wm_levels;
	uint32_t bitmask = nlevels ? 1u << (nlevels - 1) : 0u;
	// rank keys: no level is constant (rank 0 has every bit clear, rank m-1 has the top bit set), so every level is walked
	for (int level = 0; level < nlevels; ++level, lv += stride, bitmask >>= 1) {
		if (level >= 4 && (level & 3) == 0) { // a checkpoint: finish whatever has become small
			const uint32_t *ck = ch->wm_ckpt[(level >> 2) - 1];
			if (joint) {
				if (a.r - a.l <= MG_WM_FINISH) {
					mg_wm_finish(ck, a.l, a.r, a.k, true, key1, key2);
					x1 = mg_fkey_inv(key1);
					x2 = mg_fkey_inv(key2);
					return;
				}
			} else {
				uint32_t dummy;
				if (!a_done && a.r - a.l <= MG_WM_FINISH) {
					mg_wm_finish(ck, a.l, a.r, a.k, false, key1, dummy);
					x1 = mg_fkey_inv(key1);
					a_done = true;
				}
				if (split && !b_done && b.r - b.l <= MG_WM_FINISH) {
					mg_wm_finish(ck, b.l, b.r, b.k, false, key2, dummy);
					x2 = mg_fkey_inv(key2);
					b_done = true;
				}
				if (a_done && (b_done || !split))
					break;
			}
		}
		const uint32_t zeros = __ldg(zp + level);
		if (joint) { // both statistics still in one range (which therefore holds >= 2 elements)
			uint32_t bit_l, bit_r;
			const uint32_t ol = mg_wm_rank1(lv, a.l, bit_l), orr = mg_wm_rank1(lv, a.r, bit_r);
			const uint32_t nz = (a.r - a.l) - (orr - ol);
			if (a.k + 1 < nz) {
				a.l -= ol;
				a.r -= orr;
			} else if (a.k >= nz) {
				a.k -= nz;
				a.l = zeros + ol;
				a.r = zeros + orr;
				v1 |= bitmask;
				v2 |= bitmask;
			} else { // k is the last zero, k + 1 the first one: they part here
				b.l = zeros + ol;
				b.r = zeros + orr;
				b.k = 0;
				v2 |= bitmask;
				a.l -= ol;
				a.r -= orr;
				joint = false;
				split = true;
			}
		} else {
			if (!a_done && mg_wm_step(lv, zeros, a))
				v1 |= bitmask;
			if (split && !b_done && mg_wm_step(lv, zeros, b))
				v2 |= bitmask;
		}
	}
	// whatever walked all the way down holds a rank: look its value up
	if (!a_done)
		x1 = __ldg(ch->wm_sorted + v1);
	if (!want2)
		x2 = x1;
	else if (!b_done)
		x2 = __ldg(ch->wm_sorted + v2);
}

// ------------------------------------------------------------------------------------------------------------------
// lse pyramid: fan-out 8 like the min / max pyramids, entries = log-sum-exp of their 8 children in double. A window decomposes
// into <= 7 head + <= 7 tail entries per level (at most ~45 terms for the 517 bins of a C3 window instead of 517), and
// lse(window) = M + log(sum exp(term - M)) over those terms with M their maximum -- every term is exact to double rounding,
// so the result equals the direct sum's to ~1e-15 before the cast to float. One thread per row.
template <typename T>
__global__ void __launch_bounds__(256) k_lse_build(const T *__restrict__ child, int64_t nchild, double *__restrict__ parent, int64_t nparent)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nparent)
		return;
	double v[8], m = -INFINITY;
#pragma unroll
	for (int j = 0; j < 8; ++j) {
		const int64_t c = i * 8 + j;
		const double x = c < nchild ? (double)child[c] : -INFINITY;
		v[j] = x == x ? x : -INFINITY; // NaN bins hold no value
		m = fmax(m, v[j]);
	}
	double s = 0;
	if (m > -INFINITY) {
#pragma unroll
		for (int j = 0; j < 8; ++j)
			s += exp(v[j] - m); // exp(-inf) = 0
	}
	parent[i] = m > -INFINITY ? m + log(s) : -INFINITY;
}

// calls f(term) for every term of the decomposition of bins [lo, hi) (lo < hi), NaN bins included (as NaN)
template <typename F>
__device__ __forceinline__ void mg_lse_terms(const DenseChrom *ch, int64_t lo, int64_t hi, F f)
{
	for (int level = 0; lo < hi; ++level) {
		const float *pv = ch->vals;
		const double *pd = level ? ch->plse[level - 1] : nullptr;
		const int64_t glo = (lo + 7) & ~7ll, ghi = hi & ~7ll;
		if (glo >= ghi || level == MG_PYR_MAX || !ch->plse[level]) { // no full group in between (or no level above): take the entries
			for (int64_t i = lo; i < hi; ++i)
				f(level ? __ldg(pd + i) : (double)__ldg(pv + i));
			return;
		}
		for (int64_t i = lo; i < glo; ++i)
			f(level ? __ldg(pd + i) : (double)__ldg(pv + i));
		for (int64_t i = ghi; i < hi; ++i)
			f(level ? __ldg(pd + i) : (double)__ldg(pv + i));
		lo = glo >> 3;
		hi = ghi >> 3;
	}
}

__global__ void __launch_bounds__(256) k_dense_lse(const DenseQuery q)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= q.count)
		return;
	DenseWin w;
	double lse = mg_nan();
	if (mg_dense_win(q, q.first + i, w) && w.ebin > w.sbin) {
		if (w.one_bin) // assign_single_bin_value: the bin's value
			lse = (double)__ldg(w.ch->vals + w.sbin);
		else {
			double M = -INFINITY;
			mg_lse_terms(w.ch, w.sbin, w.ebin, [&](double x) { M = fmax(M, x); }); // fmax skips NaN
			if (M > -INFINITY) {
				double s = 0;
				mg_lse_terms(w.ch, w.sbin, w.ebin, [&](double x) {
					if (x == x)
						s += exp(x - M);
				});
				lse = mg_f2d(M + log(s));
			}
		}
	}
	mg_st_stream(q.out[MG_LSE] + i, lse);
}

// lse as the reference's GENERIC path accumulates it (MG_FUNC_GENERIC): pairwise in float, bin order, from -inf
// (lse_accumulate, src/GenomeTrack1D.h:18-28; the loop of src/GenomeTrackFixedBin.cpp:881-990). One thread per row walks its window.
__device__ __forceinline__ void mg_lse_accumulate_f(float &l1, float l2)
{
	if (l1 > l2) {
		if (!isinf(l2))
			l1 += logf(1.0f + expf(l2 - l1));
	} else {
		if (isinf(l1))
			l1 = l2;
		else
			l1 = l2 + logf(1.0f + expf(l1 - l2));
	}
}
__global__ void __launch_bounds__(256) k_dense_lse_generic(const DenseQuery q)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= q.count)
		return;
	DenseWin w;
	double lse = mg_nan();
	if (mg_dense_win(q, q.first + i, w) && w.ebin > w.sbin) {
		if (w.one_bin)
			lse = (double)__ldg(w.ch->vals + w.sbin);
		else {
			float l = -INFINITY;
			int64_t n = 0;
			for (int64_t b = w.sbin; b < w.ebin; ++b) {
				const float v = __ldg(w.ch->vals + b);
				if (v == v) {
					mg_lse_accumulate_f(l, v);
					++n;
				}
			}
			if (n)
				lse = (double)l;
		}
	}
	mg_st_stream(q.out[MG_LSE] + i, lse);
}

// size / exists as the reference's GENERIC path reports them for a window inside ONE bin: 1 whatever the bin holds
// (assign_single_bin_value, src/GenomeTrackFixedBin.cpp:150-180; the bin must exist in the file, :664-712). Runs after the kernel
// that wrote the reducer-path columns and overwrites those rows only.
__global__ void __launch_bounds__(256) k_dense_generic_counts(const DenseQuery q)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= q.count)
		return;
	DenseWin w;
	if (!mg_dense_win(q, q.first + i, w) || !w.one_bin || w.ebin <= w.sbin)
		return;
	if ((q.generic_mask >> MG_SIZE) & 1u)
		q.out[MG_SIZE][i] = 1.;
	if ((q.generic_mask >> MG_EXISTS) & 1u)
		q.out[MG_EXISTS][i] = 1.;
}

// ------------------------------------------------------------------------------------------------------------------
// k_dense_argext: min.pos / max.pos, one thread per row, from the same structures as min / max: the extremum m of the window
// first (two partial groups + two sparse-table entries), then the FIRST bin holding it -- the reference's strict comparisons
// keep the first occurrence (src/GenomeTrackFixedBin.cpp:886-905) -- in window order: the head group's bins, then the full
// groups by a descent through the sparse tables (is m in the left 2^j groups? one lookup per halving, <= 12), then the tail
// group's bins. A warp-per-row sweep of the C3 windows takes 8.9 ms; windows of 2^12 groups and more (not covered by the sparse
// tables) are scanned by their thread.
template <bool IS_MAX>
__device__ __forceinline__ int64_t mg_first_bin_eq(const float *__restrict__ vals, int64_t lo, int64_t hi, float m)
{
	for (int64_t b = lo; b < hi; ++b)
		if (__ldg(vals + b) == m)
			return b;
	return -1;
}

template <bool IS_MAX>
__device__ __forceinline__ int64_t mg_arg_ext(const DenseChrom *ch, int64_t sbin, int64_t ebin)
{
	const float *vals = ch->vals;
	const float seed = IS_MAX ? __int_as_float(0xff800000) : __int_as_float(0x7f800000);
	const int64_t g0 = sbin >> 3, g1 = (ebin - 1) >> 3;
	if (g0 == g1 || g1 - g0 - 1 >= (1ll << MG_ST_LEVELS)) { // inside one group, or too long for the tables: scan
		float m = seed;
		int64_t at = -1;
		for (int64_t b = sbin; b < ebin; ++b) {
			const float v = __ldg(vals + b);
			if (v == v && (IS_MAX ? v > m : v < m)) {
				m = v;
				at = b;
			}
		}
		return at;
	}
	const int64_t gs = g0 + 1, ge = g1; // full groups
	const float mh = mg_group_ext<IS_MAX>(vals, 8 * g0, (int)(sbin & 7), 8, seed);
	const float mt = mg_group_ext<IS_MAX>(vals, 8 * g1, 0, (int)(ebin - 8 * g1), seed);
	const float mf = mg_groups_ext<IS_MAX>(ch, gs, ge, seed);
	const float m = IS_MAX ? fmaxf(mh, fmaxf(mf, mt)) : fminf(mh, fminf(mf, mt));
	if (m == seed)
		return -1; // no value in the window
	if (mh == m)
		return mg_first_bin_eq<IS_MAX>(vals, sbin, 8 * gs, m);
	if (mf == m) {
		int64_t lo = gs, hi = ge; // some group of [lo, hi) holds m, none before lo does
		while (hi - lo > 1) {
			const int j = 31 - __clz((int)(hi - lo - 1)); // 2^j < hi - lo <= 2^(j+1)
			const float left = __ldg((IS_MAX ? ch->smax[j] : ch->smin[j]) + lo);
			if (left == m)
				hi = lo + (1ll << j);
			else
				lo += 1ll << j;
		}
		return mg_first_bin_eq<IS_MAX>(vals, 8 * lo, 8 * lo + 8, m);
	}
	return mg_first_bin_eq<IS_MAX>(vals, 8 * g1, ebin, m);
}

__global__ void __launch_bounds__(256) k_dense_argext(const DenseQuery q)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= q.count)
		return;
	DenseWin w;
	const bool ok = mg_dense_win(q, q.first + i, w);
	int64_t mnb = -1, mxb = -1;
	if (ok && w.ebin > w.sbin) {
		if (w.one_bin) // assign_single_bin_value: the bin's position whatever its value (see k_dense_ext)
			mnb = mxb = w.sbin;
		else {
			if (q.out[MG_MAX_POS])
				mxb = mg_arg_ext<true>(w.ch, w.sbin, w.ebin);
			if (q.out[MG_MIN_POS])
				mnb = mg_arg_ext<false>(w.ch, w.sbin, w.ebin);
		}
	}
	const double bs = (double)q.bin_size, ws = ok ? (double)w.ws : 0.;
	if (q.out[MG_MAX_POS]) {
		double p = mg_nan();
		if (mxb >= 0) {
			p = fmax((double)mxb * bs, ws);
			if (q.pos_rel & (1u << MG_MAX_POS))
				p -= ws;
		}
		mg_st_stream(q.out[MG_MAX_POS] + i, p);
	}
	if (q.out[MG_MIN_POS]) {
		double p = mg_nan();
		if (mnb >= 0) {
			p = fmax((double)mnb * bs, ws);
			if (q.pos_rel & (1u << MG_MIN_POS))
				p -= ws;
		}
		mg_st_stream(q.out[MG_MIN_POS] + i, p);
	}
}

// ------------------------------------------------------------------------------------------------------------------
// k_dense_rowquery: one thread per row, every function from the index structures read through L1 / L2 -- the path for rows in
// any order and at any distance from each other (k_dense_tile, dense_tile.cuh, serves sorted rows out of shared memory and falls
// back to mg_rq_row for the tiles it cannot hold).
//   SUM: avg / sum / nearest / size / exists   MM: min / max   Q: quantiles
//   SD: stddev = sqrt(M2 / (n - 1)) with M2 = S2 - S1^2 / n from the prefix sums of v and v^2. When that difference has
//       lost too many digits (low variance against the mean, or a window that is tiny against the chromosome's running
//       sum of squares) the window is re-read and Welford's update is run in bin order, i.e. the reference's arithmetic.
#define MG_SD_DIRECT_BINS 8
#ifndef MG_RQ_THREADS
#define MG_RQ_THREADS 256
#endif
// Resident blocks per SM, measured on B200 at the C3 shape: the kernel is bound by memory latency, so more warps win even at the
// price of a few spills -- 8 blocks (32 registers) for one function group, 6 for several.
#ifdef MG_RQ_MB // build-time override (tuning)
#define MG_RQ_MINBLOCKS(SUM, MM, Q, SD) MG_RQ_MB
#else
#define MG_RQ_MINBLOCKS(SUM, MM, Q, SD) (((SD) ? 1024 : ((int)(SUM) + (int)(MM) + (int)(Q) <= 1 ? 2048 : 1536)) / MG_RQ_THREADS)
#endif

// lv_staged: this row's walked levels in shared memory (null: read them through L1); bar: the mbarrier the staging copies complete
// on (null: nothing to wait for) -- waited for as late as possible, the sums and extrema are computed while the copies fly
template <bool SUM, bool MM, bool Q, bool SD>
__device__ __forceinline__ void mg_rq_row(const DenseQuery &q, int64_t i, const DenseWin &w, bool ok, const uint8_t *lv_staged = nullptr,
										  uint64_t *bar = nullptr)
{
	RowSum rs;
	rs.avg = rs.sum = rs.size = rs.exists = mg_nan();
	double mn = mg_nan(), mx = mg_nan(), sd = mg_nan();
	const int64_t nb = ok ? w.ebin - w.sbin : 0;
	WinBasic wb;
	wb.n = 0;
	wb.before = 0;
	if (ok) {
		rs.size = 0;
		rs.exists = 0;
	}
	// a quantile-only row inside one block needs neither bins nor prefix entries: its non-NaN count is in the block record
	const bool count_from_block = Q && !SUM && !SD && !MM && q.use_bw && nb > 0 && nb <= MG_BW_S;
	if (count_from_block) {
		const uint32_t blk = (uint32_t)w.sbin / MG_BW_STRIDE, l = (uint32_t)w.sbin - blk * MG_BW_STRIDE;
		wb.n = mg_bw_count(w.ch->bw + (size_t)blk * MG_BW_RECORD_BYTES, l, l + (uint32_t)nb);
	} else if (nb > 0) {
		// one pass over the window's two partial groups feeds the sum, the count and level 0 of the pyramids
		mg_win_basic<SUM || SD, MM, SUM || SD || Q>(w.ch, w.sbin, w.ebin, wb);
		if (SUM || SD)
			mg_row_sum(w.ch, wb, w.sbin, w.ebin, rs);
	}
	if (SUM) { // every column is stored as soon as it is known: short live ranges
		if (q.out[MG_AVG]) mg_st_stream(q.out[MG_AVG] + i, rs.avg);
		if (q.out[MG_NEAREST]) mg_st_stream(q.out[MG_NEAREST] + i, rs.avg);
		if (q.out[MG_SUM]) mg_st_stream(q.out[MG_SUM] + i, rs.sum);
		if (q.out[MG_SIZE]) mg_st_stream(q.out[MG_SIZE] + i, rs.size);
		if (q.out[MG_EXISTS]) mg_st_stream(q.out[MG_EXISTS] + i, rs.exists);
	}
	if (nb > 0) {
		if (SD && wb.n > 1) {
			const double S1 = wb.S, S2 = __ldg(w.ch->pref2 + w.ebin) - __ldg(w.ch->pref2 + w.sbin);
			double m2 = S2 - S1 * S1 / (double)wb.n;
			// trust the difference only while it keeps ~9 significant digits
			if (nb <= MG_SD_DIRECT_BINS || !(m2 > S2 * 1e-7) || !(m2 > w.ch->sq_sum * 1e-7)) {
				long long c = 0;
				double mean = 0;
				m2 = 0;
				for (int64_t t = w.sbin; t < w.ebin; ++t) { // src/GenomeTrackFixedBin.cpp:912-917
					const float v = __ldg(w.ch->vals + t);
					if (!isnan(v)) {
						++c;
						const double d = (double)v - mean;
						mean += d / (double)c;
						m2 += d * ((double)v - mean);
					}
				}
			}
			sd = mg_f2d(sqrt(m2 / (double)(wb.n - 1)));
		}
		if (MM) { // an extremum still at its -inf / +inf seed means the window holds no value
			if (q.out[MG_MAX]) {
				const float v = mg_groups_ext<true>(w.ch, wb.gs, wb.ge, wb.mx);
				if (v != __int_as_float(0xff800000))
					mx = (double)v;
			}
			if (q.out[MG_MIN]) {
				const float v = mg_groups_ext<false>(w.ch, wb.gs, wb.ge, wb.mn);
				if (v != __int_as_float(0x7f800000))
					mn = (double)v;
			}
		}
	}
	if (SD)
		mg_st_stream(q.out[MG_STDDEV] + i, sd);
	if (MM) {
		if (q.out[MG_MIN]) mg_st_stream(q.out[MG_MIN] + i, mn);
		if (q.out[MG_MAX]) mg_st_stream(q.out[MG_MAX] + i, mx);
	}
	if (Q) {
		// every quantile column is stored as soon as it is known (a per-thread array indexed by k would live in local memory)
		const long long n = wb.n;
		const bool use_bw = q.use_bw && nb <= MG_BW_S && nb > 0;
		const uint32_t blk = nb > 0 ? (uint32_t)w.sbin / MG_BW_STRIDE : 0u;
		const uint32_t l = nb > 0 ? (uint32_t)w.sbin - blk * MG_BW_STRIDE : 0u;
		long long before = wb.before;
		if (!use_bw && n > 0 && before < 0)
			before = mg_win_before(w.ch, w.sbin);
		if (bar) // every thread of a staging block waits for the copies (none may be in flight when the block retires)
			mg_mbar_wait(bar, 0);
		for (int k = 0; k < q.nq; ++k) {
			double qv = mg_nan();
			if (n > 0) {
				const double index = __dmul_rn((double)(n - 1), q.q_pct[k]); // q_pct: clamped to [0, 1] by the host
				const double fl = floor(index);
				const long long i1 = (long long)fl;
				const bool want2 = index != fl; // ceil(index) != floor(index)
				float x1, x2;
				if (use_bw) {
					const uint8_t *rec = w.ch->bw + (size_t)blk * MG_BW_RECORD_BYTES;
					const uint16_t *pos = reinterpret_cast<const uint16_t *>(rec + MG_BW_POS_OFFSET);
					const float *vb = w.ch->vals + (size_t)blk * MG_BW_STRIDE;
					if (lv_staged)
						mg_bw_select<false, true>(lv_staged, pos, vb, l, l + (uint32_t)nb, (uint32_t)i1, want2, x1, x2);
					else
						mg_bw_select<true, true>(rec, pos, vb, l, l + (uint32_t)nb, (uint32_t)i1, want2, x1, x2);
				} else
					mg_wm_select(w.ch, (uint32_t)before, (uint32_t)(before + n), (uint32_t)i1, want2, x1, x2);
				qv = mg_interp(x1, x2, __dsub_rn(index, fl));
			}
			mg_st_stream(q.q_out[k] + i, qv);
		}
	}
}

// Staged block records (Q): the 256 rows of a thread block are neighbours in position when the rows are sorted (what the scanner
// produces), so between them they need the records of a handful of consecutive blocks. The thread block finds the range of
// block ids its rows need (a min / max reduction, no assumption about the row order), thread 0 issues one bulk copy per
// record -- at most MG_RQ_SLOTS, the walked levels only -- into shared memory, the rows compute their avg / min / max from
// global memory while the copies fly, and the walks then run on shared memory. A row whose record is not among the staged ones
// (another chromosome, rows far apart, unsorted rows) walks the record through L1: correctness never depends on the order.
#define MG_RQ_SLOTS 5
#define MG_STAGE_BLK_BITS 21
template <bool SUM, bool MM, bool Q, bool SD>
__global__ void __launch_bounds__(MG_RQ_THREADS, MG_RQ_MINBLOCKS(SUM, MM, Q, SD)) k_dense_rowquery(const __grid_constant__ DenseQuery q)
{
	__shared__ __align__(128) uint8_t s_stage[Q ? MG_RQ_SLOTS * MG_BW_STAGE_BYTES : 16];
	__shared__ uint64_t s_bar;
	__shared__ uint32_t s_kmin, s_kmax;
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const bool live = i < q.count;
	const bool staging = Q && q.use_bw && q.stage;
	if (!staging && !live)
		return;
	if (staging && threadIdx.x == 0) {
		s_kmin = 0xffffffffu;
		s_kmax = 0;
		mg_mbar_init(&s_bar, 1);
	}
	DenseWin w;
	const bool ok = live && mg_dense_win(q, q.first + i, w);
	const uint8_t *lv_staged = nullptr;
	bool staged = false;
	if (staging) {
		const int64_t nb = ok ? w.ebin - w.sbin : 0;
		// one key per row: chromosome in the high bits, block id in the low MG_STAGE_BLK_BITS (rows beyond either range walk through L1)
		const uint32_t blk = nb > 0 ? (uint32_t)w.sbin / MG_BW_STRIDE : 0;
		const uint32_t chrom = nb > 0 ? (uint32_t)(w.ch - q.table) : 0;
		const bool mine = nb > 0 && nb <= MG_BW_S && blk < (1u << MG_STAGE_BLK_BITS) && chrom < (1u << (32 - MG_STAGE_BLK_BITS)) - 1u;
		const uint32_t key = (chrom << MG_STAGE_BLK_BITS) | blk;
		__syncthreads(); // the seeds and the barrier are initialised
		const uint32_t wmin = __reduce_min_sync(0xffffffffu, mine ? key : 0xffffffffu), wmax = __reduce_max_sync(0xffffffffu, mine ? key : 0u);
		if ((threadIdx.x & 31) == 0 && wmin != 0xffffffffu) {
			atomicMin(&s_kmin, wmin);
			atomicMax(&s_kmax, wmax);
		}
		__syncthreads();
		const uint32_t kmin = s_kmin, kmax = s_kmax;
		// one chromosome, and dense enough that the staged records serve most of the rows
		staged = kmin <= kmax && (kmin >> MG_STAGE_BLK_BITS) == (kmax >> MG_STAGE_BLK_BITS) && kmax - kmin < 2u * MG_RQ_SLOTS;
		if (staged) {
			const uint32_t nslots = min(kmax - kmin + 1u, (uint32_t)MG_RQ_SLOTS);
			if (threadIdx.x == 0) {
				const uint8_t *src = q.table[kmin >> MG_STAGE_BLK_BITS].bw + (size_t)(kmin & ((1u << MG_STAGE_BLK_BITS) - 1u)) * MG_BW_RECORD_BYTES;
				mg_mbar_expect_tx(&s_bar, nslots * MG_BW_STAGE_BYTES);
				for (uint32_t sl = 0; sl < nslots; ++sl)
					mg_bulk_g2s(s_stage + sl * MG_BW_STAGE_BYTES, src + (size_t)sl * MG_BW_RECORD_BYTES, MG_BW_STAGE_BYTES, &s_bar);
			}
			if (mine && key - kmin < nslots)
				lv_staged = s_stage + (key - kmin) * MG_BW_STAGE_BYTES;
		}
		if (!live) { // a partial last block: the copies must have landed before the block retires
			if (staged)
				mg_mbar_wait(&s_bar, 0);
			return;
		}
	}
	mg_rq_row<SUM, MM, Q, SD>(q, i, w, ok, lv_staged, staged ? &s_bar : nullptr);
}
