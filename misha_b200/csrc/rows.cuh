// rows.cuh -- iterator rows on the device.
//   explicit rows : SoA (chrom int32, start int64, end int64) -- the intervals iterator / uploaded batches
//   implicit rows : fixed-bin iterator over a sorted scope; row r of scope interval k covers
//                   [max(bin*b, s_k), min((bin+1)*b, e_k)), bin = floor(s_k / b) + (r - row0_k)
//                   (src/TrackExpressionFixedBinIterator.cpp:52-60). Never materialised.
#pragma once
#include "common.cuh"

#define MG_SCOPE_INLINE 32
struct RowSrc {
	int kind; // 0 explicit, 1 fixed-bin
	int64_t n;
	// explicit
	const int32_t *chrom;
	const int64_t *start;
	const int64_t *end;
	const int64_t *scope_idx; // may be null
	// fixed-bin
	int64_t binsize;
	int64_t nscope;
	const int32_t *schrom;
	const int64_t *sstart;
	const int64_t *send;
	const int64_t *srow0; // nscope + 1 exclusive prefix of rows per scope interval
	const int64_t *sbin0; // floor(sstart / binsize) per scope interval (host-computed)
	// a copy of srow0 that travels in the kernel parameters when the scope is short (a genome's chromosomes): the search for a
	// row's scope interval then reads the constant bank instead of paying a chain of dependent global loads per thread block
	int64_t srow0_inl[MG_SCOPE_INLINE + 1];
};

struct mg_rows {
	RowSrc src{};
	bool owns = false;
	void *d_block = nullptr; // one allocation backing every owned array
	// host copy of a fixed-bin scope (few entries): lets the host split a scan into contiguous per-chromosome segments
	std::vector<int32_t> h_schrom;
	std::vector<int64_t> h_sstart, h_send, h_srow0;
};

// scope interval of implicit row r: last k with srow0[k] <= r
__device__ __forceinline__ int64_t mg_scope_of_row(const RowSrc &rs, int64_t r)
{
	if (rs.nscope <= MG_SCOPE_INLINE) {
		int lo = 0, hi = (int)rs.nscope;
		while (hi - lo > 1) {
			const int mid = (lo + hi) >> 1;
			if (rs.srow0_inl[mid] <= r)
				lo = mid;
			else
				hi = mid;
		}
		return lo;
	}
	int64_t lo = 0, hi = rs.nscope; // invariant: srow0[lo] <= r < srow0[hi]
	while (hi - lo > 1) {
		int64_t mid = (lo + hi) >> 1;
		if (__ldg(rs.srow0 + mid) <= r)
			lo = mid;
		else
			hi = mid;
	}
	return lo;
}

// floor(a / b) for a >= 0, b > 0 with a 32-bit fast path (a 64-bit division is ~100 instructions on the GPU)
__device__ __forceinline__ int64_t mg_div_pos(int64_t a, int64_t b)
{
	if (((uint64_t)a | (uint64_t)b) <= 0xffffffffull)
		return (int64_t)((uint32_t)a / (uint32_t)b);
	return a / b;
}

// scope interval cache of a thread that walks several implicit rows: the search and the division are paid once per
// scope interval, not once per row
struct RowCache {
	int64_t k = -1, row0 = 0, row1 = 0, bin_base = 0, sstart = 0, send = 0;
	int chrom = 0;
};

__device__ __forceinline__ void mg_rowcache_load(const RowSrc &rs, int64_t r, RowCache &c)
{
	c.k = mg_scope_of_row(rs, r);
	c.row0 = __ldg(rs.srow0 + c.k);
	c.row1 = __ldg(rs.srow0 + c.k + 1);
	c.sstart = __ldg(rs.sstart + c.k);
	c.send = __ldg(rs.send + c.k);
	c.chrom = __ldg(rs.schrom + c.k);
	c.bin_base = __ldg(rs.sbin0 + c.k) - c.row0;
}

__device__ __forceinline__ void mg_get_row_cached(const RowSrc &rs, int64_t r, RowCache &c, int &chrom, int64_t &s, int64_t &e, int64_t &scope)
{
	if (rs.kind == 0) {
		chrom = __ldg(rs.chrom + r);
		s = __ldg(rs.start + r);
		e = __ldg(rs.end + r);
		scope = rs.scope_idx ? __ldg(rs.scope_idx + r) : -1;
		return;
	}
	if (c.k < 0 || r < c.row0 || r >= c.row1)
		mg_rowcache_load(rs, r, c);
	const int64_t coord = (c.bin_base + r) * rs.binsize;
	chrom = c.chrom;
	s = coord > c.sstart ? coord : c.sstart;
	e = coord + rs.binsize < c.send ? coord + rs.binsize : c.send;
	scope = c.k;
}

__device__ __forceinline__ void mg_get_row(const RowSrc &rs, int64_t r, int &chrom, int64_t &s, int64_t &e, int64_t &scope)
{
	RowCache c;
	mg_get_row_cached(rs, r, c, chrom, s, e, scope);
}
