// rows.cuh -- iterator rows on the device.
//   explicit rows : SoA (chrom int32, start int64, end int64) -- the intervals iterator / uploaded batches
//   implicit rows : fixed-bin iterator over a sorted scope; row r of scope interval k covers
//                   [max(bin*b, s_k), min((bin+1)*b, e_k)), bin = floor(s_k / b) + (r - row0_k)
//                   (src/TrackExpressionFixedBinIterator.cpp:52-60). Never materialised.
#pragma once
#include "common.cuh"

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
};

struct mg_rows {
	RowSrc src{};
	bool owns = false;
	void *d_block = nullptr; // one allocation backing every owned array
};

// scope interval of implicit row r: last k with srow0[k] <= r
__device__ __forceinline__ int64_t mg_scope_of_row(const RowSrc &rs, int64_t r)
{
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

__device__ __forceinline__ void mg_get_row(const RowSrc &rs, int64_t r, int &chrom, int64_t &s, int64_t &e, int64_t &scope)
{
	if (rs.kind == 0) {
		chrom = __ldg(rs.chrom + r);
		s = __ldg(rs.start + r);
		e = __ldg(rs.end + r);
		scope = rs.scope_idx ? __ldg(rs.scope_idx + r) : -1;
	} else {
		int64_t k = mg_scope_of_row(rs, r);
		int64_t ss = __ldg(rs.sstart + k), se = __ldg(rs.send + k);
		int64_t bin = ss / rs.binsize + (r - __ldg(rs.srow0 + k));
		int64_t coord = bin * rs.binsize;
		chrom = __ldg(rs.schrom + k);
		s = coord > ss ? coord : ss;
		e = coord + rs.binsize < se ? coord + rs.binsize : se;
		scope = k;
	}
}
