// pwm.cuh -- PWM (PSSM log-likelihood) vtracks over 2-bit packed sequence.
//
// HBM layout per chromosome: codes[len/16] uint32 (16 bases x 2 bits, A=0 C=1 G=2 T=3, case-insensitive as
// DnaLookupTables::BASE_ENCODE src/DnaPSSM.h:15-22) + inval[len/32] uint32 bitmask (1 = not ACGT, e.g. N).
// Scoring (src/PWMScorer.cpp:742-861 non-spatial path, src/DnaPSSM.cpp:249-283): one warp per row; the row's bases are
// decoded into a shared-memory tile 128 anchors at a time, each lane scores 4 consecutive anchors on both strands with
// float adds in the reference's order (table {fwd, rc} float2 in shared memory, column 4 = -inf for non-ACGT),
// strands are combined with the reference's float log_sum_log, anchors are reduced per mode:
//   pwm       log-sum-exp in double with a running maximum (RunningLogSumExp::init/value), rounded to float
//   pwm.max   float maximum          pwm.count   #anchors >= threshold
// Not HBM-bound (~130 packed bytes per 500 bp row against 2 x 20 table adds per anchor): shared-memory/ALU bound.
#pragma once
#include "common.cuh"
#include "rows.cuh"

struct SeqChrom {
	const uint32_t *codes;
	const uint32_t *inval;
	int64_t len;
};

struct mg_seq {
	int nchroms = 0;
	std::vector<SeqChrom> h_table;
	SeqChrom *d_table = nullptr;
	std::vector<void *> allocs;
	int64_t bytes = 0;
};

// one thread per 32 bases
__global__ void __launch_bounds__(256) k_seq_pack(const uint8_t *__restrict__ ascii, int64_t len, uint32_t *codes, uint32_t *inval)
{
	const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const int64_t p0 = t * 32;
	if (p0 >= len)
		return;
	uint32_t c0 = 0, c1 = 0, bad = 0;
#pragma unroll 8
	for (int k = 0; k < 32; ++k) {
		const int64_t p = p0 + k;
		uint32_t code = 0, inv = 1;
		if (p < len) {
			const uint8_t ch = ascii[p] & 0xdf; // fold case
			inv = 0;
			if (ch == 'A') code = 0;
			else if (ch == 'C') code = 1;
			else if (ch == 'G') code = 2;
			else if (ch == 'T') code = 3;
			else inv = 1;
		}
		if (k < 16)
			c0 |= code << (2 * k);
		else
			c1 |= code << (2 * (k - 16));
		bad |= inv << k;
	}
	codes[2 * t] = c0;
	codes[2 * t + 1] = c1;
	inval[t] = bad;
}

#define MG_PWM_MAXL 128
#define MG_PWM_WARPS 8
#define MG_PWM_CHUNK 128 // anchors per warp pass (4 per lane)
#define MG_PWM_TILE (MG_PWM_CHUNK + MG_PWM_MAXL)

struct PwmQuery {
	RowSrc rows;
	int64_t first, count;
	int64_t sshift, eshift;
	const SeqChrom *table;
	const int64_t *chrom_sizes;
	const float2 *tab; // device, L x 5 {fwd logp[j][c], rc logp[L-1-j][3-c]}; c == 4 -> -inf
	int L;
	int bidirect, extend, mode, strand;
	float thresh;
	double *out;
};

// src/util.h:16-28. STRICT: expf / logf taken as the correctly rounded values (double evaluation rounded to float; glibc's
// are correctly rounded in all but rare cases) -- two double-precision transcendentals per anchor, which is what bounds the
// kernel. Default: CUDA's float expf (<= 2 ulp) and logf (<= 1 ulp): the strand sum moves by <= ~3e-7 absolute, i.e.
// ~1e-8 relative to a 20-mer's log-likelihood, far inside the 1e-6 bar (mg_set_option("pwm_strict", 1) restores STRICT).
template <bool STRICT>
__device__ __forceinline__ float mg_log1pexp_f(float d)
{
	if (STRICT) {
		const float t = (float)exp((double)d);
		const float u = __fadd_rn(1.0f, t);
		return (float)log((double)u);
	}
	return logf(__fadd_rn(1.0f, expf(d)));
}
template <bool STRICT>
__device__ __forceinline__ float mg_log_sum_log(float l1, float l2)
{
	if (l1 > l2) {
		if (!isinf(l2))
			l1 = __fadd_rn(l1, mg_log1pexp_f<STRICT>(__fsub_rn(l2, l1)));
		return l1;
	}
	if (isinf(l1))
		return l2;
	return __fadd_rn(l2, mg_log1pexp_f<STRICT>(__fsub_rn(l1, l2)));
}

// BID: both strands combined per anchor; MODE: MG_PWM_TOTAL / MAX / COUNT; STRICT: see mg_log1pexp_f
#ifndef MG_PWM_MINBLOCKS
#define MG_PWM_MINBLOCKS 1
#endif
template <bool BID, int MODE, bool STRICT>
__global__ void __launch_bounds__(MG_PWM_WARPS * 32, MG_PWM_MINBLOCKS) k_pwm(const PwmQuery q)
{
	__shared__ float2 stab[MG_PWM_MAXL * 5];
	__shared__ __align__(16) uint8_t stile[MG_PWM_WARPS][MG_PWM_TILE + 16];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int L = q.L;
	for (int t = threadIdx.x; t < L * 5; t += blockDim.x)
		stab[t] = q.tab[t];
	__syncthreads();
	uint8_t *tile = stile[warp];
	const float NEG_INF = __int_as_float(0xff800000);
	const bool rev = q.strand == -1; // the reference scores the reverse-complemented target: terms are added j = L-1..0
	const int64_t nwarps_total = (int64_t)gridDim.x * MG_PWM_WARPS;
	for (int64_t i = (int64_t)blockIdx.x * MG_PWM_WARPS + warp; i < q.count; i += nwarps_total) {
		int chrom;
		int64_t s, e, scope;
		mg_get_row(q.rows, q.first + i, chrom, s, e, scope);
		int64_t ws, we;
		const int64_t chromsize = __ldg(q.chrom_sizes + chrom);
		double result = mg_nan();
		bool scan = false;
		int64_t na = 0, xe = 0;
		if (mg_window(s, e, q.sshift, q.eshift, chromsize, ws, we)) {
			xe = we;
			if (q.extend && L > 1) { // src/GenomeSeqScorer.cpp:16-38: extend the END only
				xe = we + (L - 1);
				if (xe > chromsize)
					xe = chromsize;
			}
			const int64_t tlen = xe - ws;
			if (tlen < L) {
				if (q.extend) // score_without_spatial over zero anchors (src/PWMScorer.cpp:313-338)
					result = MODE == MG_PWM_COUNT ? 0. : (double)NEG_INF;
			} else {
				int64_t i_max = tlen - L;
				const int64_t want = we - ws - 1;
				if (want < i_max)
					i_max = want;
				na = i_max + 1;
				scan = true;
			}
		}
		if (scan) {
			const SeqChrom ch = q.table[chrom];
			double m_run = -INFINITY, s_run = 0; // online log-sum-exp per lane
			float vmax = NEG_INF;
			long long hits = 0;
			for (int64_t a0 = 0; a0 < na; a0 += MG_PWM_CHUNK) {
				// decode bases [ws + a0, ws + a0 + CHUNK + L - 1) into the tile
				const int need = MG_PWM_CHUNK + L - 1;
				__syncwarp();
				for (int t = lane; t < need + 3; t += 32) {
					const int64_t p = ws + a0 + t;
					uint8_t code = 4;
					if (p < xe) {
						const uint32_t w = __ldg(ch.codes + (p >> 4));
						const uint32_t bad = __ldg(ch.inval + (p >> 5));
						code = ((bad >> (p & 31)) & 1u) ? 4 : (uint8_t)((w >> (2 * (p & 15))) & 3u);
					}
					tile[t] = code;
				}
				__syncwarp();
				float f[4] = {0.f, 0.f, 0.f, 0.f}, r[4] = {0.f, 0.f, 0.f, 0.f};
				if (!rev) {
					// the lane's 4 anchors read bases 4 lane + j + u: one aligned word of the tile per 4 motif positions, byte
					// offsets known at compile time
					const uint32_t *tw = reinterpret_cast<const uint32_t *>(tile) + lane;
					uint32_t w0 = tw[0];
					for (int j0 = 0; j0 < L; j0 += 4) {
						const uint32_t w1 = tw[(j0 >> 2) + 1];
#pragma unroll
						for (int jj = 0; jj < 4; ++jj) {
							if (j0 + jj < L) {
								const float2 *row = stab + (j0 + jj) * 5;
#pragma unroll
								for (int u = 0; u < 4; ++u) {
									const int b = jj + u; // byte b of the 8 bytes {w0, w1}
									const uint32_t code = ((b < 4 ? w0 >> (8 * b) : w1 >> (8 * (b - 4))) & 0xffu);
									const float2 tt = row[code];
									f[u] = __fadd_rn(f[u], tt.x);
									r[u] = __fadd_rn(r[u], tt.y);
								}
							}
						}
						w0 = w1;
					}
				} else {
					const uint8_t *tb = tile + 4 * lane;
					for (int k = 0; k < L; ++k) {
						const int j = L - 1 - k;
						const float2 *row = stab + j * 5;
#pragma unroll
						for (int u = 0; u < 4; ++u) {
							const float2 tt = row[tb[j + u]];
							f[u] = __fadd_rn(f[u], tt.x);
							r[u] = __fadd_rn(r[u], tt.y);
						}
					}
				}
#pragma unroll
				for (int u = 0; u < 4; ++u) {
					const int64_t a = a0 + 4 * lane + u;
					if (a >= na)
						continue;
					float v;
					if (BID)
						v = mg_log_sum_log<STRICT>(f[u], r[u]);
					else
						v = rev ? r[u] : f[u];
					if (MODE == MG_PWM_TOTAL) {
						if (v != NEG_INF) {
							const double dv = (double)v;
							if (STRICT) {
								if (dv > m_run) {
									s_run = s_run * exp(m_run - dv) + 1.;
									m_run = dv;
								} else
									s_run += exp(dv - m_run);
							} else {
								// same recurrence (src/utils/RunningLogSumExp.h:32-45) with the exponential in float: each term is
								// good to ~2.4e-7 relative, the log of their sum to the same absolute error
								const float d = (float)fabs(dv - m_run);
								const double ex = (double)expf(-d);
								if (dv > m_run) {
									s_run = s_run * ex + 1.;
									m_run = dv;
								} else
									s_run += ex;
							}
						}
					} else if (MODE == MG_PWM_MAX)
						vmax = fmaxf(vmax, v);
					else
						hits += v >= q.thresh ? 1 : 0;
				}
			}
			if (MODE == MG_PWM_TOTAL) {
				double M = m_run;
#pragma unroll
				for (int k = 16; k > 0; k >>= 1)
					M = fmax(M, __shfl_xor_sync(0xffffffffu, M, k));
				double sc = (m_run == -INFINITY) ? 0. : s_run * exp(m_run - M);
				sc = mg_warp_sum_d(sc);
				result = (M == -INFINITY) ? (double)NEG_INF : mg_f2d(M + log(sc));
			} else if (MODE == MG_PWM_MAX) {
#pragma unroll
				for (int k = 16; k > 0; k >>= 1)
					vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, k));
				result = (double)vmax;
			} else {
#pragma unroll
				for (int k = 16; k > 0; k >>= 1)
					hits += __shfl_xor_sync(0xffffffffu, hits, k);
				result = (double)(float)hits;
			}
		}
		if (lane == 0)
			mg_st_stream(q.out + i, result);
	}
}
