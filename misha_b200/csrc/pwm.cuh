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
	const uint32_t *lower; // bit k of word t: base 32 t + k is a lower-case (soft-masked) letter
};

struct mg_seq {
	int nchroms = 0;
	bool has_other = false; // some staged character is neither ACGT nor N / * (see k_seq_pack)
	std::vector<SeqChrom> h_table;
	SeqChrom *d_table = nullptr;
	std::vector<void *> allocs;
	int64_t bytes = 0;
};

// one thread per 32 bases
// *other is set when a character is neither ACGT nor neutral (N, n, *: DnaLookupTables::NEUTRAL_CHAR, src/DnaPSSM.h:21): pwm.max.pos
// scores those two kinds of non-ACGT characters differently (src/DnaPSSM.cpp:306-316) and the packed form keeps one bit for both.
__global__ void __launch_bounds__(256) k_seq_pack(const uint8_t *__restrict__ ascii, int64_t len, uint32_t *codes, uint32_t *inval, int *other)
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
			else {
				inv = 1;
				if (ch != 'N' && ascii[p] != '*')
					*other = 1;
			}
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
#define MG_PWM_MAX_RANGE 1000000 // anchors 0..10^6 of the target are scored, the rest ignored (DnaPSSM::m_max_range)
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
	// spatial weights (PWMParams spat_factor / spat_bin, src/PWMScorer.cpp:112-165): log(max(1e-30f, factor)) per bin of spat_bin TARGET
	// positions, the last bin extended over whatever follows; nspat = 0: none
	const float *spat_log;
	int nspat, spat_bin;
};
#define MG_PWM_MAXSPAT 2048

// src/util.h:16-28. STRICT: expf / logf taken as the correctly rounded values (double evaluation rounded to float; glibc's
// are correctly rounded in all but rare cases) -- two double-precision transcendentals per anchor, which is what bounds the
// kernel. Default: CUDA's float expf (<= 2 ulp) and logf (<= 1 ulp): the strand sum moves by <= ~3e-7 absolute, i.e.
// ~1e-8 relative to a 20-mer's log-likelihood, far inside the 1e-6 bar (mg_set_option("pwm_strict", 1) restores STRICT).
__device__ __forceinline__ uint32_t mg_lds_u32(uint32_t a)
{
	uint32_t v;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
	return v;
}
__device__ __forceinline__ void mg_lds_f2(uint32_t a, float &x, float &y)
{
	asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x), "=f"(y) : "r"(a));
}

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
// log(exp(l1) + exp(l2)), src/util.h:16-28, without its branches: with hi = max, lo = min the reference returns
// hi + log1pexp(lo - hi), skipping the term when lo is -inf -- where it is exactly 0 anyway (exp(-inf) = 0, log(1) = 0).
// Only hi = -inf (both strands impossible: lo - hi = NaN) needs a guard. Same values, no divergence between lanes.
template <bool STRICT>
__device__ __forceinline__ float mg_log_sum_log(float l1, float l2)
{
	const float hi = fmaxf(l1, l2), lo = fminf(l1, l2);
	const float r = __fadd_rn(hi, mg_log1pexp_f<STRICT>(__fsub_rn(lo, hi)));
	return hi == __int_as_float(0xff800000) ? hi : r;
}

// BID: both strands combined per anchor; MODE: MG_PWM_TOTAL / MAX / COUNT; STRICT: see mg_log1pexp_f
#ifndef MG_PWM_MINBLOCKS
#define MG_PWM_MINBLOCKS 3
#endif
template <bool BID, int MODE, bool STRICT>
__global__ void __launch_bounds__(MG_PWM_WARPS * 32, MG_PWM_MINBLOCKS) k_pwm(const PwmQuery q)
{
	__shared__ float2 stab[MG_PWM_MAXL * 5];
	__shared__ float sspat[MODE == MG_PWM_MAX_POS ? 1 : MG_PWM_MAXSPAT];
	__shared__ __align__(16) uint8_t stile[MG_PWM_WARPS][MG_PWM_TILE + 16];
	__shared__ uint32_t spk[MG_PWM_WARPS][3 * ((MG_PWM_TILE + 16 + 31) / 32 + 2)]; // {inval, codes lo, codes hi} per 32 bases
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int L = q.L, Lpad = (L + 3) & ~3; // motif rows padded to a multiple of 4 with zeros: adding 0.f changes nothing
	for (int t = threadIdx.x; t < Lpad * 5; t += blockDim.x)
		stab[t] = t < L * 5 ? q.tab[t] : make_float2(0.f, 0.f);
	if (MODE != MG_PWM_MAX_POS)
		for (int t = threadIdx.x; t < q.nspat; t += blockDim.x)
			sspat[t] = q.spat_log[t];
	__syncthreads();
	const uint32_t stab_s = (uint32_t)__cvta_generic_to_shared(stab);
	uint8_t *tile = stile[warp];
	uint32_t *wpk = spk[warp];
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
		int64_t na = 0, xe = 0, a_base = 0;
		if (mg_window(s, e, q.sshift, q.eshift, chromsize, ws, we)) {
			xe = we;
			if (q.extend && L > 1) { // src/GenomeSeqScorer.cpp:16-38: extend the END only
				xe = we + (L - 1);
				if (xe > chromsize)
					xe = chromsize;
			}
			const int64_t tlen = xe - ws;
			if (tlen < L) {
				if (q.extend) { // score_without_spatial over zero anchors (src/PWMScorer.cpp:313-338)
					result = MODE == MG_PWM_COUNT ? 0. : (double)NEG_INF;
					if (MODE == MG_PWM_MAX_POS) {
						// max_like_match returns the target's begin and leaves best_dir unset (src/DnaPSSM.cpp:288-291): index 0 through
						// compute_position_result; with both strands the sign is the reference's uninitialised int -> NaN here
						const float pos = rev ? __fadd_rn(__fsub_rn(__fsub_rn((float)tlen, 1.0f), (float)L), 1.0f) : 1.0f;
						result = BID ? mg_nan() : (double)pos;
					}
				}
			} else {
				int64_t i_max = tlen - L;
				const int64_t want = we - ws - 1;
				if (want < i_max)
					i_max = want;
				if (i_max > MG_PWM_MAX_RANGE) // DnaPSSM's default scan range, m_max_range (src/DnaPSSM.h:186-187)
					i_max = MG_PWM_MAX_RANGE;
				na = i_max + 1;
				// strand -1 scores the reverse complement: its first na anchors are the LAST na genomic ones
				a_base = rev ? (tlen - L + 1) - na : 0;
				scan = true;
			}
		}
		if (scan) {
			const SeqChrom ch = q.table[chrom];
			double m_run = -INFINITY, s_run = 0; // online log-sum-exp per lane
			float vmax = NEG_INF;
			long long hits = 0;
			int best_i = -1, best_dir = 1; // pwm.max.pos: index in the scored target (the reverse complement for strand -1)
			for (int64_t a0 = 0; a0 < na; a0 += MG_PWM_CHUNK) {
				// decode bases [ws + a0, ws + a0 + CHUNK + L + 2) into the tile: the packed words the chunk spans are fetched once
				// (one load per lane), every base is then a shift and a mask in 32-bit arithmetic
				const int need = MG_PWM_CHUNK + Lpad - 1 + 3;
				const int64_t p0 = ws + a_base + a0, pw0 = p0 & ~31ll; // chunk start, and the 32-base boundary below it
				const int off = (int)(p0 - pw0), span = off + need; // bases from pw0 on
				__syncwarp();
				for (int t = lane; t < ((span + 31) >> 5); t += 32) { // inval word t, codes words 2t and 2t + 1
					const int64_t wi = (pw0 >> 5) + t;
					const bool in = (wi << 5) < ch.len;
					wpk[3 * t] = in ? __ldg(ch.inval + wi) : 0xffffffffu;
					wpk[3 * t + 1] = in ? __ldg(ch.codes + 2 * wi) : 0u;
					wpk[3 * t + 2] = in && (wi << 5) + 16 < ch.len ? __ldg(ch.codes + 2 * wi + 1) : 0u;
				}
				__syncwarp();
				const int lim = (int)min((int64_t)0x7fffffff, xe - pw0); // bases at or past xe read as non-ACGT
				for (int t = lane; t < need; t += 32) {
					const int b = off + t, w = b >> 5, k = b & 31;
					const uint32_t bad = (wpk[3 * w] >> k) & 1u;
					const uint32_t code = (wpk[3 * w + 1 + (k >> 4)] >> (2 * (k & 15))) & 3u;
					tile[t] = (uint8_t)(8u * ((bad || b >= lim) ? 4u : code)); // byte offset of the table column
				}
				__syncwarp();
				float f[4] = {0.f, 0.f, 0.f, 0.f}, r[4] = {0.f, 0.f, 0.f, 0.f};
				if (!rev) {
					// the lane's 4 anchors read bases 4 lane + j + u: one aligned word of the tile per 4 motif positions, byte
					// offsets known at compile time; a lookup is PRMT (column offset) + add + LDS.64 + 2 FADD
					const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(tile) + 4u * lane;
					uint32_t w0 = mg_lds_u32(tile_s), rowb = stab_s;
					for (int j0 = 0; j0 < Lpad; j0 += 4, rowb += 4 * 40) {
						const uint32_t w1 = mg_lds_u32(tile_s + j0 + 4);
#pragma unroll
						for (int jj = 0; jj < 4; ++jj) {
#pragma unroll
							for (int u = 0; u < 4; ++u) {
								// byte jj + u of the 8 bytes {w0, w1}, zero-extended: one PRMT (selector 4 = a byte of the zero operand)
								const uint32_t off = __byte_perm(jj + u < 4 ? w0 : w1, 0u, 0x4440 + ((jj + u) & 3));
								float x, y;
								mg_lds_f2(rowb + off + jj * 40, x, y);
								f[u] = __fadd_rn(f[u], x);
								r[u] = __fadd_rn(r[u], y);
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
							const float2 tt = row[tb[j + u] >> 3];
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
					if (MODE == MG_PWM_MAX_POS) {
						// DnaPSSM::max_like_match with combine_strands = false (src/DnaPSSM.cpp:285-365): the better strand of the anchor
						// (the reverse one only when strictly better), the first strictly greater anchor in target order wins
						const float l = rev ? r[u] : f[u], rl = rev ? f[u] : r[u];
						int d = 1;
						v = l;
						if (BID && rl > l) {
							v = rl;
							d = -1;
						}
						// target index: strand -1 scores the reverse complement, whose index runs against the genomic anchor; a lane
						// meets its anchors in genomic order, so there an equal later anchor (smaller target index) replaces the held one
						const int ti = rev ? (int)(na - 1 - a) : (int)a;
						if (v > vmax || (rev && v == vmax && v != NEG_INF)) {
							vmax = v;
							best_i = ti;
							best_dir = d;
						}
						continue;
					}
					if (MODE != MG_PWM_MAX_POS && q.nspat) {
						// the anchor's spatial bin by its TARGET index (integrate_energy_logspat / _max_logspat, src/DnaPSSM.cpp:933-1075:
						// each strand weighted, then united; count_motif_hits_with_spatial, src/PWMScorer.cpp:219-249: united, then weighted)
						const int ti = rev ? (int)(na - 1 - a) : (int)a, b = ti / q.spat_bin;
						const float sl = sspat[b < q.nspat ? b : q.nspat - 1];
						if (BID) {
							v = MODE == MG_PWM_COUNT ? __fadd_rn(mg_log_sum_log<STRICT>(f[u], r[u]), sl)
													 : mg_log_sum_log<STRICT>(__fadd_rn(f[u], sl), __fadd_rn(r[u], sl));
						} else
							v = __fadd_rn(rev ? r[u] : f[u], sl);
						// pwm.count: the reference's scorer answers from spat_seed, whose hit counts stop at the end of the last spatial bin
						// (src/PWMScorer.cpp:1098-1111) although pwm / pwm.max extend that bin over the rest (:934-937)
						if (MODE == MG_PWM_COUNT && b >= q.nspat)
							v = NEG_INF;
					} else if (BID)
						v = mg_log_sum_log<STRICT>(f[u], r[u]);
					else
						v = rev ? r[u] : f[u];
					if (MODE == MG_PWM_TOTAL) {
						// src/utils/RunningLogSumExp.h:32-45 as selects: the new maximum rescales the running sum, anything else adds
						// its own term; an impossible anchor (-inf) adds nothing
						const double dv = (double)v;
						const bool up = dv > m_run;
						double ex;
						if (STRICT)
							ex = exp(-fabs(dv - m_run));
						else // each term good to ~2.4e-7 relative, the log of their sum to the same absolute error
							ex = (double)expf(-(float)fabs(dv - m_run));
						if (v != NEG_INF) {
							s_run = up ? s_run * ex + 1. : s_run + ex;
							m_run = up ? dv : m_run;
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
			} else if (MODE == MG_PWM_MAX_POS) {
#pragma unroll
				for (int k = 16; k > 0; k >>= 1) { // larger score, then smaller target index
					const float ov = __shfl_xor_sync(0xffffffffu, vmax, k);
					const int oi = __shfl_xor_sync(0xffffffffu, best_i, k), od = __shfl_xor_sync(0xffffffffu, best_dir, k);
					if (oi >= 0 && (best_i < 0 || ov > vmax || (ov == vmax && oi < best_i))) {
						vmax = ov;
						best_i = oi;
						best_dir = od;
					}
				}
				if (best_i >= 0) { // PWMScorer::compute_position_result (src/PWMScorer.cpp:168-182), in float like the reference
					float pos = __fadd_rn((float)best_i, 1.0f);
					if (rev)
						pos = __fadd_rn(__fsub_rn(__fsub_rn((float)(xe - ws), pos), (float)L), 1.0f);
					if (BID)
						pos = pos * (float)best_dir;
					result = (double)pos;
				}
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

// ---- kmer.count / kmer.frac (KmerCounter::score_interval, src/KmerCounter.cpp:38-176) -------------------------------------------
// A warp per row. The reference fetches [ws, xe) (xe = the window's end extended by k - 1 when `extend`, clipped to the chromosome:
// src/GenomeSeqScorer.cpp:16-38), reverse-complements it for the reverse strand, and compares the k-mer at every start inside the
// original window that still fits. Both strands visit the same genomic starts g in [ws, xe - k]: the forward strand matches the
// k-mer at g, the reverse strand its reverse complement. A start is one 2k-bit compare on the packed codes plus the invalid bits
// (a k-mer of ACGT never matches N or any other character).
struct KmerQuery {
	RowSrc rows;
	int64_t first, count;
	int64_t sshift, eshift;
	const SeqChrom *table;
	const int64_t *chrom_sizes;
	uint64_t fwd, rev; // base j of the k-mer (of its reverse complement) at bits 2j
	int k, mode, extend, strand;
	double *out;
};

__global__ void __launch_bounds__(256) k_kmer(const KmerQuery q)
{
	const int lane = threadIdx.x & 31;
	const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
	const int k = q.k;
	const uint64_t cmask = k < 32 ? (1ull << (2 * k)) - 1ull : ~0ull;
	const uint32_t imask = k < 32 ? (1u << k) - 1u : ~0u;
	for (int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < q.count; i += nwarps) {
		int chrom;
		int64_t s, e, scope;
		mg_get_row(q.rows, q.first + i, chrom, s, e, scope);
		const int64_t chromsize = __ldg(q.chrom_sizes + chrom);
		int64_t ws, we;
		double result = mg_nan();
		if (mg_window(s, e, q.sshift, q.eshift, chromsize, ws, we)) {
			int64_t xe = we;
			if (q.extend && k > 1) {
				xe = we + (k - 1);
				if (xe > chromsize)
					xe = chromsize;
			}
			unsigned nf = 0, nr = 0;
			const bool fits = xe - ws >= k;
			if (fits) {
				const SeqChrom ch = q.table[chrom];
				const unsigned long long *codes = (const unsigned long long *)ch.codes;
				const int64_t nwords = (ch.len + 31) >> 5;
				const int64_t last = min(we - 1, xe - k);
				for (int64_t g = ws + lane; g <= last; g += 32) {
					const int64_t w = g >> 5;
					const int off = (int)(g & 31);
					const bool more = off && w + 1 < nwords;
					const unsigned long long lo = __ldg(codes + w), hi = more ? __ldg(codes + w + 1) : 0ull;
					const uint32_t ilo = __ldg(ch.inval + w), ihi = more ? __ldg(ch.inval + w + 1) : 0xffffffffu;
					const unsigned long long code = (off ? (lo >> (2 * off)) | (hi << (64 - 2 * off)) : lo) & cmask;
					const uint32_t inv = (off ? (ilo >> off) | (ihi << (32 - off)) : ilo) & imask;
					nf += inv == 0 && code == q.fwd;
					nr += inv == 0 && code == q.rev;
				}
			}
			nf = __reduce_add_sync(0xffffffffu, nf);
			nr = __reduce_add_sync(0xffffffffu, nr);
			const unsigned long long total = (q.strand >= 0 ? nf : 0u) + (unsigned long long)(q.strand <= 0 ? nr : 0u);
			if (q.mode == MG_KMER_FRAC) {
				const unsigned long long possible = fits ? (unsigned long long)(we - ws) : 0ull;
				unsigned long long valid = possible > (unsigned long long)(k - 1) ? possible - (unsigned long long)(k - 1) : 0ull;
				if (q.strand == 0)
					valid *= 2;
				result = valid > 0 ? (double)__fdiv_rn((float)total, (float)valid) : 0.;
			} else
				result = (double)(float)total;
		}
		if (lane == 0)
			q.out[i] = result;
	}
}

// ---- masked.count / masked.frac (MaskedBpCounter::score_interval, src/MaskedBpCounter.cpp:19-64) ---------------------------------
// The lower-case bit of every base, one thread per 32 bases (staging) ...
__global__ void __launch_bounds__(256) k_seq_lower(const uint8_t *__restrict__ ascii, int64_t len, uint32_t *lower)
{
	const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const int64_t p0 = t * 32;
	if (p0 >= len)
		return;
	uint32_t m = 0;
#pragma unroll 8
	for (int k = 0; k < 32; ++k) {
		const int64_t p = p0 + k;
		if (p < len) {
			const uint8_t ch = ascii[p];
			m |= (uint32_t)(ch >= 'a' && ch <= 'z') << k; // islower in the C locale
		}
	}
	lower[t] = m;
}

// ... and a warp per row counting them over the window's words: no extension, no strand; the fraction is float / size_t -> float.
struct MaskedQuery {
	RowSrc rows;
	int64_t first, count;
	int64_t sshift, eshift;
	const SeqChrom *table;
	const int64_t *chrom_sizes;
	int mode;
	double *out;
};

__global__ void __launch_bounds__(256) k_masked(const MaskedQuery q)
{
	const int lane = threadIdx.x & 31;
	const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
	for (int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < q.count; i += nwarps) {
		int chrom;
		int64_t s, e, scope;
		mg_get_row(q.rows, q.first + i, chrom, s, e, scope);
		int64_t ws, we;
		double result = mg_nan();
		if (mg_window(s, e, q.sshift, q.eshift, __ldg(q.chrom_sizes + chrom), ws, we)) {
			unsigned n = 0;
			if (we > ws) {
				const uint32_t *lower = q.table[chrom].lower;
				const int64_t w0 = ws >> 5, w1 = (we - 1) >> 5;
				for (int64_t w = w0 + lane; w <= w1; w += 32) {
					uint32_t m = __ldg(lower + w);
					if (w == w0)
						m &= 0xffffffffu << (ws & 31);
					if (w == w1)
						m &= 0xffffffffu >> (31 - ((we - 1) & 31));
					n += __popc(m);
				}
			}
			n = __reduce_add_sync(0xffffffffu, n);
			const unsigned long long len = (unsigned long long)(we - ws);
			if (len == 0)
				result = 0.;
			else
				result = q.mode == MG_MASKED_FRAC ? (double)__fdiv_rn((float)n, (float)len) : (double)(float)n;
		}
		if (lane == 0)
			q.out[i] = result;
	}
}
