/*
 * scan_v3.cuh -- warp-autonomous promiscuous access-code scan (the bulk kernel).
 *
 * Same decision per window as promiscuous_packet_search (bluetooth_packet.c:368-420), laid
 * out for the SM's issue and shared-memory budgets:
 *
 *   load     a warp owns a contiguous run of 4096-symbol strips and never meets a block
 *            barrier; lane L pulls symbols [32(32k+L), +32), k = 0..3, with one 256-bit load
 *            each (1 KiB contiguous per warp instruction) plus a 64-symbol halo;
 *   pack     32 symbols -> one word with 8 IDP.4A (byte dot products with 1,2,4,8 /
 *            16,..,128) and 3 IMAD -- FMA pipe only, the ALU pipe is kept for the logic;
 *   filter   the Barker-tail test (BARKER_DISTANCE <= 1, :385) runs bit-sliced on 32
 *            positions per lane: 7 funnel shifts, 6 LOP3;
 *   compact  the ~1/8 surviving positions are written to a per-warp queue (row-major, so the
 *            consumer's shared-memory reads stay conflict-free) after a packed warp scan;
 *   test     every lane takes queue entries (all lanes busy, two independent candidates in
 *            flight): window extraction, low 32 syndrome bits from two LUTs over codeword
 *            bits 32..44 and 45..56 (gen_syndrome, :147-159, regrouped), one probe of a
 *            2^19-bit map holding the reachable syndromes of both Barker classes
 *            (find_syndrome, :139-145, as a filter);
 *   exact    the few map positives go through exact_promisc() (the reference's test, full
 *            34-bit syndrome, error lookup, LAP) from a small second queue.
 */
#pragma once

namespace v3 {

constexpr int WARPS = 32;            /* warps per CTA, one CTA per SM */
constexpr int K = 4;                 /* rows: words (32 positions each) per lane per strip */
constexpr int SW = 32 * K;           /* words per strip */
constexpr int STRIP = SW * 32;       /* symbols per strip */
constexpr int BLOG = 19;             /* log2 bits of the syndrome map */
constexpr int QCAP = 1024;           /* queue entries (u16); a row never holds more */
constexpr int XCAP = 30;             /* map-positive queue per warp (flushed when half full) */
constexpr int LUTA_BITS = 13, LUTB_BITS = 12;
constexpr int LUT_WORDS = (1 << LUTA_BITS) + (1 << LUTB_BITS);
constexpr int MAP_WORDS = 1 << (BLOG - 5);
constexpr int S_WORDS = SW + 8;
constexpr int X_WORDS = 128;          /* [0] count, then XCAP x {pos_lo, pos_hi, lo, hi} */
constexpr int WARP_WORDS = S_WORDS + QCAP / 2 + X_WORDS;
constexpr size_t SMEM_BYTES = (size_t)(LUT_WORDS + MAP_WORDS + WARPS * WARP_WORDS) * 4;

struct args {
	const uint8_t *base;     /* 32-byte aligned; base[0] is stream position pos0 */
	int64_t pos0;
	int64_t nstrips;         /* strips [0, nstrips); base[nstrips*STRIP + 63] is readable */
	const uint32_t *lut;     /* LUT_WORDS: A then B */
	const uint32_t *map;     /* MAP_WORDS */
	uint64_t cc[2];          /* 34-bit syndrome of PN ^ (legal tail << 57), tail A / tail B */
	uint32_t m32, m33;       /* codeword bits 32..56 (as bits of `hi`) feeding syndrome bits 32 / 33 */
	scan_args common;
};

__device__ __forceinline__ void ld256(const uint8_t *p, uint32_t r[8])
{
	asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
		     : "l"(p));
}

/* 32 symbols (one byte each, 0/1) -> 32 bits, symbol i -> bit i */
__device__ __forceinline__ uint32_t pack32(const uint32_t r[8])
{
	uint32_t b0 = __dp4a(r[1], 0x80402010u, __dp4a(r[0], 0x08040201u, 0u));
	uint32_t b1 = __dp4a(r[3], 0x80402010u, __dp4a(r[2], 0x08040201u, 0u));
	uint32_t b2 = __dp4a(r[5], 0x80402010u, __dp4a(r[4], 0x08040201u, 0u));
	uint32_t b3 = __dp4a(r[7], 0x80402010u, __dp4a(r[6], 0x08040201u, 0u));
	return b0 + (b1 << 8) + (b2 << 16) + (b3 << 24);
}

__device__ __forceinline__ uint32_t lds_off(const uint32_t *base, uint32_t byteoff)
{
	return *reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(base) + byteoff);
}

__device__ __forceinline__ uint32_t bfind(uint32_t x)
{
	uint32_t r;
	asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(x));
	return r;
}

/* Barker tail within distance 1 of either legal tail, for the 32 positions whose windows
 * start in the word before w1 (tail bits 57..63 of position i are bits i+25.. of w2:w1) */
__device__ __forceinline__ uint32_t barker_mask(uint32_t w1, uint32_t w2)
{
	const uint32_t x0 = ~__funnelshift_r(w1, w2, 25), x1 = ~__funnelshift_r(w1, w2, 26),
		       x2 = ~__funnelshift_r(w1, w2, 27), x3 = __funnelshift_r(w1, w2, 28),
		       x4 = __funnelshift_r(w1, w2, 29),  x5 = ~__funnelshift_r(w1, w2, 30),
		       x6 = __funnelshift_r(w1, w2, 31);
	const uint32_t s1 = x0 ^ x1 ^ x2, c1 = maj3(x0, x1, x2);
	const uint32_t s2 = x3 ^ x4 ^ x5, c2 = maj3(x3, x4, x5);
	const uint32_t c3 = maj3(s1, s2, x6);
	return ~(c1 | c2 | c3) | (c1 & c2 & c3);
}

/* one candidate: queue entry -> true when the map says "maybe" */
__device__ __forceinline__ bool probe(uint32_t e, const uint32_t *S, const uint32_t *s_lut,
				      const uint32_t *s_map, uint32_t *lo_out, uint32_t *hi_out)
{
	const uint32_t off = e >> 5;                    /* byte offset of the window's first word */
	const uint32_t w0 = lds_off(S, off), w1 = lds_off(S, off + 4), w2 = lds_off(S, off + 8);
	const uint32_t lo = __funnelshift_r(w0, w1, e), hi = __funnelshift_r(w1, w2, e);
	const uint32_t ta = lds_off(s_lut, (hi << 2) & (uint32_t)(((1 << LUTA_BITS) - 1) << 2));
	const uint32_t tb = lds_off(s_lut + (1 << LUTA_BITS),
				    (hi >> (LUTA_BITS - 2)) & (uint32_t)(((1 << LUTB_BITS) - 1) << 2));
	const uint32_t sy = lo ^ ta ^ tb;
	const uint32_t mw = lds_off(s_map, (sy >> (32 - BLOG + 5 - 2)) & (uint32_t)((MAP_WORDS - 1) * 4));
	*lo_out = lo; *hi_out = hi;
	return (mw >> (sy & 31)) & 1;
}

/* The reference's decision for one window whose Barker tail already passed
 * (bluetooth_packet.c:387-416), exact: full 34-bit syndrome (low 32 bits from the LUTs,
 * bits 32/33 by parity), error-pattern lookup, error count, LAP. */
__device__ void exact_one(const args &a, const uint32_t *s_lut, int64_t pos, uint32_t lo, uint32_t hi)
{
	const uint32_t tail = hi >> 25;
	const int cls = __popc((tail ^ BT_BARKER_A) & 0x7f) <= 3 ? 0 : 1;
	const uint32_t ta = lds_off(s_lut, (hi << 2) & (uint32_t)(((1 << LUTA_BITS) - 1) << 2));
	const uint32_t tb = lds_off(s_lut + (1 << LUTA_BITS),
				    (hi >> (LUTA_BITS - 2)) & (uint32_t)(((1 << LUTB_BITS) - 1) << 2));
	uint64_t syn = (uint64_t)(lo ^ ta ^ tb) | ((uint64_t)(__popc(hi & a.m32) & 1) << 32) |
		       ((uint64_t)(__popc(hi & a.m33) & 1) << 33);
	syn ^= a.cc[cls];
	uint64_t sw = (((uint64_t)hi << 32) | lo) & 0x01ffffffffffffffULL;
	sw |= (uint64_t)(cls ? BT_BARKER_B : BT_BARKER_A) << 57;
	uint32_t e = 0;
	if (syn) {
		e = 0xff;
		if (a.common.err) {
			const uint64_t mask = ((uint64_t)1 << a.common.err_log2) - 1;
			uint64_t h = bt_err_hash(syn, a.common.err_log2);
			for (;;) {
				const bt_err_slot sl = a.common.err[h];
				if (sl.syn == syn) { sw ^= sl.err; e = (uint32_t)__popcll(sl.err); break; }
				if (sl.syn == 0) break;
				h = (h + 1) & mask;
			}
		}
	}
	if ((int)e <= a.common.kmax)
		push_hit(a.common, pos, (uint32_t)(sw >> 34) & 0xffffffu, e);
}

__device__ __forceinline__ void flush_exact(const args &a, const uint32_t *s_lut, uint32_t *X, int lane)
{
	__syncwarp();
	uint32_t n = X[0];
	if (n > XCAP) n = XCAP;
	if ((uint32_t)lane < n) {
		const uint32_t *x = X + 1 + 4 * lane;
		exact_one(a, s_lut, (int64_t)(((uint64_t)x[1] << 32) | x[0]), x[2], x[3]);
	}
	__syncwarp();
	if (lane == 0) X[0] = 0;
	__syncwarp();
}

__global__ void __launch_bounds__(WARPS * 32, 1) scan_promisc_v3(const args a)
{
	extern __shared__ __align__(16) uint32_t smem[];
	uint32_t *s_lut = smem;
	uint32_t *s_map = s_lut + LUT_WORDS;
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	uint32_t *S = s_map + MAP_WORDS + wid * WARP_WORDS;   /* SW + 2 bit words (+ pad) */
	uint16_t *Q = reinterpret_cast<uint16_t *>(S + S_WORDS);
	uint32_t *X = S + S_WORDS + QCAP / 2;                  /* map positives awaiting the exact test */

	for (int i = threadIdx.x; i < LUT_WORDS; i += WARPS * 32) s_lut[i] = a.lut[i];
	for (int i = threadIdx.x; i < MAP_WORDS; i += WARPS * 32) s_map[i] = a.map[i];
	if (lane == 0) X[0] = 0;
	__syncthreads();

	const int64_t gw = (int64_t)blockIdx.x * WARPS + wid, nw = (int64_t)gridDim.x * WARPS;
	const int64_t s_begin = a.nstrips * gw / nw, s_end = a.nstrips * (gw + 1) / nw;

	for (int64_t s = s_begin; s < s_end; s++) {
		/* ---- load + pack ---- */
		{
			uint32_t raw[K][8], halo[8];
			const uint8_t *p = a.base + s * STRIP + lane * 32;
			#pragma unroll
			for (int k = 0; k < K; k++) ld256(p + k * 1024, raw[k]);
			if (lane < 2) ld256(p + STRIP, halo);
			#pragma unroll
			for (int k = 0; k < K; k++) S[k * 32 + lane] = pack32(raw[k]);
			if (lane < 2) S[SW + lane] = pack32(halo);
		}
		__syncwarp();
		/* ---- filter ---- */
		uint32_t c[K];
		#pragma unroll
		for (int k = 0; k < K; k++)
			c[k] = barker_mask(S[k * 32 + lane + 1], S[k * 32 + lane + 2]);
		/* ---- packed inclusive scans of the per-row counts (two rows per register) ---- */
		uint32_t cnt01 = __popc(c[0]) | (__popc(c[1]) << 16), cnt23 = __popc(c[2]) | (__popc(c[3]) << 16);
		uint32_t inc01 = cnt01, inc23 = cnt23;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			uint32_t u = __shfl_up_sync(0xffffffffu, inc01, d), v = __shfl_up_sync(0xffffffffu, inc23, d);
			if (lane >= d) { inc01 += u; inc23 += v; }
		}
		const uint32_t tot01 = __shfl_sync(0xffffffffu, inc01, 31), tot23 = __shfl_sync(0xffffffffu, inc23, 31);
		uint32_t rowtot[K] = {tot01 & 0xffff, tot01 >> 16, tot23 & 0xffff, tot23 >> 16};
		uint32_t excl[K] = {(inc01 - cnt01) & 0xffff, (inc01 - cnt01) >> 16, (inc23 - cnt23) & 0xffff, (inc23 - cnt23) >> 16};
		const uint32_t total = rowtot[0] + rowtot[1] + rowtot[2] + rowtot[3];
		const bool one_pass = total <= QCAP;
		/* passes: the whole strip at once (normal), or row by row when the queue would overflow */
		for (int pass = 0; pass < (one_pass ? 1 : K); pass++) {
			uint32_t nq = 0;
			/* ---- compact: entry = (word byte offset << 5) | bit ---- */
			#pragma unroll
			for (int k = 0; k < K; k++) {
				if (one_pass || pass == k) {
					uint32_t m = c[k];
					uint16_t *dst = Q + nq + excl[k];
					const uint32_t ebase = (uint32_t)(k * 32 + lane) << 7;
					while (m) {
						const uint32_t q = bfind(m);
						m ^= 1u << q;
						*dst++ = (uint16_t)(ebase | q);
					}
					nq += rowtot[k];
				}
			}
			__syncwarp();
			/* ---- test: two candidates per lane per trip ---- */
			for (uint32_t i = lane; i < nq; i += 64) {
				const bool second = i + 32 < nq;
				const uint32_t e0 = Q[i], e1 = second ? Q[i + 32] : e0;
				uint32_t lo0, hi0, lo1, hi1;
				const bool m0 = probe(e0, S, s_lut, s_map, &lo0, &hi0);
				const bool m1 = probe(e1, S, s_lut, s_map, &lo1, &hi1) && second;
				if (m0 | m1) {
					for (int t = 0; t < 2; t++) {
						if (t ? m1 : m0) {
							const uint32_t e = t ? e1 : e0, lo = t ? lo1 : lo0, hi = t ? hi1 : hi0;
							const int64_t pos = a.pos0 + s * STRIP + (e >> 7) * 32 + (e & 31);
							const uint32_t slot = atomicAdd(&X[0], 1u);
							if (slot < XCAP) {
								uint32_t *x = X + 1 + 4 * slot;
								x[0] = (uint32_t)pos; x[1] = (uint32_t)(pos >> 32); x[2] = lo; x[3] = hi;
							} else
								exact_one(a, s_lut, pos, lo, hi);   /* queue full: resolve in place */
						}
					}
				}
			}
			__syncwarp();
		}
		if (X[0] >= XCAP / 2) flush_exact(a, s_lut, X, lane);
	}
	flush_exact(a, s_lut, X, lane);
}

}  // namespace v3
