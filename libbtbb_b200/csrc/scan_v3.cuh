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
 *   exact    the few map positives go through exact_one() (the reference's test: full
 *            34-bit syndrome, error lookup, error count, LAP) from a small second queue.
 *
 * Shared memory is addressed by absolute shared-window addresses (ld.shared with the table
 * base as the instruction's immediate offset, so a lookup is shift + mask + LDS): LUT B at
 * 0x4000 (16 KiB), LUT A at 0x8000 (32 KiB), the map at 0x10000 (64 KiB), the per-warp exact
 * queues below 0x4000, the per-warp bit tiles and candidate queues from 0x20000 up.
 */
#pragma once

namespace v3 {

constexpr int WARPS = 32;            /* warps per CTA, one CTA per SM */
constexpr int K = 4;                 /* rows: words (32 positions each) per lane per strip */
constexpr int SW = 32 * K;           /* words per strip */
constexpr int STRIP = SW * 32;       /* symbols per strip */
constexpr int BLOG = 19;             /* log2 bits of the syndrome map */
constexpr int QCAP = 1024;           /* queue entries (u16); a row never holds more */
constexpr int XCAP = 20;             /* map-positive queue per warp (flushed when half full) */
constexpr int LUTA_BITS = 13, LUTB_BITS = 12;
constexpr int LUT_WORDS = (1 << LUTA_BITS) + (1 << LUTB_BITS);
constexpr int MAP_WORDS = 1 << (BLOG - 5);
/* absolute shared addresses */
constexpr uint32_t SA_X = 0x0800;                 /* 32 warps x 96 words */
constexpr uint32_t X_BYTES = 96 * 4;
constexpr uint32_t SA_LUTB = 0x4000, SA_LUTA = 0x8000, SA_MAP = 0x10000, SA_WARP = 0x20000;
constexpr uint32_t S_BYTES = (SW + 8) * 4;
constexpr uint32_t WARP_BYTES = S_BYTES + QCAP * 2;
constexpr uint32_t SA_END = SA_WARP + WARPS * WARP_BYTES;
constexpr size_t SMEM_BYTES = SA_END;             /* dynamic request; the window starts at <= 0x800 */
static_assert(SA_X + WARPS * X_BYTES <= SA_LUTB, "exact queues overlap LUT B");

struct xparams;
struct args {
	const uint8_t *base;     /* 32-byte aligned; base[0] is stream position pos0 */
	int64_t pos0;
	int64_t nstrips;         /* strips [0, nstrips); base[nstrips*STRIP + 63] is readable */
	const uint32_t *lut;     /* LUT_WORDS: A then B */
	const uint32_t *map;     /* MAP_WORDS */
	const struct xparams *xp; /* everything only the exact test needs (device memory) */
};

/* Read by exact_one() only; kept out of the kernel's parameter block so the cold path is a
 * plain function taking one pointer and the hot loops keep their registers. */
struct xparams {
	uint64_t cc[2];          /* 34-bit syndrome of PN ^ (legal tail << 57), tail A / tail B */
	uint32_t m32, m33;       /* codeword bits 32..56 (as bits of `hi`) feeding syndrome bits 32 / 33 */
	int kmax, err_log2;
	const bt_err_slot *err;
	btbb_b200_hit *hits;
	int64_t max_hits;
	unsigned long long *count;
	int64_t bias;
	/* slab mode (find_ac_dev): every warp appends to its own slab, so that one sort per
	 * slab + concatenation in warp order gives the ascending list without a global sort */
	btbb_b200_hit *slab;
	uint32_t *slab_cnt;
	uint32_t slab_cap;
	uint32_t m0;             /* like m32: codeword bits 32..56 feeding syndrome bit 0 (scan_v7.cuh) */
	const uint32_t *map2g;   /* scan_v7.cuh with tables for 3 errors: 2^27-bit second-level map in global memory */
};

__device__ __forceinline__ void ld256(const uint8_t *p, uint32_t r[8])
{
	asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
		     : "l"(p));
}
__device__ __forceinline__ uint32_t ldg32(const uint32_t *p)
{
	uint32_t v;
	asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
	return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t sa)
{
	uint32_t v;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(sa));
	return v;
}
template <int OFF>
__device__ __forceinline__ uint32_t lds32o(uint32_t sa)
{
	uint32_t v;
	asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(sa), "n"(OFF));
	return v;
}
template <int OFF>
__device__ __forceinline__ uint32_t lds16o(uint32_t sa)
{
	uint32_t v;
	asm volatile("ld.shared.u16 %0, [%1+%2];" : "=r"(v) : "r"(sa), "n"(OFF));
	return v;
}
__device__ __forceinline__ void sts32(uint32_t sa, uint32_t v)
{
	asm volatile("st.shared.u32 [%0], %1;" :: "r"(sa), "r"(v) : "memory");
}
template <int OFF>
__device__ __forceinline__ void sts16o(uint32_t sa, uint32_t v)
{
	asm volatile("st.shared.u16 [%0+%1], %2;" :: "r"(sa), "n"(OFF), "r"(v) : "memory");
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

/* low 32 syndrome bits of the received part (bits 0..56) of a window */
__device__ __forceinline__ uint32_t syn_lo32(uint32_t lo, uint32_t hi)
{
	const uint32_t ta = lds32o<SA_LUTA>((hi << 2) & (uint32_t)(((1 << LUTA_BITS) - 1) << 2));
	const uint32_t tb = lds32o<SA_LUTB>((hi >> (LUTA_BITS - 2)) & (uint32_t)(((1 << LUTB_BITS) - 1) << 2));
	return lo ^ ta ^ tb;
}

/* one candidate: queue entry -> map word shifted so that bit 0 says "maybe" */
__device__ __forceinline__ uint32_t probe(uint32_t e, uint32_t s_sa, uint32_t *lo_out, uint32_t *hi_out)
{
	const uint32_t wa = s_sa + (e >> 5);            /* address of the window's first word */
	const uint32_t w0 = lds32o<0>(wa), w1 = lds32o<4>(wa), w2 = lds32o<8>(wa);
	const uint32_t lo = __funnelshift_r(w0, w1, e), hi = __funnelshift_r(w1, w2, e);
	const uint32_t sy = syn_lo32(lo, hi);
	const uint32_t mw = lds32o<SA_MAP>((sy >> (32 - BLOG + 5 - 2)) & (uint32_t)((MAP_WORDS - 1) * 4));
	*lo_out = lo; *hi_out = hi;
	return mw >> (sy & 31);
}

/* The reference's decision for one window whose Barker tail already passed
 * (bluetooth_packet.c:387-416), exact: full 34-bit syndrome (low 32 bits from the LUTs,
 * bits 32/33 by parity), error-pattern lookup, error count, LAP. */
__device__ __noinline__ void exact_one(const xparams *xp, int64_t pos, uint32_t lo, uint32_t hi)
{
	const uint32_t tail = hi >> 25;
	const int cls = __popc((tail ^ BT_BARKER_A) & 0x7f) <= 3 ? 0 : 1;
	uint64_t syn = (uint64_t)syn_lo32(lo, hi) | ((uint64_t)(__popc(hi & xp->m32) & 1) << 32) |
		       ((uint64_t)(__popc(hi & xp->m33) & 1) << 33);
	syn ^= xp->cc[cls];
	uint64_t sw = (((uint64_t)hi << 32) | lo) & 0x01ffffffffffffffULL;
	sw |= (uint64_t)(cls ? BT_BARKER_B : BT_BARKER_A) << 57;
	uint32_t e = 0;
	if (syn) {
		e = 0xff;
		const bt_err_slot *tab = xp->err;
		if (tab) {
			const int lg = xp->err_log2;
			const uint64_t mask = ((uint64_t)1 << lg) - 1;
			uint64_t h = bt_err_hash(syn, lg);
			for (;;) {
				const bt_err_slot sl = tab[h];
				if (sl.syn == syn) { sw ^= sl.err; e = (uint32_t)__popcll(sl.err); break; }
				if (sl.syn == 0) break;
				h = (h + 1) & mask;
			}
		}
	}
	if ((int)e > xp->kmax) return;
	const uint32_t lap = (uint32_t)(sw >> 34) & 0xffffffu;
	const int64_t max_hits = xp->max_hits;
	if (max_hits < 0) {          /* first-hit mode, see push_hit() */
		atomicMin(xp->count, ((unsigned long long)(pos + xp->bias) << 32) | ((unsigned long long)lap << 8) | e);
		return;
	}
	const unsigned long long slot = atomicAdd(xp->count, 1ULL);
	if ((int64_t)slot < max_hits) {
		btbb_b200_hit h;
		h.offset = pos + xp->bias; h.lap = lap; h.ac_errors = (uint8_t)e; h.pad[0] = h.pad[1] = h.pad[2] = 0;
		xp->hits[slot] = h;
	}
}

__device__ __noinline__ void flush_exact(const xparams *xp, uint32_t x_sa, int lane)
{
	__syncwarp();
	uint32_t n = lds32(x_sa);
	if (n > XCAP) n = XCAP;
	if ((uint32_t)lane < n) {
		const uint32_t xa = x_sa + 4 + 16 * lane;
		const uint32_t p0 = lds32o<0>(xa), p1 = lds32o<4>(xa), lo = lds32o<8>(xa), hi = lds32o<12>(xa);
		exact_one(xp, (int64_t)(((uint64_t)p1 << 32) | p0), lo, hi);
	}
	__syncwarp();
	if (lane == 0) sts32(x_sa, 0);
	__syncwarp();
}

/* a map positive: park it for the exact test (or resolve in place when the queue is full) */
__device__ __forceinline__ void park(const xparams *xp, uint32_t x_sa, int64_t pos, uint32_t lo, uint32_t hi)
{
	uint32_t slot;
	asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(slot) : "r"(x_sa) : "memory");
	if (slot < XCAP) {
		const uint32_t xa = x_sa + 4 + 16 * slot;
		sts32(xa, (uint32_t)pos); sts32(xa + 4, (uint32_t)(pos >> 32)); sts32(xa + 8, lo); sts32(xa + 12, hi);
	} else
		exact_one(xp, pos, lo, hi);
}

__global__ void __launch_bounds__(WARPS * 32, 1) scan_promisc_v3(const args a)
{
	extern __shared__ __align__(16) uint32_t smem[];
	const uint32_t smem_sa = (uint32_t)__cvta_generic_to_shared(smem);
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const xparams *xp = a.xp;
	if (smem_sa > SA_X) { if (threadIdx.x == 0) atomicAdd(xp->count, 1ULL << 62); return; }  /* never: layout assumption */

	for (int i = threadIdx.x; i < (1 << LUTA_BITS); i += WARPS * 32) sts32(SA_LUTA + 4 * i, a.lut[i]);
	for (int i = threadIdx.x; i < (1 << LUTB_BITS); i += WARPS * 32) sts32(SA_LUTB + 4 * i, a.lut[(1 << LUTA_BITS) + i]);
	for (int i = threadIdx.x; i < MAP_WORDS; i += WARPS * 32) sts32(SA_MAP + 4 * i, a.map[i]);
	const uint32_t x_sa = SA_X + wid * X_BYTES;
	const uint32_t s_sa = SA_WARP + wid * WARP_BYTES;       /* SW + 2 bit words (+ pad) */
	const uint32_t q_sa = s_sa + S_BYTES;                   /* QCAP u16 entries */
	if (lane == 0) sts32(x_sa, 0);
	__syncthreads();

	const int64_t gw = (int64_t)blockIdx.x * WARPS + wid, nw = (int64_t)gridDim.x * WARPS;
	const int64_t s_begin = a.nstrips * gw / nw, s_end = a.nstrips * (gw + 1) / nw;
	const uint32_t my_sa = s_sa + 4 * lane;

	for (int64_t s = s_begin; s < s_end; s++) {
		/* ---- load + pack ---- */
		{
			uint32_t raw[K][8];
			const uint8_t *p = a.base + s * STRIP + lane * 32;
			#pragma unroll
			for (int k = 0; k < K; k++) ld256(p + k * 1024, raw[k]);
			#pragma unroll
			for (int k = 0; k < K; k++) sts32(my_sa + 128 * k, pack32(raw[k]));
			/* 64-symbol halo: the first two words of the next strip (in L2 already: it was
			 * prefetched while the previous strip was processed) */
			if (lane < 2) {
				ld256(p + STRIP, raw[0]);
				sts32(my_sa + 128 * K, pack32(raw[0]));
			}
			/* pull the next strip into L2 while this one is processed */
			if (s + 1 < s_end) {
				#pragma unroll
				for (int k = 0; k < K; k++)
					asm volatile("prefetch.global.L2 [%0];" :: "l"(p + STRIP + k * 1024));
			}
		}
		__syncwarp();
		/* ---- filter ---- */
		uint32_t c[K];
		c[0] = barker_mask(lds32o<4>(my_sa), lds32o<8>(my_sa));
		c[1] = barker_mask(lds32o<128 + 4>(my_sa), lds32o<128 + 8>(my_sa));
		c[2] = barker_mask(lds32o<256 + 4>(my_sa), lds32o<256 + 8>(my_sa));
		c[3] = barker_mask(lds32o<384 + 4>(my_sa), lds32o<384 + 8>(my_sa));
		/* ---- packed inclusive scans of the per-row counts (two rows per register) ---- */
		const uint32_t cnt01 = __popc(c[0]) | (__popc(c[1]) << 16), cnt23 = __popc(c[2]) | (__popc(c[3]) << 16);
		uint32_t inc01 = cnt01, inc23 = cnt23;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t u = __shfl_up_sync(0xffffffffu, inc01, d), v = __shfl_up_sync(0xffffffffu, inc23, d);
			if (lane >= d) { inc01 += u; inc23 += v; }
		}
		const uint32_t tot01 = __shfl_sync(0xffffffffu, inc01, 31), tot23 = __shfl_sync(0xffffffffu, inc23, 31);
		const uint32_t rowtot[K] = {tot01 & 0xffff, tot01 >> 16, tot23 & 0xffff, tot23 >> 16};
		const uint32_t excl[K] = {(inc01 - cnt01) & 0xffff, (inc01 - cnt01) >> 16,
					  (inc23 - cnt23) & 0xffff, (inc23 - cnt23) >> 16};
		const uint32_t total = rowtot[0] + rowtot[1] + rowtot[2] + rowtot[3];
		const bool one_pass = total <= QCAP;
		/* passes: the whole strip at once (normal), or row by row when the queue would overflow */
		for (int pass = 0; pass < (one_pass ? 1 : K); pass++) {
			uint32_t nq = 0;
			/* ---- compact: entry = (word byte offset << 5) | bit, two per trip ---- */
			#pragma unroll
			for (int k = 0; k < K; k++) {
				if (one_pass || pass == k) {
					uint32_t m = c[k];
					uint32_t dst = q_sa + 2 * (nq + excl[k]);
					const uint32_t ebase = (uint32_t)(k * 32 + lane) << 7;
					while (m) {
						const uint32_t q0 = bfind(m);
						m ^= 1u << q0;
						sts16o<0>(dst, ebase | q0);
						if (m) {
							const uint32_t q1 = bfind(m);
							m ^= 1u << q1;
							sts16o<2>(dst, ebase | q1);
						}
						dst += 4;
					}
					nq += rowtot[k];
				}
			}
			__syncwarp();
			/* ---- test: two candidates per lane per trip, no predication in the body ---- */
			const uint32_t full = nq & ~63u;
			uint32_t qa = q_sa + 2 * lane;
			for (uint32_t i = 0; i < full; i += 64, qa += 128) {
				const uint32_t e0 = lds16o<0>(qa), e1 = lds16o<64>(qa);
				uint32_t lo0, hi0, lo1, hi1;
				const uint32_t b0 = probe(e0, s_sa, &lo0, &hi0);
				const uint32_t b1 = probe(e1, s_sa, &lo1, &hi1);
				if ((b0 | b1) & 1) {
					const int64_t sp = a.pos0 + s * STRIP;
					if (b0 & 1) park(xp, x_sa, sp + (e0 >> 7) * 32 + (e0 & 31), lo0, hi0);
					if (b1 & 1) park(xp, x_sa, sp + (e1 >> 7) * 32 + (e1 & 31), lo1, hi1);
				}
			}
			for (uint32_t i = full + lane; i < nq; i += 32) {      /* ragged tail */
				const uint32_t e0 = lds16o<0>(q_sa + 2 * i);
				uint32_t lo0, hi0;
				if (probe(e0, s_sa, &lo0, &hi0) & 1)
					park(xp, x_sa, a.pos0 + s * STRIP + (e0 >> 7) * 32 + (e0 & 31), lo0, hi0);
			}
			__syncwarp();
		}
		if (lds32(x_sa) >= XCAP / 2) flush_exact(xp, x_sa, lane);
	}
	flush_exact(xp, x_sa, lane);
}

}  // namespace v3
