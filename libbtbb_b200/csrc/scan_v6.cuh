/*
 * scan_v6.cuh -- the bulk promiscuous access-code scan, third generation.
 *
 * Same decision per window as promiscuous_packet_search (bluetooth_packet.c:368-420) and the
 * same structure as scan_v4.cuh (warp-autonomous strips, bit-sliced Barker filter, five
 * branch-free in-place candidates per word, overflow queue, two-level syndrome map).  v4 was
 * bound by the ALU pipe (LOP3 / SHF / PRMT issue at half rate), and a third of a candidate's
 * ALU work was shifting and masking table indices.  v6 removes most of that:
 *
 *   - a lane's candidate q of word L is the window that STARTS ONE SYMBOL EARLIER, at
 *     32 L + q - 1.  Its codeword bits 1..32 and 33..64 are then the two funnel shifts the
 *     lane does anyway, and the 24 received bits above the parity part, 33..56, are exactly
 *     bytes 0, 1, 2 of the second one;
 *   - the filter value is syndrome bits 1..32 instead of 0..31 (codeword bit 0 only feeds
 *     syndrome bit 0, which the map never looked at more than any other bit), so the first
 *     funnel shift is XORed in as it is;
 *   - tables A / B over bytes 1 / 2 are lane-private with a 256-byte entry pitch and share
 *     pages (A in the first 128 bytes, B in the second), so an index is ONE byte permute:
 *     prmt(hi, 4*lane) = (byte << 8) | 4*lane.  Table C over byte 0 keeps the 128-byte pitch
 *     (mask on the ALU pipe, shift-and-add as an IMAD on the FMA pipe).
 *
 * The exact test needs codeword bit 0 again: it reads that one symbol from global memory
 * (a few times per strip).
 *
 * Shared memory by absolute shared-window address: exact queues from 0x800, table C at
 * 0x4000, the second-level map at 0xC000, the first-level map at 0x10000, tables A/B at
 * 0x20000, per-warp bit tile + overflow queue from 0x30000.
 */
#pragma once

namespace v6 {

using v3::ld256;
using v3::lds32;
using v3::lds32o;
using v3::sts32;
using v3::pack32;
using v3::bfind;
using v3::xparams;
using v4::onebit;
using v4::exact_tail;

constexpr int MAXWARPS = 32;
constexpr int K = 4;
constexpr int SW = 32 * K;
constexpr int STRIP = SW * 32;
constexpr int BLOG = 19;
constexpr int MAP_WORDS = 1 << (BLOG - 5);
constexpr int M2_WORDS = 1 << 12;
constexpr int XCAP = 20;
constexpr uint32_t SA_X = 0x0800, X_BYTES = 96 * 4;
constexpr uint32_t SA_TC = 0x4000, SA_M2 = 0xC000, SA_MAP = 0x10000, SA_AB = 0x20000, SA_WARP = 0x30000;
constexpr uint32_t S_BYTES = (SW + 8) * 4;
constexpr uint32_t WARP_BYTES = S_BYTES;
constexpr size_t SMEM_BYTES = SA_WARP + MAXWARPS * WARP_BYTES;
static_assert(SA_X + MAXWARPS * X_BYTES <= SA_TC, "exact queues overlap table C");
static_assert(SMEM_BYTES <= 232448, "shared memory budget");

struct args {
	const uint8_t *base;     /* 32-byte aligned; base[0] is stream position pos0 (>= 1) */
	int64_t pos0;
	int64_t nstrips;         /* windows [pos0 - 1, pos0 - 1 + nstrips * STRIP) */
	const uint32_t *lut;     /* 768 words: tables A, B, C */
	const uint32_t *map;     /* MAP_WORDS, then M2_WORDS */
	const xparams *xp;
};

/* Barker tail (codeword bits 57..63 = bits 24..30 of the second funnel shift) within distance
 * 1 of either legal tail, for the 32 candidates of a word */
__device__ __forceinline__ uint32_t barker_mask6(uint32_t w1, uint32_t w2)
{
	const uint32_t x0 = ~__funnelshift_r(w1, w2, 24), x1 = ~__funnelshift_r(w1, w2, 25),
		       x2 = ~__funnelshift_r(w1, w2, 26), x3 = __funnelshift_r(w1, w2, 27),
		       x4 = __funnelshift_r(w1, w2, 28),  x5 = ~__funnelshift_r(w1, w2, 29),
		       x6 = __funnelshift_r(w1, w2, 30);
	const uint32_t s1 = x0 ^ x1 ^ x2, c1 = maj3(x0, x1, x2);
	const uint32_t s2 = x3 ^ x4 ^ x5, c2 = maj3(x3, x4, x5);
	const uint32_t c3 = maj3(s1, s2, x6);
	return ~(c1 | c2 | c3) | (c1 & c2 & c3);
}

__device__ __forceinline__ uint32_t ldu8(const uint8_t *p)
{
	uint32_t v;
	asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p));
	return v;
}

/* syndrome bits 1..32 of the received part: lo = codeword bits 1..32, hi = bits 33..64 */
__device__ __forceinline__ uint32_t fp32(uint32_t lo, uint32_t hi, uint32_t lane4)
{
	const uint32_t ta = lds32o<SA_AB>(__byte_perm(hi, lane4, 0x5514));
	const uint32_t tb = lds32o<SA_AB + 128>(__byte_perm(hi, lane4, 0x5524));
	uint32_t ac;
	asm("mad.lo.u32 %0, %1, 128, %2;" : "=r"(ac) : "r"(hi & 0xffu), "r"(lane4));
	const uint32_t tc = lds32o<SA_TC>(ac);
	return lo ^ ta ^ tb ^ tc;
}

__device__ __forceinline__ uint32_t map_bit(uint32_t sy)
{
	const uint32_t mw = lds32o<SA_MAP>((sy >> (32 - BLOG + 5 - 2)) & (uint32_t)((MAP_WORDS - 1) * 4));
	return (mw >> (sy & 31)) & 1;
}
__device__ __forceinline__ uint32_t map2_bit(uint32_t sy)
{
	const uint32_t mw = lds32o<SA_M2>((sy >> 8) & (uint32_t)((M2_WORDS - 1) * 4));
	return (mw >> ((sy >> 5) & 31)) & 1;
}

/* exact test of the window at stream position pos; lo / hi as in fp32() */
__device__ __forceinline__ void exact6(const xparams *xp, int64_t pos, uint32_t lo, uint32_t hi)
{
	const uint32_t lane4 = (threadIdx.x & 31) * 4;
	const uint32_t b0 = xp->stream[pos] & 1u;
	const uint32_t flo = (lo << 1) | b0, fhi = (hi << 1) | (lo >> 31);    /* codeword bits 0..31 / 32..63 */
	const uint64_t syn = ((uint64_t)fp32(lo, hi, lane4) << 1) | (uint64_t)((b0 ^ __popc(fhi & xp->m0)) & 1) |
			     ((uint64_t)(__popc(fhi & xp->m33) & 1) << 33);
	exact_tail(xp, pos, flo, fhi, syn);
}

__device__ __noinline__ void flush6(const xparams *xp, uint32_t x_sa)
{
	const int lane = threadIdx.x & 31;
	__syncwarp();
	uint32_t n = lds32(x_sa);
	if (n > XCAP) n = XCAP;
	if ((uint32_t)lane < n) {
		const uint32_t xa = x_sa + 4 + 16 * lane;
		const uint32_t p0 = lds32o<0>(xa), p1 = lds32o<4>(xa), lo = lds32o<8>(xa), hi = lds32o<12>(xa);
		exact6(xp, (int64_t)(((uint64_t)p1 << 32) | p0), lo, hi);
	}
	__syncwarp();
	if (lane == 0) sts32(x_sa, 0);
	__syncwarp();
}

/* rel = candidate index relative to the warp's run; the stream position of the run's
 * candidate 0 sits in the warp's exact-queue block (words 93/94) */
__device__ __noinline__ void park6(const xparams *xp, uint32_t x_sa, uint32_t rel, uint32_t lo, uint32_t hi)
{
	const int64_t pos = (int64_t)(((uint64_t)lds32o<94 * 4>(x_sa) << 32) | lds32o<93 * 4>(x_sa)) + rel;
	uint32_t slot;
	asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(slot) : "r"(x_sa) : "memory");
	if (slot < XCAP) {
		const uint32_t xa = x_sa + 4 + 16 * slot;
		sts32(xa, (uint32_t)pos); sts32(xa + 4, (uint32_t)(pos >> 32)); sts32(xa + 8, lo); sts32(xa + 12, hi);
	} else
		exact6(xp, pos, lo, hi);
}

/* highest remaining candidate of this lane's word, branch-free: lanes without one run the
 * same instructions on a dummy window (bfind(0) = -1, onebit() gives 0) and add nothing */
__device__ __forceinline__ void slot6(uint32_t &c, uint32_t &hitm, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t lane4)
{
	const uint32_t q = bfind(c);
	const uint32_t bit = onebit(q);
	const uint32_t lo = __funnelshift_r(w0, w1, q), hi = __funnelshift_r(w1, w2, q);
	asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hitm) : "r"(map_bit(fp32(lo, hi, lane4))), "r"(bit));
	c ^= bit;
}

/* One pending candidate of this lane (a first-level positive of the in-place slots, or a
 * candidate beyond them), taken from whichever of its four words still has one: full test
 * from the bit tile in shared memory -- the window words are no longer in registers here,
 * the next strip's loads are in flight in their place. */
__device__ __forceinline__ uint32_t pending_one(uint32_t &p0, uint32_t &p1, uint32_t &p2, uint32_t &p3,
					    uint32_t my_sa, uint32_t lane4, uint32_t lane_pos,
					    const xparams *xp, uint32_t x_sa)
{
	const uint32_t m = p0 ? p0 : p1 ? p1 : p2 ? p2 : p3;
	const uint32_t ko = p0 ? 0u : p1 ? 128u : p2 ? 256u : 384u;     /* byte offset of the word's row */
	const uint32_t q = bfind(m), bit = 1u << q;
	if (ko == 0) p0 ^= bit;
	else if (ko == 128) p1 ^= bit;
	else if (ko == 256) p2 ^= bit;
	else p3 ^= bit;
	const uint32_t wa = my_sa + ko;
	const uint32_t w0 = lds32o<0>(wa), w1 = lds32o<4>(wa), w2 = lds32o<8>(wa);
	const uint32_t lo = __funnelshift_r(w0, w1, q), hi = __funnelshift_r(w1, w2, q);
	const uint32_t sy = fp32(lo, hi, lane4);
	if (map_bit(sy) && map2_bit(sy)) {
		park6(xp, x_sa, lane_pos + ko * 8 + q, lo, hi);
		return 1;
	}
	return 0;
}

/*
 * Per strip and warp: pack the 4096 symbols loaded during the previous strip's tail into
 * the bit tile, filter, test up to NSLOTS candidates per word in place, then -- with the
 * next strip's loads already issued -- work off what is left (7 % of the candidates plus the
 * first-level map positives) in one divergent per-lane loop.
 */
template <int NSLOTS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) scan_promisc_v6(const args a)
{
	extern __shared__ __align__(16) uint32_t smem[];
	const uint32_t smem_sa = (uint32_t)__cvta_generic_to_shared(smem);
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const xparams *xp = a.xp;
	if (smem_sa > SA_X) { if (threadIdx.x == 0) atomicAdd(xp->count, 1ULL << 62); return; }  /* never: layout assumption */

	/* lane-private copies: A entry e -> SA_AB + 256 e + 4 lane, B -> the same + 128,
	 * C entry e -> SA_TC + 128 e + 4 lane */
	for (int i = threadIdx.x; i < 768 * 32; i += WARPS * 32) {
		const int e = i >> 5, l = i & 31;
		const uint32_t base = e < 256 ? SA_AB + 256 * e : e < 512 ? SA_AB + 128 + 256 * (e - 256) : SA_TC + 128 * (e - 512);
		sts32(base + 4 * l, a.lut[e]);
	}
	for (int i = threadIdx.x; i < MAP_WORDS; i += WARPS * 32) sts32(SA_MAP + 4 * i, a.map[i]);
	for (int i = threadIdx.x; i < M2_WORDS; i += WARPS * 32) sts32(SA_M2 + 4 * i, a.map[MAP_WORDS + i]);
	const uint32_t x_sa = SA_X + wid * X_BYTES;
	const uint32_t lane4 = 4 * lane, my_sa = SA_WARP + wid * WARP_BYTES + lane4;
	const int64_t nw = (int64_t)gridDim.x * WARPS;
	const int64_t gw = (int64_t)blockIdx.x * WARPS + wid;
	const int64_t s_begin = a.nstrips * gw / nw;
	const uint32_t ns = (uint32_t)(a.nstrips * (gw + 1) / nw - s_begin);
	if (lane == 0) {
		const int64_t run_pos = a.pos0 + s_begin * STRIP - 1;    /* candidate 0 of the run */
		sts32(x_sa, 0);
		sts32(x_sa + 93 * 4, (uint32_t)run_pos); sts32(x_sa + 94 * 4, (uint32_t)(run_pos >> 32));
	}
	__syncthreads();
	if (ns == 0) return;

	const uint8_t *p = a.base + s_begin * STRIP + lane * 32;
	/* the 64-symbol halo (head of the next strip) costs two registers: lane j loads symbols
	 * j and 32 + j, two ballots pack them */
	const uint8_t *hp = a.base + s_begin * STRIP + STRIP + lane;
	uint32_t raw[K][8], h0, h1;
	#pragma unroll
	for (int k = 0; k < K; k++) ld256(p + k * 1024, raw[k]);
	h0 = ldu8(hp); h1 = ldu8(hp + 32);

	uint32_t dbg_trips = 0, dbg_parks = 0, dbg_worst = 0, dbg_worst_i = 0, dbg_wait = 0, dbg_hot = 0, dbg_cold = 0;
	const long long dbg_t0 = clock64();
	for (uint32_t i = 0; i < ns; i++) {
		const long long dbg_ts = clock64();
		uint32_t wv[K];
		#pragma unroll
		for (int k = 0; k < K; k++) { wv[k] = pack32(raw[k]); sts32(my_sa + 128 * k, wv[k]); }
		{
			const uint32_t b0 = __ballot_sync(0xffffffffu, h0 & 1), b1 = __ballot_sync(0xffffffffu, h1 & 1);
			if (lane4 < 8) sts32(my_sa + 128 * K, lane4 ? b1 : b0);
		}
		p += STRIP; hp += STRIP;
		if (i + 2 < ns) {          /* pull the strip after the next one into L2 */
			#pragma unroll
			for (int k = 0; k < K; k++)
				asm volatile("prefetch.global.L2 [%0];" :: "l"(p + STRIP + k * 1024));
		}
		__syncwarp();
		dbg_wait += (uint32_t)(clock64() - dbg_ts);
		const uint32_t lane_pos = i * STRIP + lane4 * 8;          /* run-relative */
		uint32_t pend[K];
		{
			uint32_t rem[K], w1[K], w2[K], hitm[K] = {0, 0, 0, 0};
			#pragma unroll
			for (int k = 0; k < K; k++) {
				w1[k] = lds32(my_sa + 128 * k + 4); w2[k] = lds32(my_sa + 128 * k + 8);
				rem[k] = barker_mask6(w1[k], w2[k]);
			}
			#pragma unroll
			for (int t = 0; t < NSLOTS; t++) {
				#pragma unroll
				for (int k = 0; k < K; k++)
					slot6(rem[k], hitm[k], wv[k], w1[k], w2[k], lane4);
			}
			#pragma unroll
			for (int k = 0; k < K; k++) pend[k] = rem[k] | hitm[k];
		}
		const long long dbg_th = clock64();
		dbg_hot += (uint32_t)(dbg_th - dbg_ts);
		/* next strip's loads go out before the leftovers, which only need the tile */
		if (i + 1 < ns) {
			#pragma unroll
			for (int k = 0; k < K; k++) ld256(p + k * 1024, raw[k]);
			h0 = ldu8(hp); h1 = ldu8(hp + 32);
		}
		while (pend[0] | pend[1] | pend[2] | pend[3]) {
			dbg_trips++;
			dbg_parks += pending_one(pend[0], pend[1], pend[2], pend[3], my_sa, lane4, lane_pos, xp, x_sa);
		}
		__syncwarp();
		dbg_cold += (uint32_t)(clock64() - dbg_th);
		if ((i & 3) == 3 && lds32(x_sa) >= XCAP / 2) flush6(xp, x_sa);
		{
			const uint32_t dt = (uint32_t)(clock64() - dbg_ts);
			if (dt > dbg_worst) { dbg_worst = dt; dbg_worst_i = i; }
		}
	}
	flush6(xp, x_sa);
	if (xp->dbg) {
		const uint32_t mx = __reduce_max_sync(0xffffffffu, dbg_trips), sm = __reduce_add_sync(0xffffffffu, dbg_trips);
		const uint32_t pk = __reduce_add_sync(0xffffffffu, dbg_parks);
		if (lane == 0) {
			xp->dbg[8 * gw] = dbg_hot; xp->dbg[8 * gw + 1] = dbg_cold;
			xp->dbg[8 * gw + 2] = (uint32_t)(clock64() - dbg_t0); xp->dbg[8 * gw + 3] = ns;
			xp->dbg[8 * gw + 4] = pk; xp->dbg[8 * gw + 5] = dbg_worst; xp->dbg[8 * gw + 6] = dbg_worst_i; xp->dbg[8 * gw + 7] = dbg_wait;
		}
	}
}

}  // namespace v6
