/*
 * scan_v7.cuh -- the bulk promiscuous access-code scan, fourth generation (the shipped kernel).
 *
 * Same decision per window as promiscuous_packet_search (bluetooth_packet.c:368-420) and the
 * same skeleton as scan_v4.cuh (warp-autonomous 4096-symbol strips, 256-bit loads, DP4A pack,
 * bit-sliced Barker filter, five branch-free in-place candidates per word, a per-warp queue
 * for the rest).  v4 ran at 80 % of the ALU pipe (LOP3 / SHF / PRMT / BMSK issue every second
 * cycle) with the FMA pipe at 12 %; v7 shortens a candidate's ALU work and moves what has a cheap
 * integer-multiply form to the FMA pipe:
 *
 *   syndrome   the filter works on syndrome bits 1..32, whose identity part is window bits 1..32;
 *              the 24 window bits above, 33..56, are then bytes 0..2 of the second window word,
 *              so with the lane's words pre-shifted by one the two window words are two funnel
 *              shifts and a table index is ONE byte permute, prmt(hi, 4 lane) = (byte << 8) |
 *              4 lane.  Tables B / C (bits 41..48 / 49..56) are lane-private with a 256-byte entry
 *              pitch and share pages.  Table A over bits 34..40 either keeps a 128-byte pitch
 *              (TA = 0: one mask, one IMAD) or is stored with every entry twice at a 256-byte
 *              pitch (TA = 1, shipped: byte 0 is bits 33..40 and bit 33 only reaches syndrome
 *              bit 33, so it too is one byte permute; the map shrinks to 32 KiB to make room);
 *   map        the first-level map is addressed by BYTE (address = the top 16 / 15 bits of the
 *              value, one LEA.HI, ld.shared.u8), the byte is replicated by an IMAD so that the
 *              shift by the raw value (mod 32) lands on bit (value & 7);
 *   enumerate  candidate bit b by FLO + BMSK, c - b and the hit mask update as IMADs (WIN = 0,
 *              shipped).  WIN = 1 keeps the mask bit-reversed (b = c & -c) and extracts the
 *              windows as (W * b) >> 32 with IMAD.HI + IMAD instead of funnel shifts: lighter on
 *              the ALU pipe but slower, because IMAD.HI issues every FOURTH cycle (tools/ibench.cu);
 *   halo       the 64 symbols after the strip are loaded with the strip (two byte loads per
 *              lane, two ballots) instead of by two lanes after the pack, which put a full
 *              memory round trip on every strip's critical path;
 *   positives  first-level positives are not followed up in place: they join the candidates
 *              beyond the fifth of a word in the per-warp queue, whose all-lanes-busy consumer
 *              runs both map levels and parks the rare survivors for the exact test.
 *
 * Template parameters: WIN (window extraction), NSLOTS (in-place candidates per word), TA (table A
 * / map layout), M2G (0: both maps in shared memory, tables for <= 2 errors; 1: second level in
 * global memory, 3 errors; 2: FIRST level in global memory, 4 / 5 errors), PACKED (input already
 * 32 symbols per word).  Shared memory by absolute shared-window address, see layout<TA>.
 */
#pragma once

namespace v7 {

using sc::ld256;
using sc::lds32;
using sc::lds32o;
using sc::lds16o;
using sc::sts32;
using sc::sts16o;
using sc::pack32;
using sc::bfind;
using sc::xparams;
using sc::onebit;
using sc::exact_tail;

constexpr int WARPS = 32;
constexpr int K = 4;
constexpr int SW = 32 * K;
constexpr int STRIP = SW * 32;
constexpr int XCAP = 20;
constexpr int QCAP = 256;                       /* queue entries (u16) per warp */
constexpr int M2_WORDS = 1 << 12;
constexpr int LUT_WORDS = 128 + 256 + 256;      /* tables A, B, C */
constexpr uint32_t SA_X = 0x0800, X_BYTES = 96 * 4;
constexpr uint32_t S_BYTES = (SW + 8) * 4;

/* TA = 0: exact queues 0x800, table A 0x4000 (16 KiB), second-level map 0x8000, candidate
 *         queues 0xC000, byte map 0x10000 (64 KiB), tables B/C 0x20000, bit tiles 0x30000;
 * TA = 1: exact queues 0x800, second-level map 0x4000, byte map 0x8000 (32 KiB), tables B/C
 *         0x10000, table A 0x20000 (64 KiB, every other 128 bytes unused), bit tiles 0x30000,
 *         candidate queues behind them. */
template <int TA> struct layout {
	static constexpr int map_log2 = TA ? 15 : 16;      /* bytes of the first-level map */
	static constexpr uint32_t sa_a = TA ? 0x20000 : 0x4000;
	static constexpr uint32_t sa_m2 = TA ? 0x4000 : 0x8000;
	static constexpr uint32_t sa_map = TA ? 0x8000 : 0x10000;
	static constexpr uint32_t sa_bc = TA ? 0x10000 : 0x20000;
	static constexpr uint32_t sa_warp = 0x30000;
	static constexpr uint32_t sa_q = TA ? sa_warp + WARPS * S_BYTES : 0xC000;
	static constexpr size_t smem_bytes = TA ? sa_q + WARPS * QCAP * 2 : sa_warp + WARPS * S_BYTES;
	static_assert(SA_X + WARPS * X_BYTES <= 0x4000, "exact queues overlap the next block");
	static_assert(smem_bytes <= 232448, "shared memory budget");
};

struct args {
	const uint8_t *base;     /* 32-byte aligned; base[0] is stream position pos0 */
	int64_t pos0;
	int64_t nstrips;
	const uint32_t *lut;     /* LUT_WORDS: tables A (codeword bits 34..40), B (41..48), C (49..56) -> syndrome bits 1..32 */
	const uint32_t *map;     /* first-level byte map (2^map_log2 bytes), then M2_WORDS */
	const xparams *xp;
	uint32_t m1;             /* 0xffffffff, opaque to the compiler: a * m1 + b is a - b on the FMA pipe */
	uint32_t c64;            /* 64, likewise: (x & 0xfe) * c64 + 4 lane stays an IMAD */
	const uint8_t *map1g;    /* M2G = 2 (tables for 4 / 5 errors): first-level map in global memory, */
	uint32_t m1g_shift;      /*   byte = value >> m1g_shift, bit = value & 7 */
};

__device__ __forceinline__ uint32_t mul_lo(uint32_t a, uint32_t b)
{
	uint32_t r;
	asm("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
	return r;
}
__device__ __forceinline__ uint32_t mad_lo(uint32_t a, uint32_t b, uint32_t c)
{
	uint32_t r;
	asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
	return r;
}
__device__ __forceinline__ uint32_t mul_hi(uint32_t a, uint32_t b)
{
	uint32_t r;
	asm("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
	return r;
}
template <uint32_t BASE>
__device__ __forceinline__ uint32_t lds8o(uint32_t sa)
{
	uint32_t v;
	asm volatile("ld.shared.u8 %0, [%1+%2];" : "=r"(v) : "r"(sa), "n"(BASE));
	return v;
}
__device__ __forceinline__ uint32_t ldg8(const uint8_t *p)
{
	uint32_t v;
	asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p));
	return v;
}
__device__ __forceinline__ uint32_t lop3_81(uint32_t a, uint32_t b, uint32_t c)     /* all three equal */
{
	uint32_t r;
	asm("lop3.b32 %0, %1, %2, %3, 0x81;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
	return r;
}

/* Barker tail within distance 1 of either legal tail (BARKER_DISTANCE <= 1, :385) for the 32
 * windows that start in the word before w1: tail bit j of window i is bit i + 25 + j of w2:w1.
 * x_j = tail bit j XOR tail A's bit j; the mismatch count is 0, 1, 6 or 7 exactly when the
 * three carries of the adder tree agree. */
__device__ __forceinline__ uint32_t barker_mask7(uint32_t w1, uint32_t w2)
{
	const uint32_t x0 = ~__funnelshift_r(w1, w2, 25), x1 = ~__funnelshift_r(w1, w2, 26),
		       x2 = ~__funnelshift_r(w1, w2, 27), x3 = __funnelshift_r(w1, w2, 28),
		       x4 = __funnelshift_r(w1, w2, 29),  x5 = ~__funnelshift_r(w1, w2, 30),
		       x6 = __funnelshift_r(w1, w2, 31);
	const uint32_t s1 = x0 ^ x1 ^ x2, c1 = maj3(x0, x1, x2);
	const uint32_t s2 = x3 ^ x4 ^ x5, c2 = maj3(x3, x4, x5);
	const uint32_t c3 = maj3(s1, s2, x6);
	return lop3_81(c1, c2, c3);
}

/* syndrome bits 1..32 of the received part (window bits 0..56): lo1 = window bits 1..32,
 * hi1 = window bits 33..64 (bit 33 and the bits above 56 are ignored) */
template <int TA>
__device__ __forceinline__ uint32_t fp7(uint32_t lo1, uint32_t hi1, uint32_t lane4, uint32_t c64)
{
	typedef layout<TA> L;
	const uint32_t tb = lds32o<L::sa_bc>(__byte_perm(hi1, lane4, 0x5514));
	const uint32_t tc = lds32o<L::sa_bc + 128>(__byte_perm(hi1, lane4, 0x5524));
	const uint32_t ta = TA ? lds32o<L::sa_a>(__byte_perm(hi1, lane4, 0x5504))
			       : lds32o<L::sa_a>(mad_lo(hi1 & 0xfeu, c64, lane4));
	return lo1 ^ ta ^ tb ^ tc;
}

/* M2G = 2: with 10^6 (4 errors) or 10^7 (5 errors) reachable values no map that fits in shared
 * memory filters anything, so the first level itself lives in global memory -- 2^27 / 2^29 bits,
 * L2-resident next to the stream, whose loads do not allocate in L1 -- and every real candidate
 * costs one 32-byte L2 sector.  live == 0 (a lane without a candidate) skips the load. */
struct gmap { const uint8_t *p; uint32_t shift; };
__device__ __forceinline__ uint32_t map1g_bit(uint32_t sy, gmap m, uint32_t live)
{
	uint32_t v = 0;
	if (live) asm volatile("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=r"(v) : "l"(m.p + (sy >> m.shift)));
	return (v >> (sy & 7)) & 1u;
}

/* first-level map: byte = the top map_log2 bits of the value, bit = (value & 7); returns 0 / 1 */
template <int TA>
__device__ __forceinline__ uint32_t map1_bit(uint32_t sy)
{
	typedef layout<TA> L;
	const uint32_t rep = mul_lo(lds8o<L::sa_map>(sy >> (32 - L::map_log2)), 0x01010101u);
	return (rep >> (sy & 31)) & 1u;
}
/* second-level map: word = bits 8..19, bit = bits 3..7.  M2G (tables for 3 errors: 68 558
 * reachable values would fill the 2^17-bit shared map to 41 %): a 2^27-bit map in global memory,
 * word = bits 10..31, bit = bits 5..9; it is L2-resident (16 MiB) and consulted about 60 times
 * per strip and warp, by all lanes of a consumer round at once */
constexpr int M2G_LOG2 = 27;
template <int TA, int M2G>
__device__ __forceinline__ uint32_t map2_bit(uint32_t sy, const xparams *xp)
{
	if (M2G == 2) return 1u;                 /* straight to the exact test (a few per strip) */
	if (M2G) {
		const uint32_t mw = __ldg(xp->map2g + (sy >> 10));
		return (mw >> ((sy >> 5) & 31)) & 1u;
	}
	const uint32_t mw = lds32o<layout<TA>::sa_m2>((sy >> 6) & (uint32_t)((M2_WORDS - 1) * 4));
	return (mw >> ((sy >> 3) & 31)) & 1u;
}

/* exact test of a window: lo / hi = window bits 0..31 / 32..63 */
template <int TA>
__device__ __forceinline__ void exact7(const xparams *xp, int64_t pos, uint32_t lo, uint32_t hi)
{
	const uint32_t lane4 = (threadIdx.x & 31) * 4;
	const uint32_t mid = fp7<TA>(__funnelshift_r(lo, hi, 1), hi >> 1, lane4, 64u);
	const uint64_t syn = ((uint64_t)mid << 1) | (uint64_t)((lo ^ __popc(hi & xp->m0)) & 1) |
			     ((uint64_t)(__popc(hi & xp->m33) & 1) << 33);
	exact_tail(xp, pos, lo, hi, syn);
}

template <int TA>
__device__ __noinline__ void flush7(const xparams *xp, uint32_t x_sa, int lane)
{
	__syncwarp();
	uint32_t n = lds32(x_sa);
	if (n > XCAP) n = XCAP;
	if ((uint32_t)lane < n) {
		const uint32_t xa = x_sa + 4 + 16 * lane;
		const uint32_t p0 = lds32o<0>(xa), p1 = lds32o<4>(xa), lo = lds32o<8>(xa), hi = lds32o<12>(xa);
		exact7<TA>(xp, (int64_t)(((uint64_t)p1 << 32) | p0), lo, hi);
	}
	__syncwarp();
	if (lane == 0) sts32(x_sa, 0);
	__syncwarp();
}

/* rel = symbol index relative to the warp's run (its stream position sits in words 93/94 of
 * the warp's exact-queue block) */
template <int TA>
__device__ __noinline__ void park7(const xparams *xp, uint32_t x_sa, uint32_t rel, uint32_t lo, uint32_t hi)
{
	const int64_t pos = (int64_t)(((uint64_t)lds32o<94 * 4>(x_sa) << 32) | lds32o<93 * 4>(x_sa)) + rel;
	uint32_t slot;
	asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(slot) : "r"(x_sa) : "memory");
	if (slot < XCAP) {
		const uint32_t xa = x_sa + 4 + 16 * slot;
		sts32(xa, (uint32_t)pos); sts32(xa + 4, (uint32_t)(pos >> 32)); sts32(xa + 8, lo); sts32(xa + 12, hi);
	} else
		exact7<TA>(xp, pos, lo, hi);
}

/* One in-place candidate, branch-free.  WIN = 1: c is the bit-reversed candidate mask and
 * r0..r2 the lane's stream words w0, w1, w2; WIN = 0: c is the plain mask and r0..r2 are
 * words 0..2 of W >> 1 (W = w0 | w1 << 32 | w2 << 64).  A lane without a candidate runs the
 * same instructions on an all-zero (WIN = 1) or dummy (WIN = 0) window and adds nothing: its
 * b is 0.  A first-level positive sets the candidate's bit in hitm (same bit order as c). */
template <int WIN, int TA, int M2G>
__device__ __forceinline__ void slot7(uint32_t &c, uint32_t &hitm, uint32_t r0, uint32_t r1, uint32_t r2,
				      uint32_t lane4, uint32_t m1, uint32_t c64, gmap gm)
{
	uint32_t b, lo1, hi1;
	if (WIN == 1) {
		b = c & mul_lo(c, m1);
		c = mad_lo(b, m1, c);
		/* (W * b) >> 32 = W >> (q + 1), word by word; the partial products never overlap */
		lo1 = mad_lo(r1, b, mul_hi(r0, b));
		hi1 = mad_lo(r2, b, mul_hi(r1, b));
	} else {
		const uint32_t q = bfind(c);
		b = onebit(q);
		c = mad_lo(b, m1, c);
		lo1 = __funnelshift_r(r0, r1, q);
		hi1 = __funnelshift_r(r1, r2, q);
	}
	const uint32_t sy = fp7<TA>(lo1, hi1, lane4, c64);
	hitm = mad_lo(M2G == 2 ? map1g_bit(sy, gm, b) : map1_bit<TA>(sy), b, hitm);
}

/* one queued candidate (or any candidate, from the bit tile): both map levels, then park */
template <int TA, int M2G>
__device__ __forceinline__ void tile_candidate(const xparams *xp, uint32_t x_sa, uint32_t wa, uint32_t q, uint32_t rel,
					       uint32_t lane4, uint32_t c64, gmap gm)
{
	const uint32_t w0 = lds32o<0>(wa), x1 = lds32o<4>(wa), x2 = lds32o<8>(wa);
	const uint32_t lo = __funnelshift_r(w0, x1, q), hi = __funnelshift_r(x1, x2, q);
	const uint32_t sy = fp7<TA>(__funnelshift_r(lo, hi, 1), hi >> 1, lane4, c64);
	if ((M2G == 2 ? map1g_bit(sy, gm, 1u) : map1_bit<TA>(sy)) && map2_bit<TA, M2G>(sy, xp))
		park7<TA>(xp, x_sa, rel, lo, hi);
}

/* PACKED: a.base points at the stream already packed 32 symbols per word, LSB first (format B of
 * SURVEY.md 8d; the host entry points pack before the PCIe copy): the load / pack stage becomes
 * one 4-byte load per lane and row */
template <int WIN, int NSLOTS, int TA, int M2G = 0, bool PACKED = false>
__global__ void __launch_bounds__(WARPS * 32, 1) scan_promisc_v7(const args a)
{
	typedef layout<TA> L;
	extern __shared__ __align__(16) uint32_t smem[];
	const uint32_t smem_sa = (uint32_t)__cvta_generic_to_shared(smem);
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const xparams *xp = a.xp;
	if (smem_sa > SA_X) { if (threadIdx.x == 0) atomicAdd(xp->count, 1ULL << 62); return; }  /* never: layout assumption */

	/* lane-private copies.  TA = 0: A entry e -> sa_a + 128 e + 4 lane; TA = 1: A entry e ->
	 * sa_a + 512 e + 4 lane and again 256 bytes further (index bit 0 is a don't-care).
	 * B entry e -> sa_bc + 256 e + 4 lane, C -> the same + 128 */
	for (int i = threadIdx.x; i < LUT_WORDS * 32; i += WARPS * 32) {
		const int e = i >> 5, l = i & 31;
		const uint32_t v = a.lut[e];
		if (e < 128) {
			if (TA) { sts32(L::sa_a + 512 * e + 4 * l, v); sts32(L::sa_a + 512 * e + 256 + 4 * l, v); }
			else sts32(L::sa_a + 128 * e + 4 * l, v);
		} else
			sts32((e < 384 ? L::sa_bc + 256 * (e - 128) : L::sa_bc + 128 + 256 * (e - 384)) + 4 * l, v);
	}
	if (M2G != 2)
		for (int i = threadIdx.x; i < (1 << (L::map_log2 - 2)); i += WARPS * 32) sts32(L::sa_map + 4 * i, a.map[i]);
	if (!M2G)
		for (int i = threadIdx.x; i < M2_WORDS; i += WARPS * 32) sts32(L::sa_m2 + 4 * i, a.map[(1 << (L::map_log2 - 2)) + i]);
	const uint32_t x_sa = SA_X + wid * X_BYTES;
	const uint32_t s_sa = L::sa_warp + wid * S_BYTES;
	const uint32_t q_sa = L::sa_q + wid * QCAP * 2;
	const uint32_t qn_sa = x_sa + 95 * 4;                   /* candidate-queue fill level */
	const int64_t gw = (int64_t)blockIdx.x * WARPS + wid, nw = (int64_t)gridDim.x * WARPS;
	const int64_t s_begin = a.nstrips * gw / nw, s_end = a.nstrips * (gw + 1) / nw;
	if (lane == 0) {
		const int64_t run_pos = a.pos0 + s_begin * STRIP;
		sts32(x_sa, 0); sts32(qn_sa, 0);
		sts32(x_sa + 93 * 4, (uint32_t)run_pos); sts32(x_sa + 94 * 4, (uint32_t)(run_pos >> 32));
	}
	__syncthreads();
	const uint32_t lane4 = 4 * lane, my_sa = s_sa + lane4;
	const uint32_t m1 = a.m1, c64 = a.c64;
	const gmap gm = {a.map1g, a.m1g_shift};

	for (int64_t s = s_begin; s < s_end; s++) {
		uint32_t wv[K];
		/* ---- load + pack ---- */
		if (PACKED) {
			const uint32_t *pw = reinterpret_cast<const uint32_t *>(a.base) + s * SW + lane;
			#pragma unroll
			for (int k = 0; k < K; k++) wv[k] = sc::ldg32(pw + 32 * k);
			uint32_t halo = 0;
			if (lane < 2) halo = sc::ldg32(pw + SW);                /* words 0 / 1 of the next strip */
			if (s + 1 < s_end && lane < K)
				asm volatile("prefetch.global.L2 [%0];" :: "l"(pw - lane + SW + 32 * lane));
			#pragma unroll
			for (int k = 0; k < K; k++) sts32(my_sa + 128 * k, wv[k]);
			if (lane < 2) sts32(my_sa + 128 * K, halo);
		} else {
			uint32_t raw[K][8];
			const uint8_t *p = a.base + s * STRIP + lane * 32;
			#pragma unroll
			for (int k = 0; k < K; k++) ld256(p + k * 1024, raw[k]);
			/* 64-symbol halo = head of the next strip: lane j takes symbols j and 32 + j */
			const uint8_t *hp = a.base + (s + 1) * STRIP + lane;
			const uint32_t h0 = ldg8(hp), h1 = ldg8(hp + 32);
			if (s + 1 < s_end) {   /* pull the next strip into L2 while this one is processed */
				#pragma unroll
				for (int k = 0; k < K; k++)
					asm volatile("prefetch.global.L2 [%0];" :: "l"(p + STRIP + k * 1024));
			}
			#pragma unroll
			for (int k = 0; k < K; k++) { wv[k] = pack32(raw[k]); sts32(my_sa + 128 * k, wv[k]); }
			const uint32_t b0 = __ballot_sync(0xffffffffu, h0 & 1), b1 = __ballot_sync(0xffffffffu, h1 & 1);
			if (lane < 2) sts32(my_sa + 128 * K, lane ? b1 : b0);
		}
		__syncwarp();
		const uint32_t strip_pos = (uint32_t)(s - s_begin) * STRIP;   /* run-relative */
		/* ---- filter all rows, then the first candidates of every word in place; the rows
		 * are independent dependency chains, so slot t of all four rows is issued together ---- */
		uint32_t rem[K], hitm[K], r0[K], r1[K], r2[K];
		#pragma unroll
		for (int k = 0; k < K; k++) {
			const uint32_t w1 = lds32(my_sa + 128 * k + 4), w2 = lds32(my_sa + 128 * k + 8);
			const uint32_t c = barker_mask7(w1, w2);
			hitm[k] = 0;
			if (WIN == 1) {
				rem[k] = __brev(c);
				r0[k] = wv[k]; r1[k] = w1; r2[k] = w2;
			} else {
				rem[k] = c;
				r0[k] = __funnelshift_r(wv[k], w1, 1); r1[k] = __funnelshift_r(w1, w2, 1); r2[k] = w2 >> 1;
			}
		}
		if (M2G == 2 && WIN == 0) {
			/* Global first-level map: a probe is an L2 round trip, so the slots run in two phases.
			 * A: window + syndrome value of every in-place candidate, parked in a lane-private
			 * column of the (otherwise unused) map area of shared memory. */
			static_assert(M2G != 2 || NSLOTS * K <= 20, "parking area holds 20 values per lane");
			uint32_t c0[K];
			#pragma unroll
			for (int k = 0; k < K; k++) c0[k] = rem[k];
			const uint32_t pend_a = L::sa_map + wid * 2048 + lane4, pend_b = L::sa_m2 + wid * 512 + lane4;
			#pragma unroll
			for (int t = 0; t < NSLOTS; t++) {
				#pragma unroll
				for (int k = 0; k < K; k++) {
					const int j = t * K + k;
					const uint32_t q = bfind(rem[k]);
					rem[k] = mad_lo(onebit(q), m1, rem[k]);
					const uint32_t sy = fp7<TA>(__funnelshift_r(r0[k], r1[k], q), __funnelshift_r(r1[k], r2[k], q), lane4, c64);
					sts32(j < 16 ? pend_a + 128 * j : pend_b + 128 * (j - 16), sy);
				}
			}
			/* B: all NSLOTS * K probes of the lane are issued back to back (one byte load each, an L2
			 * round trip), then looked at: the syndrome values come back from the parking area and the
			 * candidate bits are enumerated a second time instead of being kept in registers */
			uint32_t v[K][NSLOTS];
			{
				uint32_t c1[K];
				#pragma unroll
				for (int k = 0; k < K; k++) c1[k] = c0[k];
				#pragma unroll
				for (int t = 0; t < NSLOTS; t++) {
					#pragma unroll
					for (int k = 0; k < K; k++) {
						const int j = t * K + k;
						const uint32_t sy = lds32(j < 16 ? pend_a + 128 * j : pend_b + 128 * (j - 16));
						const uint32_t b1 = onebit(bfind(c1[k]));
						c1[k] = mad_lo(b1, m1, c1[k]);
						v[k][t] = 0;
						if (b1)
							asm volatile("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=r"(v[k][t]) : "l"(gm.p + (sy >> gm.shift)));
					}
				}
			}
			#pragma unroll
			for (int t = 0; t < NSLOTS; t++) {
				#pragma unroll
				for (int k = 0; k < K; k++) {
					const int j = t * K + k;
					const uint32_t sy = lds32(j < 16 ? pend_a + 128 * j : pend_b + 128 * (j - 16));
					const uint32_t b1 = onebit(bfind(c0[k]));
					c0[k] = mad_lo(b1, m1, c0[k]);
					const uint32_t x = (mul_lo(v[k][t], 0x01010101u) >> (sy & 31)) & 1u;
					hitm[k] = mad_lo(x, b1, hitm[k]);
				}
			}
		} else {
			#pragma unroll
			for (int t = 0; t < NSLOTS; t++) {
				#pragma unroll
				for (int k = 0; k < K; k++)
					slot7<WIN, TA, M2G>(rem[k], hitm[k], r0[k], r1[k], r2[k], lane4, m1, c64, gm);
			}
		}
		/* ---- what is left: candidates beyond the inline slots (7 %) and first-level
		 * positives (0.7 % / 1.3 %), as plain masks again ---- */
		#pragma unroll
		for (int k = 0; k < K; k++) {
			rem[k] |= hitm[k];
			if (WIN == 1) rem[k] = __brev(rem[k]);
		}
		const uint32_t mine = __popc(rem[0]) + __popc(rem[1]) + __popc(rem[2]) + __popc(rem[3]);
		if (__any_sync(0xffffffffu, mine != 0)) {
			const uint32_t ov = __reduce_add_sync(0xffffffffu, mine);
			if (ov > (uint32_t)QCAP) {
				/* more than the queue holds: dense / adversarial input only -- finish in
				 * place, one candidate per lane and trip, straight from the bit tile */
				#pragma unroll
				for (int k = 0; k < K; k++) {
					uint32_t c = rem[k];
					while (c) {
						const uint32_t q = bfind(c);
						c ^= 1u << q;
						tile_candidate<TA, M2G>(xp, x_sa, my_sa + 128 * k, q, strip_pos + (k * 32 + lane) * 32 + q, lane4, c64, gm);
					}
				}
			} else {
				/* queue + all-lanes-busy consumer; one shared atomic per lane reserves its entries */
				if (mine) {
					uint32_t at;
					asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(at) : "r"(qn_sa), "r"(mine) : "memory");
					uint32_t dst = q_sa + 2 * at;
					#pragma unroll
					for (int k = 0; k < K; k++) {
						uint32_t m = rem[k];
						const uint32_t ebase = (uint32_t)(k * 32 + lane) << 7;
						while (m) {
							const uint32_t q0 = bfind(m);
							m ^= 1u << q0;
							sts16o<0>(dst, ebase | q0);
							dst += 2;
						}
					}
				}
				__syncwarp();
				for (uint32_t i = lane; i < ov; i += 32) {
					const uint32_t e = lds16o<0>(q_sa + 2 * i);
					tile_candidate<TA, M2G>(xp, x_sa, s_sa + (e >> 5), e, strip_pos + (e >> 7) * 32 + (e & 31), lane4, c64, gm);
				}
				__syncwarp();
				if (lane == 0) sts32(qn_sa, 0);
			}
		}
		__syncwarp();
		if (lds32(x_sa) >= XCAP / 2) flush7<TA>(xp, x_sa, lane);
	}
	flush7<TA>(xp, x_sa, lane);
}

}  // namespace v7
