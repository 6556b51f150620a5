/*
 * scan_v4.cuh -- the bulk promiscuous access-code scan, second generation.
 *
 * Same decision per window as promiscuous_packet_search (bluetooth_packet.c:368-420) and the
 * same load / pack / bit-sliced Barker filter as scan_v3.cuh.  What changed is how the ~1/8
 * surviving positions are tested, because v3 was bound by shared-memory wavefronts (80 % of
 * peak at 46 % of the HBM roofline):
 *
 *   - each lane tests the first five candidates of each of its words IN PLACE, from the
 *     window words it already holds in registers (no queue write, no queue read, no
 *     re-fetch of the window): that covers ~93 % of all candidates;
 *   - the low 32 syndrome bits come from four LANE-PRIVATE tables over codeword bits
 *     32..38 / 39..44 / 45..50 / 51..56 (gen_syndrome, :147-159, regrouped): entry e of lane
 *     L lives in bank L, so every lookup is a single conflict-free wavefront;
 *   - only the candidates beyond the fifth of a word (and nothing else) go through the
 *     row-major queue + all-lanes-busy consumer of v3;
 *   - a row with more than 255 candidates (never on real captures, only on adversarial
 *     input) is simply looped in place until every lane is done.
 *
 * Shared memory, by absolute shared-window address: exact queues from 0x800, the four tables
 * at 0x4000/0x8000/0xA000/0xC000 (128+64+64+64 entries x 32 lanes x 4 B), the 2^19-bit
 * syndrome map at 0x10000, per-warp bit tile + overflow queue from 0x20000.
 */
#pragma once

namespace v4 {

using v3::ld256;
using v3::ldg32;
using v3::lds32;
using v3::lds32o;
using v3::lds16o;
using v3::sts32;
using v3::sts16o;
using v3::pack32;
using v3::bfind;
using v3::barker_mask;
using v3::xparams;

constexpr int WARPS = 32;
constexpr int K = 4;
constexpr int SW = 32 * K;
constexpr int STRIP = SW * 32;
constexpr int BLOG = 19;
constexpr int MAP_WORDS = 1 << (BLOG - 5);
constexpr int INLINE_SLOTS = 5;
constexpr int QCAP = 1024;
constexpr int XCAP = 20;
constexpr int LUT_ENTRIES = 128 + 64 + 64 + 64;    /* fields of 7, 6, 6, 6 bits */
constexpr uint32_t SA_X = 0x0800, X_BYTES = 96 * 4;
constexpr uint32_t SA_T0 = 0x4000, SA_T1 = 0x8000, SA_T2 = 0xA000, SA_T3 = 0xC000;
constexpr uint32_t SA_MAP = 0x10000, SA_WARP = 0x20000;
constexpr uint32_t S_BYTES = (SW + 8) * 4;
constexpr uint32_t WARP_BYTES = S_BYTES + QCAP * 2;
constexpr uint32_t SA_END = SA_WARP + WARPS * WARP_BYTES;
constexpr size_t SMEM_BYTES = SA_END + (1 << 12) * 4;
/* LUTMODE 2: three lane-private field tables over codeword bits 34..41 / 42..49 / 50..56
 * (256 + 256 + 128 entries x 128 B = 80 KiB): T0 at 0x4000, T2 at 0xC000, T1 behind the map
 * at 0x20000; the per-warp blocks move to 0x28000 and the overflow queue shrinks to 512. */
constexpr uint32_t SA_F0 = 0x4000, SA_F2 = 0xC000, SA_F1 = 0x20000, SA_WARP_M2 = 0x28000;
constexpr int QCAP_M2 = 512;
/* second-level map: 2^17 bits over syndrome bits 5..21, looked at only after a positive in the
 * first map (which is 0.7 % full, i.e. ~3.5 false positives per warp and strip): it removes
 * 97 % of them before they reach the exact queue */
constexpr int M2_WORDS = 1 << 12;
constexpr int LUT3_ENTRIES = 256 + 256 + 128;
template <int LUTMODE> struct layout {
	static constexpr uint32_t sa_warp = LUTMODE == 2 ? SA_WARP_M2 : SA_WARP;
	static constexpr int qcap = LUTMODE == 2 ? QCAP_M2 : QCAP;
	static constexpr uint32_t warp_bytes = S_BYTES + qcap * 2;
	static constexpr uint32_t sa_m2 = sa_warp + WARPS * warp_bytes;       /* second-level map */
	static constexpr size_t smem_bytes = sa_m2 + M2_WORDS * 4;
};

struct args {
	const uint8_t *base;     /* 32-byte aligned; base[0] is stream position pos0 */
	int64_t pos0;
	int64_t nstrips;
	const uint32_t *lut;     /* LUT_ENTRIES words: the four field tables back to back (LUTMODE 0) */
	const uint32_t *lut3;    /* LUT3_ENTRIES words: three field tables (LUTMODE 2) */
	const uint32_t *lut2;    /* LUT A (2^13 entries, codeword bits 34..46) then LUT B (2^10, bits 47..56) (LUTMODE 1) */
	const uint32_t *map;     /* MAP_WORDS, then M2_WORDS of the second-level map */
	const xparams *xp;
};

/* low 32 syndrome bits of the received part (bits 0..56): four conflict-free lookups.
 * lane4 = 4 * lane; a table entry e of this lane sits at base + 128 e + 4 lane. */
__device__ __forceinline__ uint32_t syn_lo32(uint32_t lo, uint32_t hi, uint32_t lane4)
{
	const uint32_t t0 = lds32o<SA_T0>(((hi << 7) & (127u << 7)) | lane4);
	const uint32_t t1 = lds32o<SA_T1>((hi & (63u << 7)) | lane4);
	const uint32_t t2 = lds32o<SA_T2>(((hi >> 6) & (63u << 7)) | lane4);
	const uint32_t t3 = lds32o<SA_T3>(((hi >> 12) & (63u << 7)) | lane4);
	return lo ^ t0 ^ t1 ^ t2 ^ t3;
}

/* LUTMODE 1: two shared tables over codeword bits 34..46 / 47..56 (bits 32 and 33 pass
 * straight into syndrome bits 32/33, which the filter does not look at).  Table A's 13
 * index bits sit at bits 2..14 of `hi`, i.e. already scaled to a byte offset: one LOP3.
 * Fewer instructions than the lane-private tables, but the lookups collide in the banks
 * like any random access. */
__device__ __forceinline__ uint32_t syn_lo32_2(uint32_t lo, uint32_t hi)
{
	const uint32_t ta = lds32o<0x8000>(hi & (8191u << 2));
	const uint32_t tb = lds32o<0x4000>((hi >> 13) & (1023u << 2));
	return lo ^ ta ^ tb;
}
/* 1 << q, or 0 when q >= 32 (bfind of an empty mask) */
__device__ __forceinline__ uint32_t onebit(uint32_t q)
{
	uint32_t r;
	asm("bmsk.clamp.b32 %0, %1, 1;" : "=r"(r) : "r"(q));
	return r;
}
/* LUTMODE 2: three conflict-free lookups (bits 32/33 skipped as in LUTMODE 1) */
__device__ __forceinline__ uint32_t syn_lo32_3(uint32_t lo, uint32_t hi, uint32_t lane4)
{
	const uint32_t t0 = lds32o<SA_F0>(((hi << 5) & (255u << 7)) | lane4);
	const uint32_t t1 = lds32o<SA_F1>(((hi >> 3) & (255u << 7)) | lane4);
	const uint32_t t2 = lds32o<SA_F2>(((hi >> 11) & (127u << 7)) | lane4);
	return lo ^ t0 ^ t1 ^ t2;
}
template <int LUTMODE>
__device__ __forceinline__ uint32_t syn32(uint32_t lo, uint32_t hi, uint32_t lane4)
{
	return LUTMODE == 1 ? syn_lo32_2(lo, hi) : LUTMODE == 2 ? syn_lo32_3(lo, hi, lane4) : syn_lo32(lo, hi, lane4);
}

__device__ __forceinline__ uint32_t map_bit(uint32_t sy)
{
	const uint32_t mw = lds32o<SA_MAP>((sy >> (32 - BLOG + 5 - 2)) & (uint32_t)((MAP_WORDS - 1) * 4));
	return (mw >> (sy & 31)) & 1;
}

template <uint32_t SA_M2>
__device__ __forceinline__ uint32_t map2_bit(uint32_t sy)
{
	const uint32_t mw = lds32o<SA_M2>((sy >> 8) & (uint32_t)((M2_WORDS - 1) * 4));
	return (mw >> ((sy >> 5) & 31)) & 1;
}

/* Second half of the reference's decision (bluetooth_packet.c:387-416) for a window whose
 * 34-bit syndrome of the received part is known: fold in the tail constant, look the error
 * pattern up, count, extract the LAP, emit. */
__device__ __noinline__ void exact_tail(const xparams *xp, int64_t pos, uint32_t lo, uint32_t hi, uint64_t syn)
{
	const uint32_t tail = hi >> 25;
	const int cls = __popc((tail ^ BT_BARKER_A) & 0x7f) <= 3 ? 0 : 1;
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
	if (xp->slab_cnt) {          /* slab mode: this warp's own slab (see find_ac.cu) */
		const uint32_t gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), cap = xp->slab_cap;
		const uint32_t i = atomicAdd(&xp->slab_cnt[gw], 1u);
		if (i < cap) {
			btbb_b200_hit h;
			h.offset = pos + xp->bias; h.lap = lap; h.ac_errors = (uint8_t)e; h.pad[0] = h.pad[1] = h.pad[2] = 0;
			xp->slab[(size_t)gw * cap + i] = h;
		}
		return;
	}
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

/* exact test: the low 32 syndrome bits come from the same tables as the filter's */
template <int LUTMODE>
__device__ __forceinline__ void exact4(const xparams *xp, int64_t pos, uint32_t lo, uint32_t hi)
{
	const uint32_t lane4 = (threadIdx.x & 31) * 4;
	const uint64_t syn = (uint64_t)syn32<LUTMODE>(lo, hi, lane4) | ((uint64_t)(__popc(hi & xp->m32) & 1) << 32) |
			     ((uint64_t)(__popc(hi & xp->m33) & 1) << 33);
	exact_tail(xp, pos, lo, hi, syn);
}

template <int LUTMODE>
__device__ __noinline__ void flush4(const xparams *xp, uint32_t x_sa, int lane)
{
	__syncwarp();
	uint32_t n = lds32(x_sa);
	if (n > XCAP) n = XCAP;
	if ((uint32_t)lane < n) {
		const uint32_t xa = x_sa + 4 + 16 * lane;
		const uint32_t p0 = lds32o<0>(xa), p1 = lds32o<4>(xa), lo = lds32o<8>(xa), hi = lds32o<12>(xa);
		exact4<LUTMODE>(xp, (int64_t)(((uint64_t)p1 << 32) | p0), lo, hi);
	}
	__syncwarp();
	if (lane == 0) sts32(x_sa, 0);
	__syncwarp();
}

/* rel = symbol index relative to the warp's run; the run's stream position sits in the
 * warp's exact-queue block (words 93/94) so the hot loops carry 32-bit positions only */
template <int LUTMODE>
__device__ __noinline__ void park4(const xparams *xp, uint32_t x_sa, uint32_t rel, uint32_t lo, uint32_t hi)
{
	const int64_t pos = (int64_t)(((uint64_t)lds32o<94 * 4>(x_sa) << 32) | lds32o<93 * 4>(x_sa)) + rel;
	uint32_t slot;
	asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(slot) : "r"(x_sa) : "memory");
	if (slot < XCAP) {
		const uint32_t xa = x_sa + 4 + 16 * slot;
		sts32(xa, (uint32_t)pos); sts32(xa + 4, (uint32_t)(pos >> 32)); sts32(xa + 8, lo); sts32(xa + 12, hi);
	} else
		exact4<LUTMODE>(xp, pos, lo, hi);
}

/* take the highest remaining candidate of this lane's word (if any) and test it in place.
 * BF (branch-free): lanes without a candidate run the same instructions on a dummy window
 * (bfind(0) = -1, onebit() then gives 0) and contribute nothing; a map positive only sets
 * the candidate's bit in `hitm` (one IMAD), the rare follow-up happens once per strip. */
template <int LUTMODE, bool BF>
__device__ __forceinline__ void slot(uint32_t &c, uint32_t &hitm, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t lane4,
				     const xparams *xp, uint32_t x_sa, uint32_t word_pos)
{
	if (BF) {
		const uint32_t q = bfind(c);
		const uint32_t bit = onebit(q);
		const uint32_t lo = __funnelshift_r(w0, w1, q), hi = __funnelshift_r(w1, w2, q);
		/* hitm += mapbit * bit as one IMAD (FMA pipe) instead of compare + select + add */
		asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hitm) : "r"(map_bit(syn32<LUTMODE>(lo, hi, lane4))), "r"(bit));
		c ^= bit;
	} else if (c) {
		const uint32_t q = bfind(c);
		c ^= 1u << q;
		const uint32_t lo = __funnelshift_r(w0, w1, q), hi = __funnelshift_r(w1, w2, q);
		const uint32_t sy = syn32<LUTMODE>(lo, hi, lane4);
		if (map_bit(sy) && map2_bit<layout<LUTMODE>::sa_m2>(sy))
			park4<LUTMODE>(xp, x_sa, word_pos + q, lo, hi);
	}
}

/* the map positives a row collected in its inline slots */
template <int LUTMODE>
__device__ __forceinline__ void park_row(uint32_t hitm, uint32_t w0, uint32_t w1, uint32_t w2,
					 const xparams *xp, uint32_t x_sa, uint32_t word_pos)
{
	while (hitm) {
		const uint32_t q = bfind(hitm);
		hitm ^= 1u << q;
		const uint32_t lo = __funnelshift_r(w0, w1, q), hi = __funnelshift_r(w1, w2, q);
		if (map2_bit<layout<LUTMODE>::sa_m2>(syn32<LUTMODE>(lo, hi, (threadIdx.x & 31) * 4)))
			park4<LUTMODE>(xp, x_sa, word_pos + q, lo, hi);
	}
}

/* PACKED: a.base points at the stream already packed 32 symbols per word, LSB first (format B
 * of SURVEY.md 8d; the host entry points pack before the PCIe copy): the load / pack stage
 * becomes one 4-byte load per lane and row */
template <int LUTMODE, int NSLOTS, bool BF, bool PACKED = false>
__global__ void __launch_bounds__(WARPS * 32, 1) scan_promisc_v4(const args a)
{
	extern __shared__ __align__(16) uint32_t smem[];
	const uint32_t smem_sa = (uint32_t)__cvta_generic_to_shared(smem);
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const xparams *xp = a.xp;
	if (smem_sa > SA_X) { if (threadIdx.x == 0) atomicAdd(xp->count, 1ULL << 62); return; }  /* never: layout assumption */

	if (LUTMODE == 0) {
		/* lane-private copies of the four field tables: entry e -> base + 128 e + 4 lane */
		for (int i = threadIdx.x; i < LUT_ENTRIES * 32; i += WARPS * 32) {
			const int e = i >> 5, l = i & 31;
			const uint32_t base = e < 128 ? SA_T0 + 128 * e : e < 192 ? SA_T1 + 128 * (e - 128)
					    : e < 256 ? SA_T2 + 128 * (e - 192) : SA_T3 + 128 * (e - 256);
			sts32(base + 4 * l, a.lut[e]);
		}
	} else if (LUTMODE == 2) {
		for (int i = threadIdx.x; i < LUT3_ENTRIES * 32; i += WARPS * 32) {
			const int e = i >> 5, l = i & 31;
			const uint32_t base = e < 256 ? SA_F0 + 128 * e : e < 512 ? SA_F1 + 128 * (e - 256) : SA_F2 + 128 * (e - 512);
			sts32(base + 4 * l, a.lut3[e]);
		}
	} else {
		for (int i = threadIdx.x; i < 8192; i += WARPS * 32) sts32(0x8000 + 4 * i, a.lut2[i]);
		for (int i = threadIdx.x; i < 1024; i += WARPS * 32) sts32(0x4000 + 4 * i, a.lut2[8192 + i]);
	}
	for (int i = threadIdx.x; i < MAP_WORDS; i += WARPS * 32) sts32(SA_MAP + 4 * i, a.map[i]);
	for (int i = threadIdx.x; i < M2_WORDS; i += WARPS * 32) sts32(layout<LUTMODE>::sa_m2 + 4 * i, a.map[MAP_WORDS + i]);
	const uint32_t x_sa = SA_X + wid * X_BYTES;
	const uint32_t s_sa = layout<LUTMODE>::sa_warp + wid * layout<LUTMODE>::warp_bytes;
	const uint32_t q_sa = s_sa + S_BYTES;
	const uint32_t qn_sa = x_sa + 95 * 4;                   /* overflow-queue fill level */
	const int64_t gw = (int64_t)blockIdx.x * WARPS + wid, nw = (int64_t)gridDim.x * WARPS;
	const int64_t s_begin = a.nstrips * gw / nw, s_end = a.nstrips * (gw + 1) / nw;
	if (lane == 0) {
		const int64_t run_pos = a.pos0 + s_begin * STRIP;
		sts32(x_sa, 0); sts32(qn_sa, 0);
		sts32(x_sa + 93 * 4, (uint32_t)run_pos); sts32(x_sa + 94 * 4, (uint32_t)(run_pos >> 32));
	}
	__syncthreads();
	const uint32_t lane4 = 4 * lane, my_sa = s_sa + lane4;

	for (int64_t s = s_begin; s < s_end; s++) {
		uint32_t wv[K];
		/* ---- load + pack ---- */
		if (PACKED) {
			const uint32_t *pw = reinterpret_cast<const uint32_t *>(a.base) + s * SW + lane;
			#pragma unroll
			for (int k = 0; k < K; k++) wv[k] = ldg32(pw + 32 * k);
			#pragma unroll
			for (int k = 0; k < K; k++) sts32(my_sa + 128 * k, wv[k]);
			if (lane < 2) sts32(my_sa + 128 * K, ldg32(pw + SW));
			if (s + 1 < s_end && lane < K)
				asm volatile("prefetch.global.L2 [%0];" :: "l"(pw - lane + SW + 32 * lane));
		} else {
			uint32_t raw[K][8];
			const uint8_t *p = a.base + s * STRIP + lane * 32;
			#pragma unroll
			for (int k = 0; k < K; k++) ld256(p + k * 1024, raw[k]);
			#pragma unroll
			for (int k = 0; k < K; k++) { wv[k] = pack32(raw[k]); sts32(my_sa + 128 * k, wv[k]); }
			if (lane < 2) {        /* 64-symbol halo = first two words of the next strip (L2-resident) */
				ld256(p + STRIP, raw[0]);
				sts32(my_sa + 128 * K, pack32(raw[0]));
			}
			if (s + 1 < s_end) {   /* pull the next strip into L2 while this one is processed */
				#pragma unroll
				for (int k = 0; k < K; k++)
					asm volatile("prefetch.global.L2 [%0];" :: "l"(p + STRIP + k * 1024));
			}
		}
		__syncwarp();
		const uint32_t strip_pos = (uint32_t)(s - s_begin) * STRIP;   /* run-relative */
		const uint32_t lane_pos = strip_pos + lane * 32;
		/* ---- filter all rows, then the first candidates of every word in place; the rows
		 * are independent dependency chains, so slot t of all four rows is issued together ---- */
		uint32_t rem[K], w1[K], w2[K];
		#pragma unroll
		for (int k = 0; k < K; k++) {
			w1[k] = lds32(my_sa + 128 * k + 4); w2[k] = lds32(my_sa + 128 * k + 8);
			rem[k] = barker_mask(w1[k], w2[k]);
		}
		{
			uint32_t hitm[K] = {0, 0, 0, 0};
			#pragma unroll
			for (int t = 0; t < NSLOTS; t++) {
				#pragma unroll
				for (int k = 0; k < K; k++)
					slot<LUTMODE, BF>(rem[k], hitm[k], wv[k], w1[k], w2[k], lane4, xp, x_sa, lane_pos + k * 1024);
			}
			if (BF && (hitm[0] | hitm[1] | hitm[2] | hitm[3])) {
				#pragma unroll
				for (int k = 0; k < K; k++)
					park_row<LUTMODE>(hitm[k], wv[k], w1[k], w2[k], xp, x_sa, lane_pos + k * 1024);
			}
		}
		/* ---- candidates beyond the inline slots (7 %) ---- */
		const uint32_t mine = __popc(rem[0]) + __popc(rem[1]) + __popc(rem[2]) + __popc(rem[3]);
		if (__any_sync(0xffffffffu, mine != 0)) {
			const uint32_t ov = __reduce_add_sync(0xffffffffu, mine);
			if (ov > (uint32_t)layout<LUTMODE>::qcap) {
				/* more than the queue holds: adversarial input only -- finish in place */
				#pragma unroll
				for (int k = 0; k < K; k++) {
					uint32_t c = rem[k], dummy = 0;
					while (__any_sync(0xffffffffu, c != 0))
						slot<LUTMODE, false>(c, dummy, wv[k], w1[k], w2[k], lane4, xp, x_sa, lane_pos + k * 1024);
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
					const uint32_t wa = s_sa + (e >> 5);
					const uint32_t w0 = lds32o<0>(wa), x1 = lds32o<4>(wa), x2 = lds32o<8>(wa);
					const uint32_t lo = __funnelshift_r(w0, x1, e), hi = __funnelshift_r(x1, x2, e);
					const uint32_t sy = syn32<LUTMODE>(lo, hi, lane4);
					if (map_bit(sy) && map2_bit<layout<LUTMODE>::sa_m2>(sy))
						park4<LUTMODE>(xp, x_sa, strip_pos + (e >> 7) * 32 + (e & 31), lo, hi);
				}
				__syncwarp();
				if (lane == 0) sts32(qn_sa, 0);
			}
		}
		__syncwarp();
		if (lds32(x_sa) >= XCAP / 2) flush4<LUTMODE>(xp, x_sa, lane);
	}
	flush4<LUTMODE>(xp, x_sa, lane);
}

}  // namespace v4
