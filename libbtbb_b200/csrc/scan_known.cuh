/*
 * scan_known.cuh -- bulk kernel for the known-LAP access-code search (find_known_lap,
 * bluetooth_packet.c:423-441: first/every position with popcount(window ^ syncword) <= k).
 *
 * Same skeleton as scan_v4.cuh (warp-autonomous 4096-symbol strips, 256-bit loads, DP4A pack)
 * but the per-position work is almost entirely bit-sliced, which makes this mode HBM-bound:
 *
 *   prefilter  a window can only match if it already has <= k mismatches among any 16 of its
 *              64 bits.  The host picks 16 positions inside the first 32 sync-word bits where
 *              the sync word is all 0 (or all 1 -- one of the two always has 16), so the
 *              mismatch vectors are the shifted stream words themselves (or their complement,
 *              folded into the LOP3 tables): 16 funnel shifts, a carry-save adder tree and a
 *              bit-sliced "count <= k" compare for 32 positions at once.  Random data passes
 *              with probability 0.2 % at k = 2.
 *   exact      survivors get the reference's test, popcount of the full 64-bit XOR.
 */
#pragma once

namespace vk {

using sc::ld256;
using sc::ldg32;
using sc::lds32;
using sc::lds32o;
using sc::sts32;
using sc::pack32;
using sc::bfind;
using sc::xparams;

constexpr int WARPS = 32;
constexpr int K = 4;
constexpr int STRIP = 4096;
constexpr uint32_t S_BYTES = (32 * K + 8) * 4;
constexpr size_t SMEM_BYTES = 0x800 + WARPS * S_BYTES;

struct args {
	const uint8_t *base;     /* 32-byte aligned; base[0] is stream position pos0 */
	int64_t pos0;
	int64_t nstrips;
	uint64_t ac;             /* sync word of the LAP (btbb_gen_syncword, :188-199) */
	uint32_t lap;
	int kk;                  /* min(max_ac_errors, 16), or -1 when nothing can match */
	int kmax;                /* max_ac_errors */
	uint32_t sh[16];         /* prefilter bit positions (0..31) */
	uint32_t sh2[16];        /* second group (k >= 3): positions 32..63 minus 32 */
	const xparams *xp;
};

/* append one hit (slab mode, first-hit mode or the plain unordered list; see find_ac.cu) */
__device__ __noinline__ void emit_hit(const xparams *xp, int64_t pos, uint32_t lap, uint32_t e)
{
	if (xp->slab_cnt) {
		const uint32_t gw = blockIdx.x * WARPS + (threadIdx.x >> 5), cap = xp->slab_cap;
		const uint32_t i = atomicAdd(&xp->slab_cnt[gw], 1u);
		if (i < cap) {
			btbb_b200_hit h;
			h.offset = pos + xp->bias; h.lap = lap; h.ac_errors = (uint8_t)e; h.pad[0] = h.pad[1] = h.pad[2] = 0;
			xp->slab[(size_t)gw * cap + i] = h;
		}
		return;
	}
	const int64_t max_hits = xp->max_hits;
	if (max_hits < 0) {
		atomicMin(xp->count, ((unsigned long long)(pos + xp->bias) << 32) | ((unsigned long long)lap << 8) | (e & 0xff));
		return;
	}
	const unsigned long long slot = atomicAdd(xp->count, 1ULL);
	if ((int64_t)slot < max_hits) {
		btbb_b200_hit h;
		h.offset = pos + xp->bias; h.lap = lap; h.ac_errors = (uint8_t)e; h.pad[0] = h.pad[1] = h.pad[2] = 0;
		xp->hits[slot] = h;
	}
}

/* 5 count planes of the 16 mismatch vectors at positions sh[] of the pair (lo, hi) */
template <bool INV>
__device__ __forceinline__ void count16(uint32_t lo, uint32_t hi, const uint32_t sh[16], uint32_t cnt[5])
{
	uint32_t x[16];
	#pragma unroll
	for (int j = 0; j < 16; j++) {
		const uint32_t s = __funnelshift_r(lo, hi, sh[j]);
		x[j] = INV ? ~s : s;
	}
	csa16(x, cnt);
}

/* positions with count <= kk, count given as NP bit planes */
template <int NP>
__device__ __forceinline__ uint32_t le_const(const uint32_t *cnt, int kk)
{
	uint32_t less = 0, eq = 0xffffffffu;
	#pragma unroll
	for (int i = NP - 1; i >= 0; i--) {
		const uint32_t kb = ((kk >> i) & 1) ? 0xffffffffu : 0u;
		less |= eq & ~cnt[i] & kb;
		eq &= ~(cnt[i] ^ kb);
	}
	return less | eq;
}

/* TWO: a second group of 16 positions in the upper sync-word half (used for k >= 3, where 16
 * positions alone would let through 1 .. 10 % of random windows) */
template <bool INV, bool TWO, bool INV2, bool PACKED = false>
__global__ void __launch_bounds__(WARPS * 32, 1) scan_known_v4(const args a)
{
	extern __shared__ __align__(16) uint32_t smem[];
	const uint32_t smem_sa = (uint32_t)__cvta_generic_to_shared(smem);
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t s_sa = ((smem_sa + 127) & ~127u) + wid * S_BYTES;
	const uint32_t my_sa = s_sa + 4 * lane;
	const int64_t gw = (int64_t)blockIdx.x * WARPS + wid, nw = (int64_t)gridDim.x * WARPS;
	const int64_t s_begin = a.nstrips * gw / nw, s_end = a.nstrips * (gw + 1) / nw;
	const uint32_t ac_lo = (uint32_t)a.ac, ac_hi = (uint32_t)(a.ac >> 32);
	uint32_t sh[16], sh2[16];
	#pragma unroll
	for (int j = 0; j < 16; j++) { sh[j] = a.sh[j]; sh2[j] = a.sh2[j]; }
	if (a.kk < 0) return;

	for (int64_t s = s_begin; s < s_end; s++) {
		uint32_t wv[K];
		if (PACKED) {      /* stream already packed 32 symbols per word (see scan_v4.cuh) */
			const uint32_t *pw = reinterpret_cast<const uint32_t *>(a.base) + s * (32 * K) + lane;
			#pragma unroll
			for (int k = 0; k < K; k++) wv[k] = ldg32(pw + 32 * k);
			#pragma unroll
			for (int k = 0; k < K; k++) sts32(my_sa + 128 * k, wv[k]);
			if (lane < 2) sts32(my_sa + 128 * K, ldg32(pw + 32 * K));
			if (s + 1 < s_end && lane < K)
				asm volatile("prefetch.global.L2 [%0];" :: "l"(pw - lane + 32 * K + 32 * lane));
		} else {
			uint32_t raw[K][8];
			const uint8_t *p = a.base + s * STRIP + lane * 32;
			#pragma unroll
			for (int k = 0; k < K; k++) ld256(p + k * 1024, raw[k]);
			#pragma unroll
			for (int k = 0; k < K; k++) { wv[k] = pack32(raw[k]); sts32(my_sa + 128 * k, wv[k]); }
			if (lane < 2) {
				ld256(p + STRIP, raw[0]);
				sts32(my_sa + 128 * K, pack32(raw[0]));
			}
			if (s + 1 < s_end) {
				#pragma unroll
				for (int k = 0; k < K; k++)
					asm volatile("prefetch.global.L2 [%0];" :: "l"(p + STRIP + k * 1024));
			}
		}
		__syncwarp();
		#pragma unroll
		for (int k = 0; k < K; k++) {
			const uint32_t w1 = lds32(my_sa + 128 * k + 4), w2 = lds32(my_sa + 128 * k + 8);
			uint32_t c;
			if (TWO) {
				uint32_t ca[5], cb[5], sum[6], carry = 0;
				count16<INV>(wv[k], w1, sh, ca);
				count16<INV2>(w1, w2, sh2, cb);
				#pragma unroll
				for (int i = 0; i < 5; i++) {        /* ripple add of the two 5-plane counts */
					sum[i] = ca[i] ^ cb[i] ^ carry;
					carry = maj3(ca[i], cb[i], carry);
				}
				sum[5] = carry;
				c = le_const<6>(sum, a.kk);
			} else {
				uint32_t ca[5];
				count16<INV>(wv[k], w1, sh, ca);
				c = le_const<5>(ca, a.kk);
			}
			if (c) {
				do {
					const uint32_t q = bfind(c);
					c ^= 1u << q;
					const uint32_t lo = __funnelshift_r(wv[k], w1, q), hi = __funnelshift_r(w1, w2, q);
					const int d = __popc(lo ^ ac_lo) + __popc(hi ^ ac_hi);
					if (d <= a.kmax)
						emit_hit(a.xp, a.pos0 + s * STRIP + (k * 32 + lane) * 32 + q, a.lap, (uint32_t)(uint8_t)d);
				} while (c);
			}
		}
		__syncwarp();
	}
}

}  // namespace vk
