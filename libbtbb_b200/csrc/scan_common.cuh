/*
 * scan_common.cuh -- what the bulk scan kernels (scan_v7.cuh, scan_known.cuh) share: strip
 * geometry, the parameter block of the exact test, load / shared-memory / pack primitives, and
 * the second half of the reference's decision for one window (bluetooth_packet.c:387-416).
 *
 * Geometry: a warp owns a contiguous run of 4096-symbol strips and never meets a block barrier;
 * lane L pulls symbols [32 (32 k + L), +32), k = 0..3, with one 256-bit load each (1 KiB contiguous
 * per warp instruction); 32 symbols become one word with 8 IDP.4A + 3 IMAD (FMA pipe only).
 * Shared memory is addressed by absolute shared-window addresses (ld.shared with the table base as
 * the instruction's immediate offset).
 */
#pragma once

namespace sc {

constexpr int WARPS = 32;            /* warps per CTA, one CTA per SM */
constexpr int K = 4;                 /* rows: words (32 positions each) per lane per strip */
constexpr int SW = 32 * K;           /* words per strip */
constexpr int STRIP = SW * 32;       /* symbols per strip */

/* Read by the exact test only; kept out of the kernel's parameter block so the cold path is a
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
	const uint32_t *map2g;   /* scan_v7.cuh, tables for 3 errors: 2^27-bit second-level map in global memory */
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

/* 1 << q, or 0 when q >= 32 (bfind of an empty mask) */
__device__ __forceinline__ uint32_t onebit(uint32_t q)
{
	uint32_t r;
	asm("bmsk.clamp.b32 %0, %1, 1;" : "=r"(r) : "r"(q));
	return r;
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


}  // namespace sc
