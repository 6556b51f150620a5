/*
 * find_ac.cu -- K1/K2: the sliding 64-bit access-code correlator on sm_100a.
 *
 * Replaces btbb_find_ac (bluetooth_packet.c:444-464) and its two search loops,
 * find_known_lap (:423-441) and promiscuous_packet_search (:368-420), for a whole
 * byte-per-symbol stream at once.  The reference tests one window per iteration with a
 * 64-step byte->bit pack (air_to_host64, :235-242); here
 *
 *   phase 1  each thread pulls 2x16 symbols with coalesced 128-bit loads and squeezes
 *            them to bits (one IMAD per 4 symbols) into a shared-memory bit tile,
 *   phase 2  each lane owns 32 consecutive window positions and evaluates the cheap
 *            part of the test for all 32 at once with bit-sliced logic on funnel-shifted
 *            words (promiscuous: the 7-bit Barker tail within distance 1 of either legal
 *            tail, BARKER_DISTANCE :55-59; known LAP: at most k mismatches among the
 *            first 16 sync-word bits),
 *   phase 3  the surviving ~1/8 (promiscuous) positions get the 32 low syndrome bits from
 *            shared-memory byte LUTs (gen_syndrome :147-159) and a two-hash Bloom probe of
 *            the syndrome->error map (find_syndrome :139-145),
 *   phase 4  the handful of Bloom positives run the exact reference test (full 34-bit
 *            syndrome, error lookup, error count, LAP extraction :390-416).
 *
 * Hits are appended unordered and then radix-sorted by offset, which yields exactly the
 * ascending list the reference produces when iterated with restart at offset+1.
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "bt_math.h"
#include "scan_hash.h"
#include "capi_internal.h"

namespace {

constexpr int QCAP = 512;   /* Bloom-positive queue entries per tile */

struct scan_args {
	const uint8_t *abase;   /* stream pointer rounded down to 16 bytes */
	int head;               /* abase[head] is stream[0] (0..15) */
	int64_t n;              /* search_length */
	int64_t bias;           /* added to every reported offset (chunked host scans) */
	int64_t vlen;           /* head + n + 63: virtual symbols readable */
	int64_t ntiles;
	int kmax;               /* max_ac_errors */
	uint64_t ac;            /* known-LAP sync word */
	uint32_t lap;
	const bt_scan_tables *tables;
	const uint32_t *bloom;
	int bloom_log2;
	const bt_err_slot *err;
	int err_log2;
	btbb_b200_hit *hits;
	int64_t max_hits;
	unsigned long long *count;
};

/* 4 symbols (bytes 0/1) -> 4 bits.  x*0x10204080 drops byte j's bit at 28+j; no two
 * partial products share a bit position, so no carries. */
__device__ __forceinline__ uint32_t pack4(uint32_t x) { return (x * 0x10204080u) >> 28; }

__device__ __forceinline__ uint32_t load_pack16(const scan_args &a, int64_t byte0)
{
	if (byte0 >= a.head && byte0 + 16 <= a.vlen) {
		uint4 v = __ldg(reinterpret_cast<const uint4 *>(a.abase + byte0));
		return pack4(v.x) | (pack4(v.y) << 4) | (pack4(v.z) << 8) | (pack4(v.w) << 12);
	}
	uint32_t r = 0;
	for (int j = 0; j < 16; j++) {
		int64_t i = byte0 + j;
		if (i >= a.head && i < a.vlen)
			r |= (uint32_t)(a.abase[i] & 1u) << j;
	}
	return r;
}

__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) { return (a & b) | (c & (a | b)); }

__device__ __forceinline__ void push_hit(const scan_args &a, int64_t off, uint32_t lap, uint32_t nerr)
{
	if (a.max_hits < 0) {
		/* first-hit mode (classic btbb_find_ac): keep the smallest offset; the record rides
		 * in the low bits of the key so one 64-bit atomicMin carries everything */
		atomicMin(a.count, ((unsigned long long)(off + a.bias) << 32) | ((unsigned long long)lap << 8) | (nerr & 0xff));
		return;
	}
	unsigned long long slot = atomicAdd(a.count, 1ULL);
	if ((int64_t)slot < a.max_hits) {
		btbb_b200_hit h;
		h.offset = off + a.bias; h.lap = lap; h.ac_errors = (uint8_t)nerr; h.pad[0] = h.pad[1] = h.pad[2] = 0;
		a.hits[slot] = h;
	}
}

/* The reference's per-window decision once the Barker tail is known to pass
 * (bluetooth_packet.c:387-416), exact. */
__device__ bool exact_promisc(const scan_args &a, uint64_t w, uint32_t *lap, uint32_t *nerr)
{
	uint32_t tail = (uint32_t)(w >> 57);
	uint64_t fixed = (uint64_t)(__popc((tail ^ BT_BARKER_A) & 0x7f) <= 3 ? BT_BARKER_A : BT_BARKER_B) << 57;
	uint64_t sw = (w & 0x01ffffffffffffffULL) | fixed;
	uint64_t syn = bt_syndrome_slow(sw ^ BT_PN);
	uint32_t e = 0;
	if (syn) {
		e = 0xff;
		if (a.err) {
			uint64_t mask = ((uint64_t)1 << a.err_log2) - 1, h = bt_err_hash(syn, a.err_log2);
			for (;;) {
				bt_err_slot s = a.err[h];
				if (s.syn == syn) { sw ^= s.err; e = (uint32_t)__popcll(s.err); break; }
				if (s.syn == 0) break;
				h = (h + 1) & mask;
			}
		}
	}
	if ((int)e > a.kmax) return false;
	*lap = (uint32_t)(sw >> 34) & 0xffffffu;
	*nerr = e;
	return true;
}

template <int NT>
__global__ void __launch_bounds__(NT) scan_promisc_kernel(const scan_args a)
{
	constexpr int TILE = NT * 32;
	extern __shared__ __align__(16) uint32_t smem[];
	uint32_t *sbits = smem;                         /* NT + 2 words (+2 pad) */
	uint32_t *s_ta = sbits + NT + 4;                /* 256 */
	uint32_t *s_tb = s_ta + 256, *s_tc = s_tb + 256;
	uint32_t *s_misc = s_tc + 256;                  /* t_56, c_class[2], pad */
	uint32_t *s_q = s_misc + 4;                     /* QCAP */
	uint32_t *s_qn = s_q + QCAP;                    /* 2 counters (tile parity) */
	uint32_t *s_bloom = s_qn + 4;                   /* 1 << (bloom_log2 - 5) */
	const int tid = threadIdx.x;
	const int blog = a.bloom_log2;

	for (int i = tid; i < 256; i += NT) { s_ta[i] = a.tables->t_a[i]; s_tb[i] = a.tables->t_b[i]; s_tc[i] = a.tables->t_c[i]; }
	if (tid == 0) { s_misc[0] = a.tables->t_56; s_misc[1] = a.tables->c_class[0]; s_misc[2] = a.tables->c_class[1]; s_qn[0] = 0; s_qn[1] = 0; }
	for (int i = tid; i < (1 << (blog - 5)); i += NT) s_bloom[i] = a.bloom[i];
	__syncthreads();
	const uint32_t t56 = s_misc[0], cls0 = s_misc[1], cls1 = s_misc[2];
	uint16_t *sb16 = reinterpret_cast<uint16_t *>(sbits);

	int par = 0;
	for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, par ^= 1) {
		const int64_t base = tile * TILE;
		/* ---- phase 1: symbols -> bit tile ---- */
		#pragma unroll
		for (int j = 0; j < 2; j++) {
			int c = j * NT + tid;
			sb16[c] = (uint16_t)load_pack16(a, base + 16 * (int64_t)c);
		}
		if (tid < 4)
			sb16[2 * NT + tid] = (uint16_t)load_pack16(a, base + 16 * (int64_t)(2 * NT + tid));
		if (tid == 0) s_qn[par ^ 1] = 0;
		__syncthreads();

		/* ---- phase 2: Barker tail filter for 32 positions per lane ---- */
		const uint32_t w0 = sbits[tid], w1 = sbits[tid + 1], w2 = sbits[tid + 2];
		/* S_j bit i = symbol (pos_i + 57 + j); mismatch against tail A = 0b0100111 */
		const uint32_t x0 = ~__funnelshift_r(w1, w2, 25), x1 = ~__funnelshift_r(w1, w2, 26),
			       x2 = ~__funnelshift_r(w1, w2, 27), x3 = __funnelshift_r(w1, w2, 28),
			       x4 = __funnelshift_r(w1, w2, 29),  x5 = ~__funnelshift_r(w1, w2, 30),
			       x6 = __funnelshift_r(w1, w2, 31);
		const uint32_t s1 = x0 ^ x1 ^ x2, c1 = maj3(x0, x1, x2);
		const uint32_t s2 = x3 ^ x4 ^ x5, c2 = maj3(x3, x4, x5);
		const uint32_t c3 = maj3(s1, s2, x6);
		const uint32_t near_a = ~(c1 | c2 | c3);        /* <= 1 mismatch with tail A */
		const uint32_t near_b = c1 & c2 & c3;           /* >= 6 mismatches = <= 1 with tail B */
		uint32_t cand = near_a | near_b;
		/* positions outside [head, head + n) are not searched */
		{
			const int64_t p0 = base + 32 * (int64_t)tid;
			const int64_t lo = a.head - p0, hi = a.head + a.n - p0;
			if (lo > 0) cand &= lo >= 32 ? 0u : (0xffffffffu << lo);
			if (hi < 32) cand &= hi <= 0 ? 0u : (0xffffffffu >> (32 - hi));
		}
		/* ---- phase 3: low-32 syndrome + Bloom probe per surviving position ---- */
		while (cand) {
			const int q = __ffs(cand) - 1;
			cand &= cand - 1;
			const uint32_t lo = __funnelshift_r(w0, w1, q);
			const uint32_t hi = __funnelshift_r(w1, w2, q);
			uint32_t s = lo ^ s_ta[hi & 255] ^ s_tb[(hi >> 8) & 255] ^ s_tc[(hi >> 16) & 255];
			s ^= ((hi >> 24) & 1) ? t56 : 0u;
			s ^= ((near_b >> q) & 1) ? cls1 : cls0;
			const uint32_t h1 = bt_bloom_h1(s, blog);
			if ((s_bloom[h1 >> 5] >> (h1 & 31)) & 1) {
				const uint32_t h2 = bt_bloom_h2(s, blog);
				if ((s_bloom[h2 >> 5] >> (h2 & 31)) & 1) {
					const uint32_t rel = 32 * tid + q;
					const uint32_t slot = atomicAdd(&s_qn[par], 1u);
					if (slot < QCAP)
						s_q[slot] = rel;
					else {  /* queue full (adversarial input): resolve in place */
						uint64_t w = ((uint64_t)hi << 32) | lo;
						uint32_t lap, ne;
						if (exact_promisc(a, w, &lap, &ne))
							push_hit(a, base + rel - a.head, lap, ne);
					}
				}
			}
		}
		__syncthreads();
		/* ---- phase 4: exact test of the Bloom positives ---- */
		uint32_t nq = s_qn[par];
		if (nq > QCAP) nq = QCAP;
		for (uint32_t i = tid; i < nq; i += NT) {
			const uint32_t rel = s_q[i], wi = rel >> 5, sh = rel & 31;
			const uint32_t lo = __funnelshift_r(sbits[wi], sbits[wi + 1], sh);
			const uint32_t hi = __funnelshift_r(sbits[wi + 1], sbits[wi + 2], sh);
			uint32_t lap, ne;
			if (exact_promisc(a, ((uint64_t)hi << 32) | lo, &lap, &ne))
				push_hit(a, base + rel - a.head, lap, ne);
		}
		__syncthreads();
	}
}

/* count of set inputs among 16 bit-sliced vectors -> 5 bit planes (carry-save adders) */
__device__ __forceinline__ void csa16(const uint32_t x[16], uint32_t cnt[5])
{
	#define FA(s, c, p, q, r) do { uint32_t p_ = (p), q_ = (q), r_ = (r); s = p_ ^ q_ ^ r_; c = maj3(p_, q_, r_); } while (0)
	#define HA(s, c, p, q) do { uint32_t p_ = (p), q_ = (q); s = p_ ^ q_; c = p_ & q_; } while (0)
	uint32_t a0, a1, a2, a3, a4, b0, b1, b2, b3, b4;   /* weight-1 sums a*, weight-2 carries b* */
	FA(a0, b0, x[0], x[1], x[2]);  FA(a1, b1, x[3], x[4], x[5]);  FA(a2, b2, x[6], x[7], x[8]);
	FA(a3, b3, x[9], x[10], x[11]); FA(a4, b4, x[12], x[13], x[14]);
	uint32_t d0, d1, e0, e1;
	FA(d0, e0, a0, a1, a2); FA(d1, e1, a3, a4, x[15]);
	uint32_t f0;
	HA(cnt[0], f0, d0, d1);
	/* weight 2: b0..b4, e0, e1, f0 */
	uint32_t g0, g1, h0, h1;
	FA(g0, h0, b0, b1, b2); FA(g1, h1, b3, b4, e0);
	uint32_t g2, h2;
	FA(g2, h2, g0, g1, e1);
	uint32_t h3;
	HA(cnt[1], h3, g2, f0);
	/* weight 4: h0, h1, h2, h3 */
	uint32_t m0, n0, n1;
	FA(m0, n0, h0, h1, h2);
	HA(cnt[2], n1, m0, h3);
	/* weight 8: n0, n1 */
	HA(cnt[3], cnt[4], n0, n1);
	#undef FA
	#undef HA
}

template <int NT>
__global__ void __launch_bounds__(NT) scan_known_kernel(const scan_args a)
{
	constexpr int TILE = NT * 32;
	__shared__ __align__(16) uint32_t sbits[NT + 4];
	const int tid = threadIdx.x;
	uint16_t *sb16 = reinterpret_cast<uint16_t *>(sbits);
	const uint32_t ac_lo = (uint32_t)a.ac, ac_hi = (uint32_t)(a.ac >> 32);
	const int kk = a.kmax > 16 ? 16 : (a.kmax < 0 ? -1 : a.kmax);

	for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
		const int64_t base = tile * TILE;
		#pragma unroll
		for (int j = 0; j < 2; j++) {
			int c = j * NT + tid;
			sb16[c] = (uint16_t)load_pack16(a, base + 16 * (int64_t)c);
		}
		if (tid < 4)
			sb16[2 * NT + tid] = (uint16_t)load_pack16(a, base + 16 * (int64_t)(2 * NT + tid));
		__syncthreads();
		const uint32_t w0 = sbits[tid], w1 = sbits[tid + 1], w2 = sbits[tid + 2];
		/* prefilter: mismatches among sync-word bits 0..15 must already be <= k */
		uint32_t x[16], cnt[5];
		#pragma unroll
		for (int j = 0; j < 16; j++)
			x[j] = __funnelshift_r(w0, w1, j) ^ (((ac_lo >> j) & 1) ? 0xffffffffu : 0u);
		csa16(x, cnt);
		uint32_t less = 0, eq = 0xffffffffu;
		#pragma unroll
		for (int i = 4; i >= 0; i--) {
			const uint32_t kb = (kk >= 0 && ((kk >> i) & 1)) ? 0xffffffffu : 0u;
			less |= eq & ~cnt[i] & kb;
			eq &= ~(cnt[i] ^ kb);
		}
		uint32_t cand = kk < 0 ? 0u : (less | eq);
		{
			const int64_t p0 = base + 32 * (int64_t)tid;
			const int64_t lo = a.head - p0, hi = a.head + a.n - p0;
			if (lo > 0) cand &= lo >= 32 ? 0u : (0xffffffffu << lo);
			if (hi < 32) cand &= hi <= 0 ? 0u : (0xffffffffu >> (32 - hi));
		}
		while (cand) {
			const int q = __ffs(cand) - 1;
			cand &= cand - 1;
			const uint32_t lo = __funnelshift_r(w0, w1, q);
			const uint32_t hi = __funnelshift_r(w1, w2, q);
			const int d = __popc(lo ^ ac_lo) + __popc(hi ^ ac_hi);
			if (d <= a.kmax)
				push_hit(a, base + 32 * (int64_t)tid + q - a.head, a.lap, (uint32_t)(uint8_t)d);
		}
		__syncthreads();
	}
}

/* ---------------- ordering pass: LSD radix sort of 16-byte hit records by offset (9-bit digits) ---------------- */
constexpr int SORT_CHUNK = 512;    /* records per warp */
constexpr int SORT_BITS = 9;       /* digit width: offsets below 2^36 sort in 4 passes */
constexpr int SORT_BINS = 1 << SORT_BITS;

__global__ void sort_hist_kernel(const btbb_b200_hit *in, int64_t n, int shift, uint32_t *hist, int nblk, int64_t key_bias)
{
	__shared__ uint32_t h[SORT_BINS];
	for (int i = threadIdx.x; i < SORT_BINS; i += 32) h[i] = 0;
	__syncwarp();
	int64_t b0 = (int64_t)blockIdx.x * SORT_CHUNK, b1 = b0 + SORT_CHUNK < n ? b0 + SORT_CHUNK : n;
	for (int64_t i = b0 + threadIdx.x; i < b1; i += 32)
		atomicAdd(&h[(uint32_t)((uint64_t)(in[i].offset - key_bias) >> shift) & (SORT_BINS - 1)], 1u);
	__syncwarp();
	for (int i = threadIdx.x; i < SORT_BINS; i += 32) hist[(int64_t)i * nblk + blockIdx.x] = h[i];
}

/* hist is [digit][block]: one warp per digit turns its row into an exclusive scan over the
 * blocks (coalesced) and leaves the digit's total in tot[digit] */
__global__ void __launch_bounds__(256) sort_rowscan_kernel(uint32_t *hist, int nblk, uint32_t *tot)
{
	const int d = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
	if (d >= SORT_BINS) return;
	uint32_t *row = hist + (int64_t)d * nblk;
	uint32_t carry = 0;
	for (int b0 = 0; b0 < nblk; b0 += 32) {
		const int b = b0 + lane;
		const uint32_t x = b < nblk ? row[b] : 0;
		uint32_t inc = x;
		#pragma unroll
		for (int s = 1; s < 32; s <<= 1) {
			const uint32_t u = __shfl_up_sync(0xffffffffu, inc, s);
			if (lane >= s) inc += u;
		}
		if (b < nblk) row[b] = carry + inc - x;
		carry += __shfl_sync(0xffffffffu, inc, 31);
	}
	if (lane == 0) tot[d] = carry;
}

__global__ void sort_scatter_kernel(const btbb_b200_hit *in, btbb_b200_hit *out, int64_t n, int shift,
				    const uint32_t *hist, int nblk, const uint32_t *tot, int64_t key_bias)
{
	__shared__ uint32_t cur[SORT_BINS];
	const int lane = threadIdx.x;
	/* start of digit d = sum of the totals of all smaller digits: scan the 512 totals here
	 * (16 per lane) instead of launching one more kernel */
	{
		constexpr int PER = SORT_BINS / 32;
		uint32_t loc[PER], sum = 0;
		#pragma unroll
		for (int i = 0; i < PER; i++) { loc[i] = tot[lane * PER + i]; sum += loc[i]; }
		uint32_t inc = sum;
		#pragma unroll
		for (int s = 1; s < 32; s <<= 1) {
			const uint32_t u = __shfl_up_sync(0xffffffffu, inc, s);
			if (lane >= s) inc += u;
		}
		uint32_t run = inc - sum;
		#pragma unroll
		for (int i = 0; i < PER; i++) {
			const int d = lane * PER + i;
			cur[d] = run + hist[(int64_t)d * nblk + blockIdx.x];
			run += loc[i];
		}
	}
	__syncwarp();
	int64_t b0 = (int64_t)blockIdx.x * SORT_CHUNK, b1 = b0 + SORT_CHUNK < n ? b0 + SORT_CHUNK : n;
	for (int64_t i0 = b0; i0 < b1; i0 += 32) {
		int64_t i = i0 + lane;
		bool live = i < b1;
		btbb_b200_hit rec;
		uint32_t dig = 0;
		if (live) { rec = in[i]; dig = (uint32_t)((uint64_t)(rec.offset - key_bias) >> shift) & (SORT_BINS - 1); }
		unsigned act = __ballot_sync(0xffffffffu, live);
		if (live) {
			unsigned peers = __match_any_sync(act, dig);
			unsigned rank = __popc(peers & ((1u << lane) - 1));
			uint32_t basepos = cur[dig];
			out[basepos + rank] = rec;
			__syncwarp(act);
			if (rank == 0) cur[dig] = basepos + __popc(peers);
		}
		__syncwarp();
	}
}

#include "scan_common.cuh"
#include "scan_v7.cuh"
#include "scan_known.cuh"

}  // namespace

static int scan_launch_v1(btbb_b200_ctx *ctx, const uint8_t *d_stream, int64_t n, uint32_t lap, int k,
		   btbb_b200_hit *d_out, int64_t max_hits, unsigned long long *d_count,
		   int64_t bias, cudaStream_t st)
{
	scan_args a;
	uintptr_t p = reinterpret_cast<uintptr_t>(d_stream);
	a.head = (int)(p & 15);
	a.abase = reinterpret_cast<const uint8_t *>(p - a.head);
	a.n = n;
	a.bias = bias;
	a.vlen = a.head + n + 63;
	a.kmax = k;
	a.lap = lap;
	a.ac = lap == BTBB_B200_LAP_ANY ? 0 : bt_gen_syncword(lap);
	a.tables = ctx->d_tables;
	a.bloom = ctx->d_bloom;
	a.bloom_log2 = ctx->bloom_log2;
	a.err = ctx->d_err;
	a.err_log2 = ctx->err_log2;
	a.hits = d_out;
	a.max_hits = max_hits;
	a.count = d_count;
	if (n <= 0) return BTBB_B200_OK;
	if (lap == BTBB_B200_LAP_ANY) {
		size_t bloom_bytes = (size_t)4 << (ctx->bloom_log2 - 5);
		if (ctx->bloom_log2 <= 17) {
			constexpr int NT = 256;
			size_t smem = (NT + 4 + 768 + 4 + QCAP + 4) * 4 + bloom_bytes;
			a.ntiles = (a.head + n + NT * 32 - 1) / (NT * 32);
			int cta_per_sm = 0;
			BT_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cta_per_sm, scan_promisc_kernel<NT>, NT, smem));
			if (cta_per_sm < 1) cta_per_sm = 1;
			int64_t grid = (int64_t)ctx->sm_count * cta_per_sm;
			if (grid > a.ntiles) grid = a.ntiles;
			scan_promisc_kernel<NT><<<(unsigned)grid, NT, smem, st>>>(a);
		} else {
			constexpr int NT = 1024;
			size_t smem = (NT + 4 + 768 + 4 + QCAP + 4) * 4 + bloom_bytes;
			a.ntiles = (a.head + n + NT * 32 - 1) / (NT * 32);
			BT_CUDA_TRY(cudaFuncSetAttribute(scan_promisc_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
			int64_t grid = ctx->sm_count;
			if (grid > a.ntiles) grid = a.ntiles;
			scan_promisc_kernel<NT><<<(unsigned)grid, NT, smem, st>>>(a);
		}
	} else {
		constexpr int NT = 256;
		a.ntiles = (a.head + n + NT * 32 - 1) / (NT * 32);
		int64_t grid = (int64_t)ctx->sm_count * 8;
		if (grid > a.ntiles) grid = a.ntiles;
		scan_known_kernel<NT><<<(unsigned)grid, NT, 0, st>>>(a);
	}
	BT_CUDA_TRY(cudaGetLastError());
	return BTBB_B200_OK;
}

/* ---------------- packed input (format B): expand a range to the byte format ---------------- */
namespace {
__global__ void unpack_kernel(const uint32_t *__restrict__ words, int64_t first, int64_t count, uint8_t *__restrict__ out)
{
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
		const int64_t p = first + i;
		out[i] = (uint8_t)((words[p >> 5] >> (p & 31)) & 1u);
	}
}
}  // namespace

/* symbols [first, first + count) of a packed stream -> ctx->d_unpack, one byte each */
static int unpack_to_bytes(btbb_b200_ctx *ctx, const uint32_t *d_words, int64_t first, int64_t count, cudaStream_t st)
{
	if (count > ctx->unpack_cap) {
		if (ctx->d_unpack) cudaFree(ctx->d_unpack);
		ctx->d_unpack = NULL; ctx->unpack_cap = 0;
		int64_t cap = count < 16384 ? 16384 : count;
		BT_CUDA_TRY(cudaMalloc(&ctx->d_unpack, (size_t)cap + 64));
		ctx->unpack_cap = cap;
	}
	int64_t blocks = (count + 255) / 256;
	if (blocks > 148 * 16) blocks = 148 * 16;
	unpack_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_words, first, count, ctx->d_unpack);
	BT_CUDA_TRY(cudaGetLastError());
	return BTBB_B200_OK;
}

/* What a scan needs set up in device memory before its kernels run -- the exact test's parameter block
 * and zeroed counters -- done by one small kernel instead of a host-to-device copy and two memsets:
 * nothing on a scan's stream then needs a copy engine, which the multi-GPU exchange keeps busy with
 * peer-to-peer copies of hit records (a 128-byte parameter copy queued behind eight 16 MB pushes was
 * what held back the next scan on 8 GPUs). */
__global__ void __launch_bounds__(256) scan_prep_kernel(sc::xparams xp, sc::xparams *slot, uint32_t *zero_a, int na,
							 unsigned long long *zero_b, int nb)
{
	if (threadIdx.x == 0 && blockIdx.x == 0) *slot = xp;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < na; i += gridDim.x * blockDim.x) zero_a[i] = 0;
	if (blockIdx.x == 0 && threadIdx.x < nb) zero_b[threadIdx.x] = 0;
}

/*
 * Dispatcher.  Every whole 4096-symbol strip that starts on a 32-byte boundary goes through a bulk
 * kernel -- scan_v7.cuh for promiscuous scans (its map hierarchy depends on the error tables the
 * context was built with), scan_known.cuh for a known LAP; the unaligned head and the ragged tail go
 * through the tile kernels above, as do known-LAP scans with k > 16 and streams shorter than a
 * strip.  BTBB_B200_OPT_TILE_KERNEL_ONLY (btbb_b200_set_option) sends everything through the tile
 * kernels (an A/B check the tests use).
 */
int bt_scan_launch(btbb_b200_ctx *ctx, const uint8_t *d_stream, int64_t n, uint32_t lap, int k,
		   btbb_b200_hit *d_out, int64_t max_hits, unsigned long long *d_count,
		   int64_t bias, cudaStream_t st)
{
	return bt_scan_launch_ex(ctx, d_stream, n, lap, k, d_out, max_hits, d_count, bias, st, NULL);
}

/* slab != NULL: try slab mode (promiscuous bulk path only); *slab->used tells the caller
 * whether it was taken and how many warps the bulk kernel ran with. */
int bt_scan_launch_ex(btbb_b200_ctx *ctx, const uint8_t *d_stream, int64_t n, uint32_t lap, int k,
		      btbb_b200_hit *d_out, int64_t max_hits, unsigned long long *d_count,
		      int64_t bias, cudaStream_t st, bt_slab_req *slab, int packed)
{
	if (slab) slab->used = 0;
	if (n <= 0) return BTBB_B200_OK;
	const bool known = lap != BTBB_B200_LAP_ANY;
	const bool k3 = !known && ctx->table_k >= 3;      /* second (3 errors) or first (4 / 5) map level in global memory */
	const bool k45 = !known && ctx->table_k >= 4;
	if (ctx->opt_tile_only || (known && k > 16) || (packed && n - 1 < sc::STRIP)) {
		if (slab) BT_CUDA_TRY(cudaMemsetAsync(d_count, 0, 2 * sizeof(unsigned long long), st));
		if (packed) {      /* no bulk kernel for this case: expand to the byte format and take the tile kernel */
			int rc0 = unpack_to_bytes(ctx, reinterpret_cast<const uint32_t *>(d_stream), 0, n + 63, st);
			if (rc0) return rc0;
			d_stream = ctx->d_unpack;
		}
		return scan_launch_v1(ctx, d_stream, n, lap, k, d_out, max_hits, d_count, bias, st);
	}
	int64_t al = packed ? 0 : (int64_t)((32 - (reinterpret_cast<uintptr_t>(d_stream) & 31)) & 31);   /* first 32-byte boundary */
	const int64_t head = al;                                                               /* first window of the bulk kernel */
	/* a strip reads 64 symbols past its end and the stream holds n + 63 */
	int64_t nstrips = n - 1 > al ? (n - 1 - al) / sc::STRIP : 0;
	/* the bulk kernels carry 32-bit positions relative to a warp's run: keep a launch below
	 * 2^31 symbols per warp (148 x 32 warps -> ~10^13 symbols); beyond that the tail kernel
	 * below simply takes the rest */
	if (nstrips > ((int64_t)1 << 31) / sc::STRIP * 4096) nstrips = ((int64_t)1 << 31) / sc::STRIP * 4096;
	if (nstrips < 1) {    /* (never with packed input: checked above) */
		if (slab) BT_CUDA_TRY(cudaMemsetAsync(d_count, 0, 2 * sizeof(unsigned long long), st));
		return scan_launch_v1(ctx, d_stream, n, lap, k, d_out, max_hits, d_count, bias, st);
	}
	const int64_t body_end = head + nstrips * sc::STRIP;
	/* packed input: the tile kernel takes the ragged tail from an expanded copy */
	const uint8_t *d_tail = d_stream + body_end;
	if (packed && body_end < n) {
		int rc0 = unpack_to_bytes(ctx, reinterpret_cast<const uint32_t *>(d_stream), body_end, n - body_end + 63, st);
		if (rc0) return rc0;
		d_tail = ctx->d_unpack;
	}
	sc::xparams xp;
	memset(&xp, 0, sizeof(xp));
	xp.cc[0] = ctx->cc[0]; xp.cc[1] = ctx->cc[1];
	xp.m32 = ctx->m32; xp.m33 = ctx->m33; xp.m0 = ctx->m0;
	xp.kmax = k; xp.err_log2 = ctx->err_log2; xp.err = ctx->d_err; xp.map2g = ctx->d_map7g;
	xp.hits = d_out; xp.max_hits = max_hits; xp.count = d_count; xp.bias = bias;
	const int bulk_warps = sc::WARPS;
	int64_t grid = ctx->sm_count;
	const int64_t need = (nstrips + bulk_warps - 1) / bulk_warps;
	if (grid > need) grid = need;
	const bool slab_mode = slab != NULL;
	if (slab_mode) {
		const int nw = (int)grid * bulk_warps;
		int rc2 = bt_ensure_slab(ctx, nw + 2);
		if (rc2) return rc2;
		xp.slab = ctx->d_slab; xp.slab_cnt = ctx->d_slab_cnt; xp.slab_cap = BT_SLAB_CAP;
		slab->used = 1; slab->nw = nw;
	}
	if (!ctx->d_xp)
		BT_CUDA_TRY(cudaMalloc(&ctx->d_xp, 16 * 128));
	static_assert(sizeof(sc::xparams) <= 128, "xparams slot");
	void *slot = (char *)ctx->d_xp + 128 * (ctx->xp_next++ & 15);
	{
		/* slab mode: the per-slab fill counts and the two 64-bit edge counters behind them, plus the
		 * context's hit counter pair (the caller's d_count) */
		const int na = slab_mode ? ((slab->nw + 2 + 1) & ~1) + 4 : 0;
		scan_prep_kernel<<<slab_mode ? 8 : 1, 256, 0, st>>>(xp, static_cast<sc::xparams *>(slot), ctx->d_slab_cnt, na,
								   slab_mode ? d_count : NULL, slab_mode ? 2 : 0);
	}
	if (known) {
		/* known LAP: bit-sliced prefilter on 16 sync-word bits that are all 0 (or all 1) */
		vk::args a;
		a.base = d_stream + al; a.pos0 = al; a.nstrips = nstrips;
		a.ac = bt_gen_syncword(lap); a.lap = lap; a.kmax = k;
		a.kk = k < 0 ? -1 : (k > 16 ? 16 : k);
		a.xp = (const sc::xparams *)slot;
		/* each 32-bit half of the sync word has at least 16 zeros or 16 ones */
		const uint32_t lo = (uint32_t)a.ac, hi = (uint32_t)(a.ac >> 32);
		const bool inv = __builtin_popcount(lo) > 16, inv2 = __builtin_popcount(hi) > 16;
		int cnt = 0;
		for (int j = 0; j < 32 && cnt < 16; j++)
			if (((lo >> j) & 1u) == (inv ? 1u : 0u)) a.sh[cnt++] = (uint32_t)j;
		cnt = 0;
		for (int j = 0; j < 32 && cnt < 16; j++)
			if (((hi >> j) & 1u) == (inv2 ? 1u : 0u)) a.sh2[cnt++] = (uint32_t)j;
		const bool two = k >= 3;
		if (two && a.kk > 16) a.kk = 16;
		void (*kern)(const vk::args);
		if (!packed) {
			if (!two) kern = inv ? vk::scan_known_v4<true, false, false> : vk::scan_known_v4<false, false, false>;
			else if (inv) kern = inv2 ? vk::scan_known_v4<true, true, true> : vk::scan_known_v4<true, true, false>;
			else kern = inv2 ? vk::scan_known_v4<false, true, true> : vk::scan_known_v4<false, true, false>;
		} else {
			if (!two) kern = inv ? vk::scan_known_v4<true, false, false, true> : vk::scan_known_v4<false, false, false, true>;
			else if (inv) kern = inv2 ? vk::scan_known_v4<true, true, true, true> : vk::scan_known_v4<true, true, false, true>;
			else kern = inv2 ? vk::scan_known_v4<false, true, true, true> : vk::scan_known_v4<false, true, false, true>;
		}
		if (ctx->prof_on) cudaEventRecord(ctx->prof_ev[0], st);
		kern<<<(unsigned)grid, vk::WARPS * 32, vk::SMEM_BYTES, st>>>(a);
		if (ctx->prof_on) { cudaEventRecord(ctx->prof_ev[1], st); ctx->prof_valid = 1; }
	} else {
		/* tables for <= 2 errors: both map levels in shared memory, table A by byte permute (layout<1>);
		 * 3 errors: the 64 KiB first-level map (12 % of it set) and a global second level;
		 * 4 / 5 errors: first level in global memory, positives straight to the exact test */
		v7::args a;
		const bool ta = !k3;
		a.base = d_stream + al; a.pos0 = al; a.nstrips = nstrips;
		a.lut = ctx->d_lut7; a.map = ta ? ctx->d_map7b : ctx->d_map7; a.xp = (const sc::xparams *)slot;
		a.m1 = 0xffffffffu; a.c64 = 64u;
		a.map1g = reinterpret_cast<const uint8_t *>(ctx->d_map7g); a.m1g_shift = 32 - (ctx->map7g_log2 - 3);
		void (*kern)(const v7::args);
		if (packed) kern = k45 ? v7::scan_promisc_v7<0, 5, 0, 2, true> : k3 ? v7::scan_promisc_v7<0, 5, 0, 1, true>
				       : v7::scan_promisc_v7<0, 5, 1, 0, true>;
		else kern = k45 ? v7::scan_promisc_v7<0, 5, 0, 2> : k3 ? v7::scan_promisc_v7<0, 5, 0, 1> : v7::scan_promisc_v7<0, 5, 1>;
		const size_t smem = ta ? v7::layout<1>::smem_bytes : v7::layout<0>::smem_bytes;
		BT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		if (k45) {
			/* the global first-level map is probed once per candidate at random: keep it in the
			 * persisting part of L2 while the stream (10^3 times its size) flows past it */
			const size_t map_bytes = (size_t)1 << (ctx->map7g_log2 - 3);
			if (!ctx->l2_persist_set) {
				int max_persist = 0;
				cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, ctx->device);
				size_t want = map_bytes + (map_bytes >> 2);
				if (want > (size_t)max_persist) want = (size_t)max_persist;
				if (want) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
				ctx->l2_persist_bytes = want;
				ctx->l2_persist_set = 1;
			}
			if (ctx->l2_persist_bytes) {
				cudaStreamAttrValue av;
				memset(&av, 0, sizeof(av));
				av.accessPolicyWindow.base_ptr = ctx->d_map7g;
				av.accessPolicyWindow.num_bytes = map_bytes;
				av.accessPolicyWindow.hitRatio = ctx->l2_persist_bytes >= map_bytes ? 1.0f : (float)ctx->l2_persist_bytes / (float)map_bytes;
				av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
				av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
				cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av);
			}
		}
		if (ctx->prof_on) cudaEventRecord(ctx->prof_ev[0], st);
		kern<<<(unsigned)grid, v7::WARPS * 32, smem, st>>>(a);
		if (ctx->prof_on) { cudaEventRecord(ctx->prof_ev[1], st); ctx->prof_valid = 1; }
		if (k45 && ctx->l2_persist_bytes) {
			cudaStreamAttrValue av;
			memset(&av, 0, sizeof(av));
			av.accessPolicyWindow.num_bytes = 0;
			cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av);
		}
	}
	BT_CUDA_TRY(cudaGetLastError());
	int rc = BTBB_B200_OK;
	if (slab_mode) {
		/* head / tail go to slabs nw and nw+1; their 64-bit counters sit behind the 32-bit ones */
		const int nw = slab->nw;
		unsigned long long *edge_cnt = (unsigned long long *)(ctx->d_slab_cnt + ((nw + 2 + 1) & ~1));
		if (head > 0)
			rc = scan_launch_v1(ctx, d_stream, head, lap, k, ctx->d_slab + (size_t)nw * BT_SLAB_CAP, BT_SLAB_CAP, edge_cnt, bias, st);
		if (!rc && body_end < n)
			rc = scan_launch_v1(ctx, d_tail, n - body_end, lap, k, ctx->d_slab + (size_t)(nw + 1) * BT_SLAB_CAP,
					    BT_SLAB_CAP, edge_cnt + 1, bias + body_end, st);
		return rc;
	}
	if (head > 0)
		rc = scan_launch_v1(ctx, d_stream, head, lap, k, d_out, max_hits, d_count, bias, st);
	if (!rc && body_end < n)
		rc = scan_launch_v1(ctx, d_tail, n - body_end, lap, k, d_out, max_hits, d_count, bias + body_end, st);
	return rc;
}

/* LSD radix sort by offset; `a` holds `have` records, `b` is scratch.  *result = the buffer
 * that ends up sorted (a when the number of passes is even, else b).  The digits are taken from
 * offset - key_bias, which lies in [0, span) for the span the pass count was derived from
 * (records carry offset + bias, btbb_b200_set_offset_bias). */
int bt_sort_hits(btbb_b200_ctx *ctx, btbb_b200_hit *a, btbb_b200_hit *b, int64_t have,
		 int passes, int64_t key_bias, cudaStream_t st, btbb_b200_hit **result)
{
	btbb_b200_hit *src = a, *dst = b;
	if (have > 0) {
		int nblk = (int)((have + SORT_CHUNK - 1) / SORT_CHUNK);
		for (int p = 0; p < passes; p++) {
			sort_hist_kernel<<<nblk, 32, 0, st>>>(src, have, SORT_BITS * p, ctx->d_sort_hist, nblk, key_bias);
			uint32_t *tot = ctx->d_sort_hist + (int64_t)nblk * SORT_BINS;
			sort_rowscan_kernel<<<SORT_BINS / 8, 256, 0, st>>>(ctx->d_sort_hist, nblk, tot);
			sort_scatter_kernel<<<nblk, 32, 0, st>>>(src, dst, have, SORT_BITS * p, ctx->d_sort_hist, nblk, tot, key_bias);
			btbb_b200_hit *t = src; src = dst; dst = t;
		}
		BT_CUDA_TRY(cudaGetLastError());
	} else if (passes & 1)
		src = b;
	*result = src;
	return BTBB_B200_OK;
}

/* ---------------- slab ordering: per-warp slabs -> ascending list ---------------- */
namespace {

/* one block: slab order is head (index nw), runs 0..nw-1, tail (nw+1).  Writes the exclusive
 * bases (64-bit) into base[0..nw+2), the total into out[0] and an overflow flag into out[1]. */
__global__ void __launch_bounds__(1024) slab_scan_kernel(uint32_t *cnt, const unsigned long long *edge, int nw,
							 unsigned long long *base, unsigned long long *out,
							 volatile unsigned long long *host_out,
							 btbb_b200_hit *const *fan, int fan_n)
{
	__shared__ unsigned long long wsum[32];
	__shared__ unsigned long long carry_s;
	__shared__ int over;
	const int t = threadIdx.x, lane = t & 31, w = t >> 5;
	if (t == 0) {
		carry_s = 0; over = 0;
		/* fold the 64-bit edge counters into the 32-bit array */
		const unsigned long long h = edge[0], tl = edge[1];
		if (h > BT_SLAB_CAP || tl > BT_SLAB_CAP) over = 1;
		cnt[nw] = (uint32_t)(h > BT_SLAB_CAP ? BT_SLAB_CAP : h);
		cnt[nw + 1] = (uint32_t)(tl > BT_SLAB_CAP ? BT_SLAB_CAP : tl);
	}
	__syncthreads();
	const int total = nw + 2;
	for (int b0 = 0; b0 < total; b0 += 1024) {
		const int pos = b0 + t;                       /* position in output order */
		const int idx = pos == 0 ? nw : (pos <= nw ? pos - 1 : nw + 1);
		unsigned long long x = 0;
		if (pos < total) {
			uint32_t c = cnt[idx];
			if (c > BT_SLAB_CAP) { over = 1; c = BT_SLAB_CAP; }
			x = c;
		}
		unsigned long long inc = x;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const unsigned long long u = __shfl_up_sync(0xffffffffu, inc, d);
			if (lane >= d) inc += u;
		}
		if (lane == 31) wsum[w] = inc;
		__syncthreads();
		if (w == 0) {
			unsigned long long y = wsum[lane], z = y;
			#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const unsigned long long u = __shfl_up_sync(0xffffffffu, z, d);
				if (lane >= d) z += u;
			}
			wsum[lane] = z - y;
		}
		__syncthreads();
		const unsigned long long excl = carry_s + wsum[w] + inc - x;
		if (pos < total) base[idx] = excl;
		__syncthreads();
		if (t == 1023) carry_s = excl + x;
		__syncthreads();
	}
	if (t == 0) {
		out[0] = carry_s; out[1] = (unsigned long long)over;
		/* the same two words straight into pinned host memory (no device-to-host copy on the stream) */
		host_out[0] = carry_s; host_out[1] = (unsigned long long)over;
		__threadfence_system();
	}
	/* fan-out (multi-GPU): the record in front of every destination list is its header, the hit count */
	if (t < fan_n) {
		btbb_b200_hit hdr;
		hdr.offset = (int64_t)carry_s; hdr.lap = 0; hdr.ac_errors = 0; hdr.pad[0] = hdr.pad[1] = hdr.pad[2] = 0;
		fan[t][-1] = hdr;
	}
}

/* one block per slab: bitonic sort of <= BT_SLAB_CAP records by offset, written to its place */
/* fan / fan_n (multi-GPU, sharded.cu): every record is also stored into fan_n further lists -- this rank's
 * slot of the gather buffer on every GPU of the node, peer memory written straight from the kernel over
 * NVLink: the all-gather of the hit records is part of the ordering pass */
__global__ void __launch_bounds__(256) slab_sort_kernel(const btbb_b200_hit *slab, const uint32_t *cnt,
							const unsigned long long *base, btbb_b200_hit *out, int64_t max_hits,
							btbb_b200_hit *const *fan, int fan_n)
{
	__shared__ long long key[BT_SLAB_CAP];
	__shared__ unsigned long long val[BT_SLAB_CAP];
	const int s = blockIdx.x;
	uint32_t n = cnt[s];
	if (n == 0) return;
	if (n > BT_SLAB_CAP) n = BT_SLAB_CAP;
	uint32_t N = 32;
	while (N < n) N <<= 1;
	const btbb_b200_hit *src = slab + (size_t)s * BT_SLAB_CAP;
	for (uint32_t i = threadIdx.x; i < N; i += 256) {
		if (i < n) {
			const btbb_b200_hit h = src[i];
			key[i] = h.offset; val[i] = ((unsigned long long)h.lap << 8) | h.ac_errors;
		} else { key[i] = 0x7fffffffffffffffLL; val[i] = 0; }
	}
	__syncthreads();
	for (uint32_t k = 2; k <= N; k <<= 1)
		for (uint32_t j = k >> 1; j > 0; j >>= 1) {
			for (uint32_t i = threadIdx.x; i < N; i += 256) {
				const uint32_t l = i ^ j;
				if (l > i) {
					const bool up = (i & k) == 0;
					const long long a = key[i], b = key[l];
					if ((a > b) == up) {
						key[i] = b; key[l] = a;
						const unsigned long long t = val[i]; val[i] = val[l]; val[l] = t;
					}
				}
			}
			__syncthreads();
		}
	const unsigned long long b0 = base[s];
	for (uint32_t i = threadIdx.x; i < n; i += 256) {
		if ((int64_t)(b0 + i) < max_hits) {
			btbb_b200_hit h;
			h.offset = key[i]; h.lap = (uint32_t)(val[i] >> 8); h.ac_errors = (uint8_t)(val[i] & 0xff);
			h.pad[0] = h.pad[1] = h.pad[2] = 0;
			out[b0 + i] = h;
			for (int j = 0; j < fan_n; j++) fan[j][b0 + i] = h;
		}
	}
}

}  // namespace

int bt_ensure_slab(btbb_b200_ctx *ctx, int nslabs)
{
	if (nslabs > ctx->slab_n) {
		if (ctx->d_slab) cudaFree(ctx->d_slab);
		if (ctx->d_slab_cnt) cudaFree(ctx->d_slab_cnt);
		if (ctx->d_slab_base) cudaFree(ctx->d_slab_base);
		ctx->d_slab = NULL; ctx->d_slab_cnt = NULL; ctx->d_slab_base = NULL; ctx->slab_n = 0;
		BT_CUDA_TRY(cudaMalloc(&ctx->d_slab, (size_t)nslabs * BT_SLAB_CAP * sizeof(btbb_b200_hit)));
		BT_CUDA_TRY(cudaMalloc(&ctx->d_slab_cnt, (size_t)(nslabs + 2) * sizeof(uint32_t) + 2 * sizeof(unsigned long long)));
		BT_CUDA_TRY(cudaMalloc(&ctx->d_slab_base, (size_t)nslabs * sizeof(unsigned long long)));
		ctx->slab_n = nslabs;
	}
	return BTBB_B200_OK;
}

int bt_sort_passes(int64_t span)
{
	int bits = 1;
	while (bits < 63 && ((int64_t)1 << bits) < span) bits++;
	return (bits + SORT_BITS - 1) / SORT_BITS;
}

int bt_ensure_tmp(btbb_b200_ctx *ctx, int64_t hits)
{
	if (hits > ctx->tmp_cap) {
		if (ctx->d_tmp) cudaFree(ctx->d_tmp);
		ctx->d_tmp = NULL; ctx->tmp_cap = 0;
		BT_CUDA_TRY(cudaMalloc(&ctx->d_tmp, (size_t)hits * sizeof(btbb_b200_hit)));
		ctx->tmp_cap = hits;
	}
	int64_t nblk = (hits + SORT_CHUNK - 1) / SORT_CHUNK;
	if ((nblk + 1) * SORT_BINS > ctx->sort_hist_cap) {
		if (ctx->d_sort_hist) cudaFree(ctx->d_sort_hist);
		ctx->d_sort_hist = NULL; ctx->sort_hist_cap = 0;
		BT_CUDA_TRY(cudaMalloc(&ctx->d_sort_hist, (size_t)(nblk + 1) * SORT_BINS * sizeof(uint32_t)));
		ctx->sort_hist_cap = (nblk + 1) * SORT_BINS;
	}
	return BTBB_B200_OK;
}

extern "C" int btbb_b200_find_ac_enqueue(btbb_b200_ctx *ctx, const uint8_t *d_stream, int64_t search_length,
					 uint32_t lap, int max_ac_errors, btbb_b200_hit *d_hits,
					 int64_t max_hits, unsigned long long *d_count, void *cuda_stream)
{
	if (!ctx || (!d_stream && search_length > 0) || search_length < 0 || max_hits < 0 || !d_count ||
	    (lap != BTBB_B200_LAP_ANY && lap > 0xffffffu))
		return btbb_b200_set_error(BTBB_B200_EINVAL, "find_ac: bad arguments");
	BT_CUDA_TRY(cudaSetDevice(ctx->device));
	BT_CUDA_TRY(cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), (cudaStream_t)cuda_stream));
	return bt_scan_launch(ctx, d_stream, search_length, lap, max_ac_errors, d_hits, max_hits, d_count,
			      0, (cudaStream_t)cuda_stream);
}

extern "C" int btbb_b200_find_ac_dev(btbb_b200_ctx *ctx, const uint8_t *d_stream, int64_t search_length,
				     uint32_t lap, int max_ac_errors, btbb_b200_hit *d_hits,
				     int64_t max_hits, int64_t *n_hits, void *cuda_stream)
{
	if (!ctx || !n_hits || (!d_hits && max_hits > 0) || (!d_stream && search_length > 0) ||
	    search_length < 0 || max_hits < 0 || (lap != BTBB_B200_LAP_ANY && lap > 0xffffffu))
		return btbb_b200_set_error(BTBB_B200_EINVAL, "find_ac: bad arguments");
	return bt_find_ac_dev_impl(ctx, d_stream, 0, search_length, lap, max_ac_errors, d_hits, max_hits, n_hits,
				   (cudaStream_t)cuda_stream);
}

extern "C" int btbb_b200_find_ac_packed_dev(btbb_b200_ctx *ctx, const uint32_t *d_words, int64_t search_length,
					    uint32_t lap, int max_ac_errors, btbb_b200_hit *d_hits,
					    int64_t max_hits, int64_t *n_hits, void *cuda_stream)
{
	if (!ctx || !n_hits || (!d_hits && max_hits > 0) || (!d_words && search_length > 0) ||
	    search_length < 0 || max_hits < 0 || (lap != BTBB_B200_LAP_ANY && lap > 0xffffffu) ||
	    (reinterpret_cast<uintptr_t>(d_words) & 3))
		return btbb_b200_set_error(BTBB_B200_EINVAL, "find_ac_packed: bad arguments");
	return bt_find_ac_dev_impl(ctx, reinterpret_cast<const uint8_t *>(d_words), 1, search_length, lap, max_ac_errors,
				   d_hits, max_hits, n_hits, (cudaStream_t)cuda_stream);
}

/* device-resident stream (bytes, or packed words when `packed`) -> ascending hit list in d_hits.
 * Split in two so that a caller can overlap its own host work (or the next call's set-up) with
 * the kernels: begin() enqueues everything the common case needs -- scan, slab scan, slab sort,
 * the read-back of the counters into pinned memory -- and returns; end() waits, and only in the
 * uncommon cases (no slab ordering for this call, or a slab overflowed) runs the generic path. */
/* swap the context's per-scan scratch with the parked lane's so that lane `want` is the active one */
static void bt_lane_select(btbb_b200_ctx *ctx, int want)
{
	if (ctx->lane_active == want) return;
	bt_lane cur;
	cur.d_count = ctx->d_count; cur.d_tmp = ctx->d_tmp; cur.tmp_cap = ctx->tmp_cap;
	cur.d_sort_hist = ctx->d_sort_hist; cur.sort_hist_cap = ctx->sort_hist_cap;
	cur.d_slab = ctx->d_slab; cur.d_slab_cnt = ctx->d_slab_cnt; cur.d_slab_base = ctx->d_slab_base; cur.slab_n = ctx->slab_n;
	cur.h_res = ctx->h_res; cur.pending = ctx->pending; cur.ev_done = ctx->ev_done;
	cur.prof_ev[0] = ctx->prof_ev[0]; cur.prof_ev[1] = ctx->prof_ev[1]; cur.prof_valid = ctx->prof_valid;
	const bt_lane &o = ctx->parked;
	ctx->d_count = o.d_count; ctx->d_tmp = o.d_tmp; ctx->tmp_cap = o.tmp_cap;
	ctx->d_sort_hist = o.d_sort_hist; ctx->sort_hist_cap = o.sort_hist_cap;
	ctx->d_slab = o.d_slab; ctx->d_slab_cnt = o.d_slab_cnt; ctx->d_slab_base = o.d_slab_base; ctx->slab_n = o.slab_n;
	ctx->h_res = o.h_res; ctx->pending = o.pending; ctx->ev_done = o.ev_done;
	ctx->prof_ev[0] = o.prof_ev[0]; ctx->prof_ev[1] = o.prof_ev[1]; ctx->prof_valid = o.prof_valid;
	ctx->parked = cur;
	ctx->lane_active = want;
}

/* before anything that uses the scratch outside begin / end (host-buffer entry points, destroy): lane 0 */
void bt_lane_reset(btbb_b200_ctx *ctx) { bt_lane_select(ctx, 0); }

static int find_ac_begin_lane(btbb_b200_ctx *ctx, const uint8_t *d_stream, int packed, int64_t search_length, uint32_t lap,
			      int max_ac_errors, btbb_b200_hit *d_hits, int64_t max_hits, cudaStream_t st);
static int find_ac_end_lane(btbb_b200_ctx *ctx, int64_t *n_hits);

/* Up to two scans may be pending: begin takes the free lane, end completes the OLDEST pending scan. */
int bt_find_ac_dev_begin(btbb_b200_ctx *ctx, const uint8_t *d_stream, int packed, int64_t search_length, uint32_t lap,
			 int max_ac_errors, btbb_b200_hit *d_hits, int64_t max_hits, cudaStream_t st)
{
	if (ctx->lane_count >= 2)
		return btbb_b200_set_error(BTBB_B200_EINVAL, "find_ac: two calls are already pending on this context");
	BT_CUDA_TRY(cudaSetDevice(ctx->device));
	bt_lane_select(ctx, (ctx->lane_head + ctx->lane_count) & 1);
	if (!ctx->d_count) BT_CUDA_TRY(cudaMalloc(&ctx->d_count, 2 * sizeof(unsigned long long)));
	if (!ctx->ev_done) BT_CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_done, cudaEventDisableTiming));
	if (ctx->prof_on && !ctx->prof_ev[0]) {
		BT_CUDA_TRY(cudaEventCreate(&ctx->prof_ev[0]));
		BT_CUDA_TRY(cudaEventCreate(&ctx->prof_ev[1]));
	}
	const int rc = find_ac_begin_lane(ctx, d_stream, packed, search_length, lap, max_ac_errors, d_hits, max_hits, st);
	if (rc == BTBB_B200_OK) ctx->lane_count++;
	return rc;
}

int bt_find_ac_dev_end(btbb_b200_ctx *ctx, int64_t *n_hits)
{
	*n_hits = 0;
	if (ctx->lane_count == 0)
		return btbb_b200_set_error(BTBB_B200_EINVAL, "find_ac: no call is pending on this context");
	bt_lane_select(ctx, ctx->lane_head);
	ctx->lane_head ^= 1;
	ctx->lane_count--;
	return find_ac_end_lane(ctx, n_hits);
}

static int find_ac_begin_lane(btbb_b200_ctx *ctx, const uint8_t *d_stream, int packed, int64_t search_length, uint32_t lap,
			      int max_ac_errors, btbb_b200_hit *d_hits, int64_t max_hits, cudaStream_t st)
{
	bt_pending &pd = ctx->pending;
	if (pd.mode != BT_PENDING_NONE)
		return btbb_b200_set_error(BTBB_B200_EINVAL, "find_ac: a call is already pending on this context");
	pd.d_stream = d_stream; pd.packed = packed; pd.search_length = search_length; pd.lap = lap;
	pd.max_ac_errors = max_ac_errors; pd.d_hits = d_hits; pd.max_hits = max_hits; pd.st = st;
	pd.bias = ctx->hit_bias; pd.fanned = 0;
	pd.mode = BT_PENDING_GENERIC;
	if (search_length == 0) { pd.mode = BT_PENDING_EMPTY; return BTBB_B200_OK; }
	int rc = bt_ensure_tmp(ctx, max_hits > 0 ? max_hits : 1);
	if (rc) { pd.mode = BT_PENDING_NONE; return rc; }
	if (!ctx->h_res) {
		cudaError_t e = cudaMallocHost(&ctx->h_res, 2 * sizeof(unsigned long long));
		if (e != cudaSuccess) { pd.mode = BT_PENDING_NONE; return btbb_b200_cuda_fail(e, "cudaMallocHost(find_ac counters)"); }
	}
	if (lap != BTBB_B200_LAP_ANY) return BTBB_B200_OK;       /* known LAP: the generic path, in end() */
	/* fast ordering: per-warp slabs, one small sort per slab (promiscuous bulk path only) */
	bt_slab_req req;
	cudaError_t e = cudaSuccess;
	/* (bt_scan_launch_ex zeroes the counters: in the bulk path's set-up kernel, else with a memset) */
	rc = bt_scan_launch_ex(ctx, d_stream, search_length, lap, max_ac_errors, ctx->d_tmp, max_hits, ctx->d_count, pd.bias, st, &req, packed);
	if (rc) { pd.mode = BT_PENDING_NONE; return rc; }
	if (req.used) {
		const int nw = req.nw;
		unsigned long long *edge_cnt = (unsigned long long *)(ctx->d_slab_cnt + ((nw + 2 + 1) & ~1));
		slab_scan_kernel<<<1, 1024, 0, st>>>(ctx->d_slab_cnt, edge_cnt, nw, ctx->d_slab_base, ctx->d_count, ctx->h_res, ctx->d_fan, ctx->fan_n);
		/* the sort is safe to run even if a slab overflowed (counts are clamped); its output is
		 * then simply not used */
		slab_sort_kernel<<<nw + 2, 256, 0, st>>>(ctx->d_slab, ctx->d_slab_cnt, ctx->d_slab_base, d_hits, max_hits, ctx->d_fan, ctx->fan_n);
		pd.fanned = ctx->fan_n > 0;
		pd.mode = BT_PENDING_SLAB;
	} else
		pd.mode = BT_PENDING_UNORDERED;      /* the unordered list is in d_tmp, its length in d_count[0] */
	/* slab mode: slab_scan_kernel has stored the counters into the pinned host words itself */
	if (pd.mode != BT_PENDING_SLAB)
		e = cudaMemcpyAsync(ctx->h_res, ctx->d_count, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st);
	if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_done, st);
	if (e == cudaSuccess) e = cudaGetLastError();
	if (e != cudaSuccess) { pd.mode = BT_PENDING_NONE; return btbb_b200_cuda_fail(e, "find_ac: enqueue"); }
	return BTBB_B200_OK;
}

static int find_ac_end_lane(btbb_b200_ctx *ctx, int64_t *n_hits)
{
	bt_pending pd = ctx->pending;
	ctx->pending.mode = BT_PENDING_NONE;
	ctx->last_fanned = 0;
	*n_hits = 0;
	if (pd.mode == BT_PENDING_NONE)
		return btbb_b200_set_error(BTBB_B200_EINVAL, "find_ac: no call is pending on this context");
	if (pd.mode == BT_PENDING_EMPTY) return BTBB_B200_OK;
	BT_CUDA_TRY(cudaSetDevice(ctx->device));
	cudaStream_t st = pd.st;
	const uint8_t *d_stream = pd.d_stream;
	const int packed = pd.packed, max_ac_errors = pd.max_ac_errors;
	const int64_t search_length = pd.search_length, max_hits = pd.max_hits;
	const uint32_t lap = pd.lap;
	btbb_b200_hit *d_hits = pd.d_hits;
	unsigned long long total = 0;
	int rc;
	if (pd.mode == BT_PENDING_SLAB || pd.mode == BT_PENDING_UNORDERED) {
		/* this scan's own completion, not the stream's: a second scan may already be queued behind it */
		BT_CUDA_TRY(cudaEventSynchronize(ctx->ev_done));
		if (ctx->h_res[0] >> 62)
			return btbb_b200_set_error(BTBB_B200_ECUDA, "find_ac: unexpected shared-memory window layout");
		if (pd.mode == BT_PENDING_SLAB) {
			if (!ctx->h_res[1]) {
				total = ctx->h_res[0];
				*n_hits = (int64_t)total;
				ctx->last_fanned = pd.fanned && (int64_t)total <= max_hits;
				if ((int64_t)total > max_hits)
					return btbb_b200_set_error(BTBB_B200_EOVERFLOW, "find_ac: hit buffer too small");
				return BTBB_B200_OK;
			}
			/* a slab overflowed (very dense hits): fall through to the generic path */
		} else {
			total = ctx->h_res[0];
			*n_hits = (int64_t)total;
			int64_t have = (int64_t)total < max_hits ? (int64_t)total : max_hits;
			btbb_b200_hit *res = NULL;
			rc = bt_sort_hits(ctx, ctx->d_tmp, d_hits, have, bt_sort_passes(search_length), pd.bias, st, &res);
			if (rc) return rc;
			if (res != d_hits && have > 0)
				BT_CUDA_TRY(cudaMemcpyAsync(d_hits, res, (size_t)have * sizeof(btbb_b200_hit), cudaMemcpyDeviceToDevice, st));
			BT_CUDA_TRY(cudaStreamSynchronize(st));
			if ((int64_t)total > max_hits)
				return btbb_b200_set_error(BTBB_B200_EOVERFLOW, "find_ac: hit buffer too small");
			return BTBB_B200_OK;
		}
	}
	/* generic path: the scan writes into whichever buffer makes the last scatter land in d_hits */
	int passes = bt_sort_passes(search_length);
	btbb_b200_hit *first = (passes & 1) ? ctx->d_tmp : d_hits;
	btbb_b200_hit *other = (passes & 1) ? d_hits : ctx->d_tmp;
	BT_CUDA_TRY(cudaMemsetAsync(ctx->d_count, 0, sizeof(unsigned long long), st));
	rc = bt_scan_launch_ex(ctx, d_stream, search_length, lap, max_ac_errors, first, max_hits, ctx->d_count, pd.bias, st, NULL, packed);
	if (rc) return rc;
	BT_CUDA_TRY(cudaMemcpyAsync(&total, ctx->d_count, sizeof(total), cudaMemcpyDeviceToHost, st));
	BT_CUDA_TRY(cudaStreamSynchronize(st));
	if (total >> 62)
		return btbb_b200_set_error(BTBB_B200_ECUDA, "find_ac: unexpected shared-memory window layout");
	*n_hits = (int64_t)total;
	int64_t have = (int64_t)total < max_hits ? (int64_t)total : max_hits;
	btbb_b200_hit *res = NULL;
	rc = bt_sort_hits(ctx, first, other, have, passes, pd.bias, st, &res);
	if (rc) return rc;
	BT_CUDA_TRY(cudaStreamSynchronize(st));
	if ((int64_t)total > max_hits)
		return btbb_b200_set_error(BTBB_B200_EOVERFLOW, "find_ac: hit buffer too small");
	return BTBB_B200_OK;
}

int bt_find_ac_dev_impl(btbb_b200_ctx *ctx, const uint8_t *d_stream, int packed, int64_t search_length, uint32_t lap,
			int max_ac_errors, btbb_b200_hit *d_hits, int64_t max_hits, int64_t *n_hits, cudaStream_t st)
{
	*n_hits = 0;
	int rc = bt_find_ac_dev_begin(ctx, d_stream, packed, search_length, lap, max_ac_errors, d_hits, max_hits, st);
	if (rc) return rc;
	return bt_find_ac_dev_end(ctx, n_hits);
}

extern "C" int btbb_b200_find_ac_dev_begin(btbb_b200_ctx *ctx, const uint8_t *d_stream, int64_t search_length,
					   uint32_t lap, int max_ac_errors, btbb_b200_hit *d_hits, int64_t max_hits,
					   void *cuda_stream)
{
	if (!ctx || (!d_hits && max_hits > 0) || (!d_stream && search_length > 0) ||
	    search_length < 0 || max_hits < 0 || (lap != BTBB_B200_LAP_ANY && lap > 0xffffffu))
		return btbb_b200_set_error(BTBB_B200_EINVAL, "find_ac_begin: bad arguments");
	return bt_find_ac_dev_begin(ctx, d_stream, 0, search_length, lap, max_ac_errors, d_hits, max_hits, (cudaStream_t)cuda_stream);
}

extern "C" int btbb_b200_find_ac_dev_end(btbb_b200_ctx *ctx, int64_t *n_hits)
{
	if (!ctx || !n_hits) return btbb_b200_set_error(BTBB_B200_EINVAL, "find_ac_end: bad arguments");
	return bt_find_ac_dev_end(ctx, n_hits);
}

extern "C" int btbb_b200_set_option(btbb_b200_ctx *ctx, int option, int64_t value)
{
	if (!ctx) return btbb_b200_set_error(BTBB_B200_EINVAL, "set_option: bad arguments");
	switch (option) {
	case BTBB_B200_OPT_TILE_KERNEL_ONLY: ctx->opt_tile_only = value != 0; break;
	case BTBB_B200_OPT_HOST_BYTE_ROUTE: ctx->opt_host_bytes = value != 0; break;
	case BTBB_B200_OPT_HOST_SPLIT_PERMILLE: ctx->opt_host_split = (int)(value < 0 ? 0 : value > 950 ? 950 : value); break;
	case BTBB_B200_OPT_PACK_THREADS: ctx->opt_pack_threads = (int)(value < 0 ? 0 : value > 128 ? 128 : value); break;
	case BTBB_B200_OPT_TRACE: ctx->opt_trace = value != 0; break;
	case BTBB_B200_OPT_DECODE_WIDE_STAGING: ctx->opt_decode_wide = value != 0; break;
	case BTBB_B200_OPT_PACK_STREAMS: ctx->opt_pack_streams = (int)(value < 1 ? 1 : value > 8 ? 8 : value); break;
	default: return btbb_b200_set_error(BTBB_B200_EINVAL, "set_option: unknown option");
	}
	return BTBB_B200_OK;
}

/* measurement hook: CUDA events around the bulk scan kernel of every following scan, on its stream */
extern "C" int btbb_b200_set_profiling(btbb_b200_ctx *ctx, int on)
{
	if (!ctx) return btbb_b200_set_error(BTBB_B200_EINVAL, "set_profiling: bad arguments");
	ctx->prof_on = on != 0;
	ctx->prof_valid = 0; ctx->parked.prof_valid = 0;      /* the events themselves are created per lane, at its next scan */
	return BTBB_B200_OK;
}

extern "C" int btbb_b200_last_scan_kernel_ms(btbb_b200_ctx *ctx, float *ms)
{
	if (!ctx || !ms || !ctx->prof_valid) return btbb_b200_set_error(BTBB_B200_EINVAL, "last_scan_kernel_ms: no profiled scan");
	BT_CUDA_TRY(cudaEventElapsedTime(ms, ctx->prof_ev[0], ctx->prof_ev[1]));
	return BTBB_B200_OK;
}

extern "C" int btbb_b200_set_offset_bias(btbb_b200_ctx *ctx, int64_t bias)
{
	if (!ctx) return btbb_b200_set_error(BTBB_B200_EINVAL, "set_offset_bias: bad arguments");
	ctx->hit_bias = bias;
	return BTBB_B200_OK;
}

