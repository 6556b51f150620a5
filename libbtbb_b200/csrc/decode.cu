/*
 * decode.cu -- K4/K5: the per-packet chain  unfec13 -> unwhiten -> HEC  and
 * unfec23 / unwhiten / CRC  (bluetooth_packet.c:552-705, 708-1317) for a batch of detected
 * packets, one warp per packet.  The arithmetic is decode_core.h (shared with the host path of
 * the classic single-packet calls); this file is the warp-level plumbing around it:
 *
 *   ingest   128-bit loads from the 16-byte aligned base below the packet, 16 symbols -> 16 bits
 *            with one multiply per 4 symbols, pairs of lanes combined by shuffle -> packed words
 *            in shared memory (symbols past `length` read 0 like a fresh btbb_packet).  The header
 *            part (512 symbols) comes first; it decides which packet types are in play and with
 *            them how much of the packet is needed at all.
 *   once per packet, independent of the clock: the FEC 1/3 votes (unfec13 :552-568), every
 *            needed FEC 2/3 block (unfec23 :585-649; lane = block), and the CRC prefix tables of
 *            each bit source (lane = run of bytes, XOR scan across lanes).
 *   per clock  lane = CLK1-6 candidate (try_clock + crc_check, :1178-1195 / :708-769; two rounds of
 *            32) or all lanes on the packet's own clock (btbb_decode_header + _payload,
 *            :1198-1297).  A payload CRC is one table test; the length searches of EV3 / EV4 / EV5
 *            and the 33 clocks of fhs are spread over the lanes (ballot = first match).
 *   output   one 372-byte record per packet written as 93 coalesced words, or 64 records per
 *            packet staged in shared memory, 32 at a time, and stored with one bulk copy
 *            (cp.async.bulk shared -> global, 11 904 bytes), or the UAP sieve's 16-bit words.
 */
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>
#include "bt_math.h"
#include "decode_core.h"
#include "capi_internal.h"

const btd_tables *btd_host_tables();
int bt_decode_one_cpu(const char *symbols, int length, uint32_t clkn, uint8_t uap, int whitened, uint8_t type,
		      int mode, btbb_b200_decoded *out);

namespace {

enum { OUT_ONE = 0, OUT_FULL64 = 1, OUT_TC16 = 2 };
constexpr int REC_WORDS = (int)(sizeof(btbb_b200_decoded) / 4);      /* 93 */
constexpr int STAGE_BYTES = 32 * (int)sizeof(btbb_b200_decoded);     /* 11 904 = 16 x 744 */
constexpr int PKT_BYTES = (int)((sizeof(btd_pkt) + 15) & ~(size_t)15);
constexpr int NIB_BYTES = BTD_LMAX * 2 * 16 * 2;
constexpr int SMALL_BYTES = (int)((sizeof(btd_small_tables) + 15) & ~(size_t)15);
constexpr unsigned FULL = 0xffffffffu;
static_assert(sizeof(btbb_b200_decoded) == 372 && STAGE_BYTES % 16 == 0, "record layout");
static_assert(NIB_BYTES % 16 == 0 && SMALL_BYTES % 16 == 0, "table layout");

struct dec_args {
	const uint8_t *stream;
	int64_t stream_len;
	const btbb_b200_pkt_in *pkts;
	int64_t n;
	int mode, raw_payload;
	btbb_b200_decoded *out;
	uint16_t *tc16;
	const int64_t *idx;                  /* work item i is packet idx[i] (UAP sieve rounds) */
	const unsigned long long *n_dev;     /* item count in device memory */
	const btd_tables *tables;
};

template <int OUT, int WARPS, int SREC>
constexpr int smem_bytes() { return SMALL_BYTES + NIB_BYTES + WARPS * (PKT_BYTES + (OUT == OUT_FULL64 ? SREC * (int)sizeof(btbb_b200_decoded) : 0)); }

__device__ __forceinline__ uint4 ld_stream(const uint4 *p)
{
	uint4 v;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
		     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
	return v;
}

/* raw words [w_begin, w_end) (multiples of 16) from 16-symbol chunks of the aligned base; the loads of
 * up to NB iterations are issued back to back before the first one is used */
template <int NB>
__device__ __forceinline__ void ingest(btd_pkt &P, const uint4 *base, int sh, int length, int w_begin, int w_end, int lane)
{
	for (int w0 = w_begin; w0 < w_end; w0 += 16 * NB) {
		uint4 v[NB];
		uint32_t m[NB];
		#pragma unroll
		for (int i = 0; i < NB; i++) {
			const int c = 2 * (w0 + 16 * i) + lane;
			m[i] = w0 + 16 * i < w_end ? btd_chunk_mask(c, sh, length) : 0u;
			v[i] = make_uint4(0, 0, 0, 0);
			if (m[i]) v[i] = ld_stream(base + c);
		}
		#pragma unroll
		for (int i = 0; i < NB; i++) {
			if (w0 + 16 * i < w_end) {
				const uint32_t h = btd_pack16(v[i].x, v[i].y, v[i].z, v[i].w) & m[i];
				const uint32_t hi = __shfl_down_sync(FULL, h, 1);
				if (!(lane & 1)) P.raw[w0 + 16 * i + (lane >> 1)] = h | (hi << 16);
			}
		}
	}
}

/* FEC 2/3 blocks [0, nblk) from symbol `start`: lane = block; returns the first uncorrectable one */
__device__ __noinline__ int fec_blocks(const btd_pkt &P, uint32_t *dst, int start, int nblk, const uint8_t *col, int lane)
{
	const int nwords = (10 * nblk + 31) / 32 + 1;
	for (int w = lane; w < nwords; w += 32) dst[w] = 0;
	__syncwarp();
	int fail = 1 << 20;
	for (int b0 = 0; b0 < nblk; b0 += 32) {
		const int b = b0 + lane;
		const bool live = b < nblk;
		uint32_t data = 0;
		bool ok = true;
		if (live) ok = btd_fec23_block(btd_bits(P.raw, P.sh + start + 15 * b, 15), col, &data);
		const uint32_t badmask = __ballot_sync(FULL, live && !ok);
		if (badmask && fail == (1 << 20)) fail = b0 + __ffs(badmask) - 1;
		if (live && ok) {
			const int pos = 10 * b;
			atomicOr(&dst[pos >> 5], data << (pos & 31));
			if ((pos & 31) > 22) atomicOr(&dst[(pos >> 5) + 1], data >> (32 - (pos & 31)));
		}
	}
	__syncwarp();
	return fail;
}

/* dp[L] = XOR of the CRC weights of the first L bytes of a source: lane = run of bytes, XOR scan over lanes */
__device__ __noinline__ void build_dp(const btd_pkt &P, const uint16_t *nib, int src, int nbytes, uint16_t *dp, int lane)
{
	if (nbytes <= 0) { if (lane == 0) dp[0] = 0; return; }
	const int C = (nbytes + 31) >> 5;
	const int j0 = lane * C, j1 = j0 + C < nbytes ? j0 + C : nbytes;
	uint32_t acc = 0;
	for (int j = j0; j < j1; j++) {
		acc ^= btd_byte_weight(nib, j, btd_src_byte(P, src, j));
		dp[j + 1] = (uint16_t)acc;
	}
	uint32_t inc = acc;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t t = __shfl_up_sync(FULL, inc, d);
		if (lane >= d) inc ^= t;
	}
	const uint32_t base = inc ^ acc;
	for (int j = j0; j < j1; j++) dp[j + 1] ^= (uint16_t)base;
	if (lane == 0) dp[0] = 0;
}

/* first candidate in [lo, hi) that passes, all lanes working on the same search.  The fhs search is one
 * candidate clock per lane.  The length searches (EV3 / EV5: lengths 3..181, EV4: 2..121) take eight
 * consecutive lengths per lane: one 128-bit load of the packet's prefix table, one of the whitening
 * table's row, eight 16-bit compares, one ballot. */
__device__ __forceinline__ int coop_search(const btd_ctx &c, const btd_pkt &P, int pend, int q18, uint32_t uap,
					   int lo, int hi, int lane)
{
	if (pend == BTD_PEND_FHS) {
		const int cand = lo + lane;
		const uint32_t m = __ballot_sync(FULL, cand < hi && btd_cand_ok(c, P, pend, q18, uap, cand));
		return m ? lo + __ffs(m) - 1 : -1;
	}
	const uint16_t *dp = pend == BTD_PEND_EV35 ? P.dp_first8 : P.dp_fec0;
	uint4 d = *reinterpret_cast<const uint4 *>(dp + 8 * lane);
	if (c.whitened) {
		const uint4 w = __ldg(reinterpret_cast<const uint4 *>(c.wp + q18 * BTD_LMAX + 8 * lane));
		d.x ^= w.x; d.y ^= w.y; d.z ^= w.z; d.w ^= w.w;
	}
	const uint32_t want = bt_crc16_init(uap);
	const uint32_t v[4] = {d.x, d.y, d.z, d.w};
	uint32_t mine = 0;
	#pragma unroll
	for (int t = 0; t < 8; t++) {
		const uint32_t x = (t & 1) ? v[t >> 1] >> 16 : v[t >> 1] & 0xffffu;
		const int L = 8 * lane + t;
		if (x == want && L >= lo && L < hi) mine |= 1u << t;
	}
	const uint32_t m = __ballot_sync(FULL, mine != 0);
	if (!m) return -1;
	const int src = __ffs(m) - 1;
	const uint32_t bits = __shfl_sync(FULL, mine, src);
	return 8 * src + __ffs(bits) - 1;
}

/* crc_check / decode_payload / one decoder for this lane's (clock, UAP, type); searches are shared */
template <bool UNIFORM>
__device__ __forceinline__ void evaluate(const btd_ctx &c, const btd_pkt &P, btd_lane &s, int kind, int lane)
{
	btd_eval_begin(c, P, s, kind);
	int found = -1;
	if (UNIFORM) {
		if (s.pend) found = coop_search(c, P, s.pend, btd_q18(c, s.clock), s.uap, s.s_lo, s.s_hi, lane);
	} else {
		uint32_t pm = __ballot_sync(FULL, s.pend != BTD_PEND_NONE);
		while (pm) {
			const int o = __ffs(pm) - 1;
			pm &= pm - 1;
			const int pend = __shfl_sync(FULL, s.pend, o), q18 = btd_q18(c, __shfl_sync(FULL, s.clock, o));
			const uint32_t uap = __shfl_sync(FULL, s.uap, o);
			const int lo = __shfl_sync(FULL, s.s_lo, o), hi = __shfl_sync(FULL, s.s_hi, o);
			const int f = coop_search(c, P, pend, q18, uap, lo, hi, lane);
			if (lane == o) found = f;
		}
	}
	btd_eval_end(P, s, kind, found);
}

/* SREC = records staged per bulk store (OUT_FULL64): 32 = a whole round of clocks, 16 = half a round
 * (half the staging memory per warp, twice the warps per SM) */
template <int OUT, int WARPS, int SREC>
__global__ void __launch_bounds__(WARPS * 32, OUT == OUT_FULL64 ? 1 : 3) decode_kernel(const dec_args a)
{
	extern __shared__ __align__(16) unsigned char smem[];
	btd_small_tables *s_small = reinterpret_cast<btd_small_tables *>(smem);
	uint16_t *s_nib = reinterpret_cast<uint16_t *>(smem + SMALL_BYTES);
	{
		const uint4 *g0 = reinterpret_cast<const uint4 *>(&a.tables->s);
		uint4 *d0 = reinterpret_cast<uint4 *>(smem);
		for (int i = threadIdx.x; i < (int)sizeof(btd_small_tables) / 16; i += blockDim.x) d0[i] = g0[i];
		const uint4 *g1 = reinterpret_cast<const uint4 *>(a.tables->nib);
		uint4 *d1 = reinterpret_cast<uint4 *>(smem + SMALL_BYTES);
		for (int i = threadIdx.x; i < NIB_BYTES / 16; i += blockDim.x) d1[i] = g1[i];
	}
	__syncthreads();

	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	btd_pkt &P = *reinterpret_cast<btd_pkt *>(smem + SMALL_BYTES + NIB_BYTES + wid * PKT_BYTES);
	constexpr int STG_BYTES = SREC * (int)sizeof(btbb_b200_decoded);
	uint32_t *stage = reinterpret_cast<uint32_t *>(smem + SMALL_BYTES + NIB_BYTES + WARPS * PKT_BYTES + wid * STG_BYTES);
	if (lane == 0) P.raw[BTD_RAW_WORDS - 1] = 0;
	btd_ctx c;
	c.s = s_small; c.nib = s_nib; c.wp = a.tables->wp;
	int64_t n = a.n;
	if (a.n_dev) n = (int64_t)*a.n_dev;
	const int mode = a.mode;

	for (int64_t it = (int64_t)blockIdx.x * WARPS + wid; it < n; it += (int64_t)gridDim.x * WARPS) {
		const int64_t p = a.idx ? a.idx[it] : it;
		const btbb_b200_pkt_in in = a.pkts[p];
		int length = in.length;
		if (length > BT_MAX_SYMBOLS) length = BT_MAX_SYMBOLS;       /* btbb_packet_set_data :472 */
		if (in.offset < 0 || in.offset >= a.stream_len) length = 0;
		else if (in.offset + length > a.stream_len) length = (int)(a.stream_len - in.offset);
		if (length < 0) length = 0;
		const uintptr_t addr = reinterpret_cast<uintptr_t>(a.stream) + (uintptr_t)(length > 0 ? in.offset : 0);
		const int sh = (int)(addr & 15);
		const uint4 *base = reinterpret_cast<const uint4 *>(addr - (uintptr_t)sh);
		c.whitened = in.whitened;
		__syncwarp();
		if (lane == 0) { P.sh = sh; P.length = length; }
		/* ---- header part: 512 symbols from the aligned base ---- */
		ingest<1>(P, base, sh, length, 0, 16, lane);
		__syncwarp();
		uint32_t hdr, hdr_ok;
		{
			uint32_t bit = 0, bad = 0;
			if (lane < 18) btd_vote3(P, 68, lane, &bit, &bad);
			hdr = __ballot_sync(FULL, bit) & 0x3ffffu;
			hdr_ok = __popc(__ballot_sync(FULL, bad)) < 18 / 4;      /* unfec13: be < length / 4 (:563-567) */
			if (lane == 0) {
				P.hdr = hdr; P.hdr_ok = (int)hdr_ok;
				P.f8[0] = P.f8[1] = btd_bits(P.raw, sh + 122, 8) * 0x01010101u;
			}
		}
		__syncwarp();
		/* ---- which packet types are in play ---- */
		btd_lane s;
		uint32_t type_mask, hp = 0;
		int header_ok = (int)hdr_ok;
		if (OUT == OUT_ONE) {
			btd_lane_init(s, (int)(in.clkn & 63));
			if (mode == BTBB_B200_MODE_DECODE) {
				header_ok = btd_decode_header(c, P, s, in.uap, &hp);
				type_mask = header_ok ? 1u << s.type : 0u;
			} else {
				s.uap = in.uap; s.type = in.type & 15;
				type_mask = btd_kind_type_mask(mode == BTBB_B200_MODE_PAYLOAD ? BTD_KIND_PAYLOAD
							       : mode == BTBB_B200_MODE_CRC_CHECK ? BTD_KIND_CRC_CHECK
							       : BTD_KIND_RAW + (mode - BTBB_B200_MODE_RAW), s.type);
			}
		} else {
			/* packet type under clock candidates lane and lane + 32 (try_clock's type field alone) */
			uint32_t ty0 = 0, ty1 = 0;
			if (hdr_ok) {
				ty0 = ((hdr ^ btd_white(c, lane, 0, 18)) >> 3) & 15u;
				ty1 = ((hdr ^ btd_white(c, lane + 32, 0, 18)) >> 3) & 15u;
			}
			type_mask = __reduce_or_sync(FULL, (1u << ty0) | (1u << ty1));
		}
		btd_needs nd;
		btd_needs_for(type_mask, length, 0, &nd);
		/* ---- the rest of the symbols that can matter ---- */
		{
			int w_end = (sh + nd.symbols + 31) / 32 + 1;
			w_end = (w_end + 15) & ~15;
			if (w_end > BTD_RAW_WORDS - 1) w_end = BTD_RAW_WORDS - 1;
			ingest<6>(P, base, sh, length, 16, w_end, lane);
		}
		__syncwarp();
		/* ---- clock-independent work ---- */
		if (nd.hv1) {
			int nbad = 0;
			for (int r = 0; r < 3; r++) {
				const int i = r * 32 + lane;
				uint32_t bit = 0, bad = 0;
				if (i < 80) btd_vote3(P, 122, i, &bit, &bad);
				const uint32_t v = __ballot_sync(FULL, bit);
				nbad += __popc(__ballot_sync(FULL, bad));
				if (lane == 0) P.hv1[r] = v;
			}
			if (lane == 0) { P.hv1[3] = 0; P.hv1_ok = nbad < 80 / 4; }
		}
		{
			const int f0 = fec_blocks(P, P.fec0, 122, nd.nblk0, s_small->fec_col, lane);
			const int f80 = nd.nblk80 ? fec_blocks(P, P.fec80, 202, nd.nblk80, s_small->fec_col, lane) : 1 << 20;
			if (lane == 0) { P.fail0 = f0; P.fail80 = f80; }
		}
		__syncwarp();
		build_dp(P, s_nib, BTD_SRC_RAW, nd.raw_bytes, P.dp_raw, lane);
		build_dp(P, s_nib, BTD_SRC_FEC0, nd.fec0_bytes, P.dp_fec0, lane);
		build_dp(P, s_nib, BTD_SRC_FEC80, nd.fec80_bytes, P.dp_fec80, lane);
		build_dp(P, s_nib, BTD_SRC_FIRST8, nd.first8_bytes, P.dp_first8, lane);
		__syncwarp();

		if (OUT == OUT_ONE) {
			/* every lane carries the same (clock, UAP, type): searches run on all 32 lanes */
			if (mode == BTBB_B200_MODE_DECODE) {
				if (header_ok) evaluate<true>(c, P, s, BTD_KIND_PAYLOAD, lane);
			} else {
				const int kind = mode == BTBB_B200_MODE_PAYLOAD ? BTD_KIND_PAYLOAD
					       : mode == BTBB_B200_MODE_CRC_CHECK ? BTD_KIND_CRC_CHECK
					       : BTD_KIND_RAW + (mode - BTBB_B200_MODE_RAW);
				evaluate<true>(c, P, s, kind, lane);
			}
			const int nbits = btd_emit_bits(s, a.raw_payload);
			const int q18 = btd_q18(c, s.pay_clk);
			const uint32_t *sb;
			int pos0, step;
			btd_src_desc(P, s.src, &sb, &pos0, &step);
			uint32_t *o = reinterpret_cast<uint32_t *>(&a.out[p]);
			#pragma unroll
			for (int i = 0; i < 3; i++) {
				const int w = lane + 32 * i;
				if (w < REC_WORDS) {
					uint32_t v;
					if (w < 7) v = btd_record_word(s, header_ok, hp, w);
					else v = btd_pay_word(c, sb, pos0 + step * (w - 7), (q18 + 32 * (w - 7)) % 127, nbits - 32 * (w - 7));
					o[w] = v;
				}
			}
		} else {
			#pragma unroll 1
			for (int round = 0; round < 2; round++) {
				btd_lane_init(s, lane + 32 * round);
				btd_try_clock(c, P, s);
				evaluate<false>(c, P, s, BTD_KIND_CRC_CHECK, lane);
				if (OUT == OUT_TC16) {
					a.tc16[p * 64 + s.clock] = (uint16_t)btd_tc16(s);
				} else {
					#pragma unroll 1
					for (int part = 0; part < 32 / SREC; part++) {
						/* the previous bulk store has to be done reading the staging buffer */
						if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
						__syncwarp();
						{
							uint4 *z = reinterpret_cast<uint4 *>(stage);
							for (int i = lane; i < STG_BYTES / 16; i += 32) z[i] = make_uint4(0, 0, 0, 0);
						}
						__syncwarp();
						if (SREC == 32 || (lane / SREC) == part) {
							uint32_t *rec = stage + (lane % SREC) * REC_WORDS;      /* 93 words apart: conflict-free */
							#pragma unroll
							for (int w = 0; w < 7; w++) rec[w] = btd_record_word(s, (int)hdr_ok, 0, w);
							/* payload: only the words that carry bits; the bit source is a pointer, a start bit and
							 * a stride, so every lane runs the same instructions whatever its packet type is */
							const int nbits = btd_emit_bits(s, a.raw_payload);
							if (nbits > 0) {
								const int nw = (nbits + 31) >> 5;
								int q = btd_q18(c, s.pay_clk);
								const uint32_t *sb;
								int pos, step;
								btd_src_desc(P, s.src, &sb, &pos, &step);
								const uint32_t sh32 = (uint32_t)pos & 31u, adv = (uint32_t)step >> 5;
								const uint32_t *wp = sb + (pos >> 5);
								uint32_t lo_w = wp[0];
								for (int j = 0; j < nw; j++) {
									const uint32_t hi_w = wp[1];
									uint32_t d = __funnelshift_r(lo_w, hi_w, sh32);
									if (c.whitened) d ^= s_small->wrot[q];
									if (j == nw - 1 && (nbits & 31)) d &= (1u << (nbits & 31)) - 1u;
									rec[7 + j] = d;
									if (adv) lo_w = hi_w;
									wp += adv;
									q += 32; if (q >= 127) q -= 127;
								}
							}
						}
						unsigned char *dst = reinterpret_cast<unsigned char *>(&a.out[p * 64 + 32 * round + SREC * part]);
						if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
							asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
							__syncwarp();
							if (lane == 0) {
								const uint32_t src = (uint32_t)__cvta_generic_to_shared(stage);
								asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
									     :: "l"(dst), "r"(src), "r"(STG_BYTES) : "memory");
								asm volatile("cp.async.bulk.commit_group;" ::: "memory");
							}
						} else {
							__syncwarp();
							uint32_t *o = reinterpret_cast<uint32_t *>(dst);
							for (int w = lane; w < SREC * REC_WORDS; w += 32) o[w] = stage[w];
						}
						__syncwarp();
					}
				}
			}
		}
		__syncwarp();
	}
	if (OUT == OUT_FULL64 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void header_present_kernel(const uint8_t *stream, int64_t stream_len,
				      const btbb_b200_pkt_in *pkts, int64_t n, uint8_t *present)
{
	int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
	if (p >= n) return;
	const btbb_b200_pkt_in in = pkts[p];
	int length = in.length > BT_MAX_SYMBOLS ? BT_MAX_SYMBOLS : in.length;
	if (in.offset < 0 || in.offset >= stream_len) length = 0;
	else if (in.offset + length > stream_len) length = (int)(stream_len - in.offset);
	if (length < 122) { present[p] = 0; return; }     /* :1380 */
	const uint8_t *s = stream + in.offset;
	int msb = s[63] & 1, be = 0;
	for (int i = 0; i < 4; i++)                        /* trailer :1384-1388 */
		be += (s[64 + i] & 1) ^ ((i & 1) ? msb : !msb);
	for (int i = 0; i < 18; i++) {                     /* triplets :1395-1401 */
		int t = (s[68 + 3 * i] & 1) + (s[69 + 3 * i] & 1) + (s[70 + 3 * i] & 1);
		be += (t == 1 || t == 2);
	}
	present[p] = be < 5;                               /* ID_THRESHOLD, bluetooth_packet.h:33 */
}

int ensure_dec_tables(btbb_b200_ctx *ctx)
{
	if (ctx->d_dec_tables) return BTBB_B200_OK;
	const btd_tables *t = btd_host_tables();
	void *d = NULL;
	BT_CUDA_TRY(cudaMalloc(&d, sizeof(btd_tables)));
	cudaError_t e = cudaMemcpy(d, t, sizeof(btd_tables), cudaMemcpyHostToDevice);
	if (e != cudaSuccess) { cudaFree(d); return btbb_b200_cuda_fail(e, "cudaMemcpy(decode tables)"); }
	ctx->d_dec_tables = d;
	return BTBB_B200_OK;
}

template <int OUT, int WARPS, int SREC = 32>
int launch(btbb_b200_ctx *ctx, const dec_args &a, int64_t n_max, cudaStream_t st)
{
	auto kern = decode_kernel<OUT, WARPS, SREC>;
	constexpr int smem = smem_bytes<OUT, WARPS, SREC>();
	static bool configured[16];
	const int dev = ctx->device;
	if (dev < 0 || dev >= 16 || !configured[dev]) {
		BT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
		if (dev >= 0 && dev < 16) configured[dev] = true;
	}
	int per_sm = 1;
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
	int64_t blocks = (n_max + WARPS - 1) / WARPS;
	const int64_t cap = (int64_t)ctx->sm_count * per_sm;
	if (blocks > cap) blocks = cap;
	kern<<<(unsigned)blocks, WARPS * 32, smem, st>>>(a);
	BT_CUDA_TRY(cudaGetLastError());
	return BTBB_B200_OK;
}

bool mode_ok(int mode)
{
	const int m = mode & ~BTBB_B200_MODE_FLAG_RAW_PAYLOAD;
	return m >= 0 && (m <= 3 || (m >= 16 && m <= 22));
}

}  // namespace

extern "C" int btbb_b200_decode_dev(btbb_b200_ctx *ctx, const uint8_t *d_stream, int64_t stream_length,
				    const btbb_b200_pkt_in *d_pkts, int64_t n, int mode,
				    btbb_b200_decoded *d_out, void *cuda_stream)
{
	if (!ctx || n < 0 || (n > 0 && (!d_stream || !d_pkts || !d_out)) || stream_length < 0 || !mode_ok(mode) ||
	    (reinterpret_cast<uintptr_t>(d_out) & 3))
		return btbb_b200_set_error(BTBB_B200_EINVAL, "decode: bad arguments");
	BT_CUDA_TRY(cudaSetDevice(ctx->device));
	int rc = ensure_dec_tables(ctx);
	if (rc) return rc;
	if (n == 0) return BTBB_B200_OK;
	dec_args a;
	memset(&a, 0, sizeof(a));
	a.stream = d_stream; a.stream_len = stream_length; a.pkts = d_pkts; a.n = n;
	a.raw_payload = (mode & BTBB_B200_MODE_FLAG_RAW_PAYLOAD) != 0;
	a.mode = mode & ~BTBB_B200_MODE_FLAG_RAW_PAYLOAD;
	a.out = d_out; a.tables = static_cast<const btd_tables *>(ctx->d_dec_tables);
	if (a.mode == BTBB_B200_MODE_TRY_CLOCKS) {
		/* 24 warps per SM staging half a round of records each; BTBB_B200_OPT_DECODE_WIDE_STAGING: 12 warps, whole rounds */
		if (ctx->opt_decode_wide) return launch<OUT_FULL64, 12, 32>(ctx, a, n, (cudaStream_t)cuda_stream);
		return launch<OUT_FULL64, 24, 16>(ctx, a, n, (cudaStream_t)cuda_stream);
	}
	return launch<OUT_ONE, 8>(ctx, a, n, (cudaStream_t)cuda_stream);
}

/* try_clock + crc_check for CLK1-6 = 0..63 (bluetooth_piconet.c:675-689) of the packets listed in
 * d_idx[0 .. *d_n) (at most n_max of them), results as one 16-bit word per (packet, clock) at
 * d_tc[64 * packet + clock]: the UAP try_clock returned in the low byte, the crc_check class above
 * it (0, 1, 2 as returned; 3 = 10; 4 = 1000) */
int bt_try_clocks_compact(btbb_b200_ctx *ctx, const uint8_t *d_stream, int64_t stream_length,
			  const btbb_b200_pkt_in *d_pkts, const int64_t *d_idx, const unsigned long long *d_n,
			  int64_t n_max, uint16_t *d_tc, cudaStream_t st)
{
	int rc = ensure_dec_tables(ctx);
	if (rc) return rc;
	if (n_max == 0) return BTBB_B200_OK;
	dec_args a;
	memset(&a, 0, sizeof(a));
	a.stream = d_stream; a.stream_len = stream_length; a.pkts = d_pkts; a.n = n_max;
	a.mode = BTBB_B200_MODE_TRY_CLOCKS; a.tc16 = d_tc; a.idx = d_idx; a.n_dev = d_n;
	a.tables = static_cast<const btd_tables *>(ctx->d_dec_tables);
	return launch<OUT_TC16, 8>(ctx, a, n_max, st);
}

extern "C" int btbb_b200_try_clocks_compact_dev(btbb_b200_ctx *ctx, const uint8_t *d_stream, int64_t stream_length,
						const btbb_b200_pkt_in *d_pkts, int64_t n, uint16_t *d_tc, void *cuda_stream)
{
	if (!ctx || n < 0 || (n > 0 && (!d_stream || !d_pkts || !d_tc)) || stream_length < 0)
		return btbb_b200_set_error(BTBB_B200_EINVAL, "try_clocks_compact: bad arguments");
	BT_CUDA_TRY(cudaSetDevice(ctx->device));
	return bt_try_clocks_compact(ctx, d_stream, stream_length, d_pkts, NULL, NULL, n, d_tc, (cudaStream_t)cuda_stream);
}

extern "C" int btbb_b200_header_present_dev(btbb_b200_ctx *ctx, const uint8_t *d_stream, int64_t stream_length,
					    const btbb_b200_pkt_in *d_pkts, int64_t n, uint8_t *d_present,
					    void *cuda_stream)
{
	if (!ctx || n < 0 || (n > 0 && (!d_stream || !d_pkts || !d_present)) || stream_length < 0)
		return btbb_b200_set_error(BTBB_B200_EINVAL, "header_present: bad arguments");
	BT_CUDA_TRY(cudaSetDevice(ctx->device));
	if (n == 0) return BTBB_B200_OK;
	header_present_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)cuda_stream>>>(d_stream, stream_length, d_pkts, n, d_present);
	BT_CUDA_TRY(cudaGetLastError());
	return BTBB_B200_OK;
}

/* grow-only device scratch of the host-buffer entry points (no cudaMalloc per call) */
static int ensure_scratch(btbb_b200_ctx *ctx, int which, size_t bytes, void **out)
{
	if (bytes > ctx->scratch_cap[which]) {
		if (ctx->d_scratch[which]) cudaFree(ctx->d_scratch[which]);
		ctx->d_scratch[which] = NULL; ctx->scratch_cap[which] = 0;
		size_t cap = bytes < 65536 ? 65536 : bytes + bytes / 4;
		BT_CUDA_TRY(cudaMalloc(&ctx->d_scratch[which], cap));
		ctx->scratch_cap[which] = cap;
	}
	*out = ctx->d_scratch[which];
	return BTBB_B200_OK;
}

extern "C" int btbb_b200_decode_host(btbb_b200_ctx *ctx, const char *stream, int64_t stream_length,
				     const btbb_b200_pkt_in *pkts, int64_t n, int mode, btbb_b200_decoded *out)
{
	if (!ctx || n < 0 || (n > 0 && (!stream || !pkts || !out)) || stream_length < 0 || !mode_ok(mode))
		return btbb_b200_set_error(BTBB_B200_EINVAL, "decode_host: bad arguments");
	if (n == 0) return BTBB_B200_OK;
	BT_CUDA_TRY(cudaSetDevice(ctx->device));
	std::lock_guard<std::mutex> guard(*ctx->host_lock);
	void *d_s = NULL, *d_p = NULL, *d_o = NULL;
	const int64_t nout = (mode & ~BTBB_B200_MODE_FLAG_RAW_PAYLOAD) == BTBB_B200_MODE_TRY_CLOCKS ? n * 64 : n;
	int rc = ensure_scratch(ctx, 0, (size_t)stream_length + 64, &d_s);
	if (!rc) rc = ensure_scratch(ctx, 1, (size_t)n * sizeof(btbb_b200_pkt_in), &d_p);
	if (!rc) rc = ensure_scratch(ctx, 2, (size_t)nout * sizeof(btbb_b200_decoded), &d_o);
	if (rc) return rc;
	cudaError_t e;
	if ((e = cudaMemcpy(d_s, stream, (size_t)stream_length, cudaMemcpyHostToDevice)) != cudaSuccess ||
	    (e = cudaMemcpy(d_p, pkts, (size_t)n * sizeof(btbb_b200_pkt_in), cudaMemcpyHostToDevice)) != cudaSuccess)
		return btbb_b200_cuda_fail(e, "cudaMemcpy(decode_host H2D)");
	rc = btbb_b200_decode_dev(ctx, static_cast<const uint8_t *>(d_s), stream_length, static_cast<const btbb_b200_pkt_in *>(d_p),
				  n, mode, static_cast<btbb_b200_decoded *>(d_o), NULL);
	if (rc) return rc;
	if ((e = cudaMemcpy(out, d_o, (size_t)nout * sizeof(btbb_b200_decoded), cudaMemcpyDeviceToHost)) != cudaSuccess)
		return btbb_b200_cuda_fail(e, "cudaMemcpy(decode_host D2H)");
	return BTBB_B200_OK;
}

extern "C" int btbb_b200_header_present_host(btbb_b200_ctx *ctx, const char *stream, int64_t stream_length,
					     const btbb_b200_pkt_in *pkts, int64_t n, uint8_t *present)
{
	if (!ctx || n < 0 || (n > 0 && (!stream || !pkts || !present)) || stream_length < 0)
		return btbb_b200_set_error(BTBB_B200_EINVAL, "header_present_host: bad arguments");
	if (n == 0) return BTBB_B200_OK;
	BT_CUDA_TRY(cudaSetDevice(ctx->device));
	std::lock_guard<std::mutex> guard(*ctx->host_lock);
	void *d_s = NULL, *d_p = NULL, *d_o = NULL;
	int rc = ensure_scratch(ctx, 0, (size_t)stream_length + 64, &d_s);
	if (!rc) rc = ensure_scratch(ctx, 1, (size_t)n * sizeof(btbb_b200_pkt_in), &d_p);
	if (!rc) rc = ensure_scratch(ctx, 2, (size_t)n, &d_o);
	if (rc) return rc;
	BT_CUDA_TRY(cudaMemcpy(d_s, stream, (size_t)stream_length, cudaMemcpyHostToDevice));
	BT_CUDA_TRY(cudaMemcpy(d_p, pkts, (size_t)n * sizeof(btbb_b200_pkt_in), cudaMemcpyHostToDevice));
	rc = btbb_b200_header_present_dev(ctx, static_cast<const uint8_t *>(d_s), stream_length, static_cast<const btbb_b200_pkt_in *>(d_p), n,
					  static_cast<uint8_t *>(d_o), NULL);
	if (rc) return rc;
	BT_CUDA_TRY(cudaMemcpy(present, d_o, (size_t)n, cudaMemcpyDeviceToHost));
	return BTBB_B200_OK;
}

extern "C" int btbb_b200_decode_smallcall(const char *symbols, int length, uint32_t clkn, uint8_t uap,
					  int whitened, uint8_t type, int mode, btbb_b200_decoded *out)
{
	if (!symbols || !out || !mode_ok(mode))
		return btbb_b200_set_error(BTBB_B200_EINVAL, "decode_smallcall: bad arguments");
	return bt_decode_one_cpu(symbols, length, clkn, uap, whitened, type, mode, out);
}
