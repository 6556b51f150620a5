/*
 * decode.cu -- K4/K5: the per-packet chain  unfec13 -> unwhiten -> HEC  and
 * unfec23 / unwhiten / CRC  (bluetooth_packet.c:552-705, 708-1317) for a batch of
 * detected packets, one warp per packet.
 *
 * Everything that does not depend on the clock candidate is done once per packet by the
 * whole warp and parked in shared memory as packed bits:
 *   - the packet's symbols (ballot-packed, symbols past `length` read 0 like a fresh
 *     btbb_packet),
 *   - the FEC-1/3 header vote (unfec13 :552-568) and the HV1 payload vote,
 *   - every FEC-2/3 block from symbol 122 on (and from 202 on for DV, :914), corrected,
 *     with the index of the first uncorrectable block (unfec23 :585-649).
 * The clock-dependent tail (dewhiten :653-668, uap_from_hec :693-705, payload header
 * :821-895, CRC :671-690 byte-wise by table) then runs with one lane per CLK1-6 candidate
 * in mode 1 (try_clock + crc_check, :1178-1195 / :708-769) or on lane 0 in mode 0
 * (btbb_decode_header + btbb_decode_payload, :1198-1297).
 */
#include <cuda_runtime.h>
#include <string.h>
#include "bt_math.h"
#include "capi_internal.h"

namespace {

constexpr int WARPS = 4;
constexpr int RAW_WORDS = 100;      /* 3200 >= 3125 symbols */
constexpr int FEC_BLOCKS = 200;     /* (3125 - 122) / 15 */
constexpr int FEC_WORDS = 64;       /* 200 * 10 bits */

struct dec_tables {
	uint32_t wseq[13];     /* whitening m-sequence from state 0x40, three periods */
	uint8_t phase[64];     /* position in wseq where the LFSR state is 0x40 | clk */
	uint16_t crc[256];     /* reflected CRC-16/CCITT byte table */
	uint8_t fec_col[10];   /* parity column of each data bit */
	uint8_t pad[2];
};
__constant__ dec_tables c_dec;

struct warp_smem {
	uint32_t raw[RAW_WORDS + 1];
	uint32_t fec0[FEC_WORDS + 1];   /* corrected data bits, blocks from symbol 122 */
	uint32_t fec80[FEC_WORDS + 1];  /* ... from symbol 202 (DV) */
	uint32_t hv1[4];                /* FEC-1/3 vote of the 240 symbols at 122 */
	uint32_t hdr;                   /* 18 voted header bits */
	int hdr_ok, hv1_ok, fail0, fail80, length;
};

struct lane_state {
	uint32_t uap, type, lt_addr, flags, hec, llid, flow, has_payload;
	int phl, plen;
	int src, pay_clk;    /* where the payload bytes come from, for the final emit */
};
enum { SRC_NONE = 0, SRC_FEC0, SRC_FEC80, SRC_RAW, SRC_HV1, SRC_FIRST8 };

__device__ __forceinline__ uint32_t bits_at(const uint32_t *w, int pos, int n)
{
	uint32_t v = __funnelshift_r(w[pos >> 5], w[(pos >> 5) + 1], pos & 31);
	return n >= 32 ? v : v & ((1u << n) - 1);
}

__device__ __forceinline__ uint32_t whiten_bits(const uint32_t *s_wseq, const uint8_t *s_phase,
						int clk, int pos, int n, int whitened)
{
	if (!whitened) return 0;
	int p = (s_phase[clk & 63] + pos) % 127;
	return bits_at(s_wseq, p, n);
}

struct dec_ctx {
	const warp_smem *ws;
	const uint32_t *wseq;
	const uint8_t *phase;
	const uint16_t *crc;
	int whitened;
};

__device__ uint32_t pay_byte(const dec_ctx &d, int src, int clk, int i)
{
	uint32_t b;
	switch (src) {
	case SRC_FEC0:  b = bits_at(d.ws->fec0, 8 * i, 8); break;
	case SRC_FEC80: b = bits_at(d.ws->fec80, 8 * i, 8); break;
	case SRC_RAW:   b = bits_at(d.ws->raw, 122 + 8 * i, 8); break;
	case SRC_HV1:   b = bits_at(d.ws->hv1, 8 * i, 8); break;
	case SRC_FIRST8: b = bits_at(d.ws->raw, 122, 8); break;   /* EV3/EV5 quirk, :1036 */
	default: b = 0;
	}
	return b ^ whiten_bits(d.wseq, d.phase, clk, 18 + 8 * i, 8, d.whitened);
}

__device__ __forceinline__ uint32_t crc_byte(const dec_ctx &d, uint32_t reg, uint32_t byte)
{
	return (reg >> 8) ^ d.crc[(reg ^ byte) & 0xff];
}

/* payload_crc (:772-781) for plen >= 2 */
__device__ bool crc_matches(const dec_ctx &d, int src, int clk, int plen, uint32_t uap)
{
	uint32_t reg = bt_crc16_init(uap);
	for (int i = 0; i < plen - 2; i++)
		reg = crc_byte(d, reg, pay_byte(d, src, clk, i));
	uint32_t chk = pay_byte(d, src, clk, plen - 2) | (pay_byte(d, src, clk, plen - 1) << 8);
	return reg == chk;
}

/* fhs (:783-818) */
__device__ int dec_fhs(const dec_ctx &d, lane_state &s, int clock)
{
	int size = d.ws->length - 122;
	s.plen = 20;
	if (size < 240) return 1;
	if (d.ws->fail0 < 16) return 0;
	s.src = SRC_FEC0; s.pay_clk = clock;
	if (crc_matches(d, SRC_FEC0, clock, 20, s.uap)) return 1000;
	for (int c = 32; c < 64; c++) {
		s.pay_clk = c;
		if (crc_matches(d, SRC_FEC0, c, 20, s.uap)) return 1000;
	}
	return 0;
}

/* decode_payload_header (:821-895); start80 selects the DV block alignment */
__device__ int dec_pay_hdr(const dec_ctx &d, lane_state &s, int clock, int hbytes, int size, int fec, bool dv)
{
	int nb = hbytes * 8;
	if (size < nb) return 0;
	uint32_t ph;
	if (fec) {
		if (size < (hbytes == 2 ? 30 : 15)) return 0;
		int need = hbytes == 2 ? 2 : 1;
		if ((dv ? d.ws->fail80 : d.ws->fail0) < need) return 0;
		ph = bits_at(dv ? d.ws->fec80 : d.ws->fec0, 0, nb);
	} else
		ph = bits_at(d.ws->raw, 122, nb);
	ph ^= whiten_bits(d.wseq, d.phase, clock, 18, nb, d.whitened);
	s.plen = hbytes == 2 ? (int)((ph >> 3) & 0x3ff) + 4 : (int)((ph >> 3) & 0x1f) + 3;
	int maxlen;
	switch (s.type) {
	case 3: maxlen = 20; break;
	case 4: maxlen = 30; break;
	case 8: maxlen = 12; break;
	case 10: maxlen = 125; break;
	case 11: maxlen = 187; break;
	case 14: maxlen = 228; break;
	case 15: maxlen = 343; break;
	default: maxlen = 0;
	}
	if (s.plen > maxlen) s.plen = maxlen;
	s.llid = ph & 3;
	s.flow = (ph >> 2) & 1;
	s.phl = hbytes;
	return 1;
}

/* DM (:898-958) */
__device__ int dec_dm(const dec_ctx &d, lane_state &s, int clock)
{
	int size = d.ws->length - 122, hbytes = 2, maxlen;
	bool dv = false;
	switch (s.type) {
	case 8: dv = true; size -= 80; hbytes = 1; maxlen = 12; break;
	case 3: hbytes = 1; maxlen = 20; break;
	case 10: maxlen = 125; break;
	case 14: maxlen = 228; break;
	default: return 0;
	}
	if (!dec_pay_hdr(d, s, clock, hbytes, size, 1, dv)) return 0;
	if (s.plen > maxlen) return 1;
	int nbits = s.plen * 8;
	if (nbits > size) return 1;
	if ((dv ? d.ws->fail80 : d.ws->fail0) < (nbits + 9) / 10) return 0;
	s.src = dv ? SRC_FEC80 : SRC_FEC0; s.pay_clk = clock;
	return crc_matches(d, s.src, clock, s.plen, s.uap) ? 10 : 2;
}

/* DH (:962-1011) */
__device__ int dec_dh(const dec_ctx &d, lane_state &s, int clock)
{
	int size = d.ws->length - 122, hbytes = 2, maxlen;
	switch (s.type) {
	case 9: case 4: hbytes = 1; maxlen = 30; break;
	case 11: maxlen = 187; break;
	case 15: maxlen = 343; break;
	default: return 0;
	}
	if (!dec_pay_hdr(d, s, clock, hbytes, size, 0, false)) return 0;
	if (s.plen > maxlen) return 1;
	int nbits = s.plen * 8;
	if (nbits > size) return 1;
	s.src = SRC_RAW; s.pay_clk = clock;
	if (s.type == 9) return 2;
	return crc_matches(d, SRC_RAW, clock, s.plen, s.uap) ? 10 : 2;
}

/* EV3 (:1013-1042) / EV5 (:1099-1128) with an incremental CRC */
__device__ int dec_ev35(const dec_ctx &d, lane_state &s, int clock, int maxlength)
{
	int size = d.ws->length - 122;
	uint32_t reg = bt_crc16_init(s.uap);   /* CRC over bytes [0, plen-2) */
	s.src = SRC_FIRST8; s.pay_clk = clock;
	for (s.plen = 0; s.plen < maxlength; s.plen++) {
		if (s.plen * 8 + 8 > size) return 1;
		if (s.plen > 2) {
			reg = crc_byte(d, reg, pay_byte(d, SRC_FIRST8, clock, s.plen - 3));
			uint32_t chk = pay_byte(d, SRC_FIRST8, clock, s.plen - 2) |
				       (pay_byte(d, SRC_FIRST8, clock, s.plen - 1) << 8);
			if (reg == chk) return 10;
		}
	}
	return 2;
}

/* EV4 (:1044-1097) with an incremental CRC */
__device__ int dec_ev4(const dec_ctx &d, lane_state &s, int clock)
{
	int size = d.ws->length - 122, syms = 0, bits = 0, blk = 0;
	uint32_t reg = bt_crc16_init(s.uap);   /* CRC over bytes [0, plen-2) once plen >= 2 */
	s.plen = 1;
	s.src = SRC_FEC0; s.pay_clk = clock;
	while (syms < 1470) {
		if (syms + 15 > size) return 1;
		if (d.ws->fail0 <= blk) return syms < 45 ? 0 : 1;
		while (s.plen * 8 <= bits) {
			if (s.plen >= 2) {
				uint32_t chk = pay_byte(d, SRC_FEC0, clock, s.plen - 2) |
					       (pay_byte(d, SRC_FEC0, clock, s.plen - 1) << 8);
				if (reg == chk) return 10;
				reg = crc_byte(d, reg, pay_byte(d, SRC_FEC0, clock, s.plen - 2));
			}
			/* plen == 1: the reference compares the CRC preload (low byte 0) with a word
			 * whose bit 4 is the packet's own payload_length byte (=1): never equal */
			s.plen++;
		}
		syms += 15; bits += 10; blk++;
	}
	return 2;
}

/* HV (:1131-1174) */
__device__ int dec_hv(const dec_ctx &d, lane_state &s, int clock)
{
	int size = d.ws->length - 122;
	s.phl = 0;
	if (size < 240) { s.plen = 0; return 1; }
	switch (s.type) {
	case 5:
		if (!d.ws->hv1_ok) return 0;
		s.plen = 10; s.has_payload = 1; s.src = SRC_HV1; s.pay_clk = clock;
		break;
	case 6:
		if (d.ws->fail0 < 16) return 0;
		s.plen = 20; s.has_payload = 1; s.src = SRC_FEC0; s.pay_clk = clock;
		break;
	case 7:
		s.plen = 30; s.has_payload = 1; s.src = SRC_RAW; s.pay_clk = clock;
		break;
	}
	return 2;
}

/* crc_check (:708-769) */
__device__ int do_crc_check(const dec_ctx &d, lane_state &s, int clock)
{
	int rv = 1;
	switch (s.type) {
	case 2: rv = dec_fhs(d, s, clock); break;
	case 8: case 3: case 10: case 14: rv = dec_dm(d, s, clock); break;
	case 4: case 11: case 15: rv = dec_dh(d, s, clock); break;
	case 7: rv = dec_ev35(d, s, clock, 32); break;
	case 12: rv = dec_ev4(d, s, clock); break;
	case 13: rv = dec_ev35(d, s, clock, 182); break;
	case 5: rv = dec_hv(d, s, clock); break;
	default: break;
	}
	if (rv == 0 && s.type != 2 && s.type != 3 && s.type != 5) return 1;
	if (rv > 1 && (s.type == 7 || s.type == 13)) return 1;
	return rv;
}

/* btbb_decode_payload (:1223-1297) */
__device__ int do_decode_payload(const dec_ctx &d, lane_state &s, int clock)
{
	int rv = 0;
	s.phl = 0;
	switch (s.type) {
	case 0: case 1: s.plen = 0; rv = 1; break;
	case 2: rv = dec_fhs(d, s, clock); break;
	case 3: case 8: case 10: case 14: rv = dec_dm(d, s, clock); break;
	case 4: case 9: case 11: case 15: rv = dec_dh(d, s, clock); break;
	case 5: case 6: rv = dec_hv(d, s, clock); break;
	case 7:
		rv = dec_ev35(d, s, clock, 32);
		if (rv <= 1) rv = dec_hv(d, s, clock);
		break;
	case 12: rv = dec_ev4(d, s, clock); break;
	case 13: rv = dec_ev35(d, s, clock, 182); break;
	}
	s.has_payload = 1;
	return rv;
}

__device__ void emit_record(const dec_ctx &d, const lane_state &s, int header_ok, int rv,
			    uint32_t header_packed, btbb_b200_decoded *o)
{
	o->header_ok = header_ok; o->rv = rv;
	o->uap = (uint8_t)s.uap; o->type = (uint8_t)s.type; o->lt_addr = (uint8_t)s.lt_addr;
	o->flags = (uint8_t)s.flags; o->hec = (uint8_t)s.hec; o->llid = (uint8_t)s.llid;
	o->flow = (uint8_t)s.flow; o->has_payload = (uint8_t)s.has_payload;
	o->payload_header_length = s.phl; o->payload_length = s.plen;
	o->header_packed = header_packed;
	int n = (rv >= 2 && s.plen > 0 && s.plen <= 344) ? s.plen : 0;
	for (int i = 0; i < n; i++)
		o->payload[i] = (uint8_t)pay_byte(d, s.src, s.pay_clk, i);
	for (int i = n; i < 344; i++)
		o->payload[i] = 0;
}

/* corrected data bits of one (15,10) block; false when the reference would give up */
__device__ __forceinline__ bool fec23_block(uint32_t cw15, const uint8_t *col, uint32_t *data10)
{
	uint32_t data = cw15 & 0x3ff, diff = (cw15 >> 10) & 0x1f;
	#pragma unroll
	for (int i = 0; i < 10; i++)
		if ((data >> i) & 1) diff ^= col[i];
	if (diff & (diff - 1)) {
		bool fixed = false;
		#pragma unroll
		for (int i = 0; i < 10; i++)
			if (col[i] == diff) { data ^= 1u << i; fixed = true; }
		if (!fixed) return false;
	}
	*data10 = data;
	return true;
}

__global__ void __launch_bounds__(WARPS * 32) decode_kernel(const uint8_t *stream, int64_t stream_len,
							     const btbb_b200_pkt_in *pkts, int64_t n, int mode,
							     btbb_b200_decoded *out, uint16_t *tc16,
							     const int64_t *idx, const unsigned long long *n_dev)
{
	__shared__ warp_smem s_w[WARPS];
	__shared__ uint32_t s_wseq[13];
	__shared__ uint8_t s_phase[64];
	__shared__ uint16_t s_crc[256];
	__shared__ uint8_t s_col[16];
	for (int i = threadIdx.x; i < 13; i += blockDim.x) s_wseq[i] = c_dec.wseq[i];
	for (int i = threadIdx.x; i < 64; i += blockDim.x) s_phase[i] = c_dec.phase[i];
	for (int i = threadIdx.x; i < 256; i += blockDim.x) s_crc[i] = c_dec.crc[i];
	if (threadIdx.x < 10) s_col[threadIdx.x] = c_dec.fec_col[threadIdx.x];
	__syncthreads();

	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	warp_smem &ws = s_w[wid];
	/* idx / n_dev (the UAP sieve's rounds, sieve.cu): work item i is packet idx[i], and the number
	 * of items is read from device memory */
	if (n_dev) n = (int64_t)*n_dev;
	for (int64_t it = (int64_t)blockIdx.x * WARPS + wid; it < n; it += (int64_t)gridDim.x * WARPS) {
		const int64_t p = idx ? idx[it] : it;
		const btbb_b200_pkt_in in = pkts[p];
		int length = in.length;
		if (length > BT_MAX_SYMBOLS) length = BT_MAX_SYMBOLS;       /* btbb_packet_set_data :472 */
		if (in.offset < 0 || in.offset >= stream_len) length = 0;
		else if (in.offset + length > stream_len) length = (int)(stream_len - in.offset);
		if (length < 0) length = 0;
		const uint8_t *sym = stream + in.offset;
		__syncwarp();
		/* ---- pack symbols ---- */
		for (int w = 0; w <= RAW_WORDS; w++) {
			int i = w * 32 + lane;
			uint32_t b = (i < length) ? (sym[i] & 1u) : 0u;
			uint32_t word = __ballot_sync(0xffffffffu, b);
			if (lane == 0) ws.raw[w] = word;
		}
		if (lane == 0) ws.length = length;
		__syncwarp();
		/* ---- FEC 1/3 votes: header (18 triplets at 68) and HV1 payload (80 at 122) ---- */
		{
			uint32_t t = lane < 18 ? bits_at(ws.raw, 68 + 3 * lane, 3) : 0;
			uint32_t ones = __popc(t);
			uint32_t hdr = __ballot_sync(0xffffffffu, ones >= 2);
			uint32_t bad = __ballot_sync(0xffffffffu, lane < 18 && (ones == 1 || ones == 2));
			if (lane == 0) { ws.hdr = hdr & 0x3ffff; ws.hdr_ok = __popc(bad) < 18 / 4; }
			int nbad = 0;
			for (int r = 0; r < 3; r++) {
				int i = r * 32 + lane;
				uint32_t tt = i < 80 ? bits_at(ws.raw, 122 + 3 * i, 3) : 0;
				uint32_t o = __popc(tt);
				uint32_t v = __ballot_sync(0xffffffffu, o >= 2);
				nbad += __popc(__ballot_sync(0xffffffffu, i < 80 && (o == 1 || o == 2)));
				if (lane == 0) ws.hv1[r] = v;
			}
			if (lane == 0) { ws.hv1[3] = 0; ws.hv1_ok = nbad < 80 / 4; }
		}
		/* ---- FEC 2/3: all blocks, both alignments ---- */
		for (int al = 0; al < 2; al++) {
			uint32_t *dst = al ? ws.fec80 : ws.fec0;
			const int start = al ? 202 : 122;
			for (int w = lane; w <= FEC_WORDS; w += 32) dst[w] = 0;
			__syncwarp();
			int fail = 1 << 20;
			for (int b0 = 0; b0 < FEC_BLOCKS; b0 += 32) {
				int b = b0 + lane;
				bool live = b < FEC_BLOCKS && start + 15 * b + 15 <= RAW_WORDS * 32;
				uint32_t data = 0;
				bool ok = true;
				if (live) ok = fec23_block(bits_at(ws.raw, start + 15 * b, 15), s_col, &data);
				uint32_t badmask = __ballot_sync(0xffffffffu, live && !ok);
				if (badmask && fail == (1 << 20)) fail = b0 + __ffs(badmask) - 1;
				if (live && ok) {
					int pos = 10 * b;
					atomicOr(&dst[pos >> 5], data << (pos & 31));
					if ((pos & 31) > 22) atomicOr(&dst[(pos >> 5) + 1], data >> (32 - (pos & 31)));
				}
			}
			if (lane == 0) { if (al) ws.fail80 = fail; else ws.fail0 = fail; }
		}
		__syncwarp();

		dec_ctx d;
		d.ws = &ws; d.wseq = s_wseq; d.phase = s_phase; d.crc = s_crc; d.whitened = in.whitened;
		if (mode >= 2) {
			/* single-function entry points of the classic API: type/UAP/clock supplied */
			if (lane == 0) {
				lane_state s;
				memset(&s, 0, sizeof(s));
				s.uap = in.uap; s.type = in.type & 15;
				const int clock = (int)(in.clkn & 63);
				int rv;
				if (mode == BTBB_B200_MODE_PAYLOAD) rv = do_decode_payload(d, s, clock);
				else if (mode == BTBB_B200_MODE_CRC_CHECK) rv = do_crc_check(d, s, clock);
				else switch (mode - BTBB_B200_MODE_RAW) {
				case 0: rv = dec_fhs(d, s, clock); break;
				case 1: rv = dec_dm(d, s, clock); break;
				case 2: rv = dec_dh(d, s, clock); break;
				case 3: rv = dec_ev35(d, s, clock, 32); break;
				case 4: rv = dec_ev4(d, s, clock); break;
				case 5: rv = dec_ev35(d, s, clock, 182); break;
				default: rv = dec_hv(d, s, clock); break;
				}
				emit_record(d, s, ws.hdr_ok, rv, 0, &out[p]);
			}
		} else if (mode == 0) {
			if (lane == 0) {
				lane_state s;
				memset(&s, 0, sizeof(s));
				s.uap = in.uap;
				int ok = 0, rv = 0;
				uint32_t hp = 0;
				if (ws.hdr_ok) {
					hp = ws.hdr ^ whiten_bits(s_wseq, s_phase, (int)in.clkn, 0, 18, in.whitened);
					uint32_t d10 = hp & 0x3ff, hec = hp >> 10;
					if (bt_uap_from_hec(d10, hec) == in.uap) {
						s.lt_addr = hp & 7; s.type = (hp >> 3) & 15; s.flags = (hp >> 7) & 7; s.hec = hec;
						ok = 1;
					}
				}
				if (ok) rv = do_decode_payload(d, s, (int)(in.clkn & 63));
				emit_record(d, s, ok, rv, hp, &out[p]);
			}
		} else {
			for (int clock = lane; clock < 64; clock += 32) {
				lane_state s;
				memset(&s, 0, sizeof(s));
				int ok = ws.hdr_ok;
				if (ok) {
					uint32_t hp = ws.hdr ^ whiten_bits(s_wseq, s_phase, clock, 0, 18, in.whitened);
					s.uap = bt_uap_from_hec(hp & 0x3ff, hp >> 10);
					s.type = (hp >> 3) & 15;
				}
				int rv = do_crc_check(d, s, clock);
				if (tc16)      /* compact table for the UAP / CLK1-6 sieve (sieve.cu): UAP | class << 8 */
					tc16[p * 64 + clock] = (uint16_t)(s.uap | ((rv == 0 ? 0 : rv == 1 ? 1 : rv == 2 ? 2 : rv == 10 ? 3 : 4) << 8));
				else
					emit_record(d, s, ok, rv, 0, &out[p * 64 + clock]);
			}
		}
		__syncwarp();
	}
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

bool g_dec_tables_ready[16];

int upload_dec_tables(int device)
{
	if (device >= 0 && device < 16 && g_dec_tables_ready[device]) return BTBB_B200_OK;
	dec_tables t;
	memset(&t, 0, sizeof(t));
	uint8_t seq[127];
	uint32_t s = bt_whiten_seed(0);
	for (int i = 0; i < 127; i++) {
		if ((s & 0x40) && (s & 0x3f) < 64) t.phase[s & 0x3f] = (uint8_t)i;
		seq[i] = (uint8_t)bt_whiten_step(&s);
	}
	for (int i = 0; i < 13 * 32; i++)
		t.wseq[i >> 5] |= (uint32_t)seq[i % 127] << (i & 31);
	for (int b = 0; b < 256; b++) {
		uint32_t reg = (uint32_t)b;
		for (int i = 0; i < 8; i++) reg = (reg & 1) ? (reg >> 1) ^ 0x8408u : reg >> 1;
		t.crc[b] = (uint16_t)reg;
	}
	for (int i = 0; i < 10; i++) t.fec_col[i] = (uint8_t)bt_fec23_parity(1u << i);
	BT_CUDA_TRY(cudaMemcpyToSymbol(c_dec, &t, sizeof(t)));
	if (device >= 0 && device < 16) g_dec_tables_ready[device] = true;
	return BTBB_B200_OK;
}

}  // namespace

extern "C" int btbb_b200_decode_dev(btbb_b200_ctx *ctx, const uint8_t *d_stream, int64_t stream_length,
				    const btbb_b200_pkt_in *d_pkts, int64_t n, int mode,
				    btbb_b200_decoded *d_out, void *cuda_stream)
{
	if (!ctx || n < 0 || (n > 0 && (!d_stream || !d_pkts || !d_out)) || stream_length < 0 || mode < 0 || (mode > 3 && (mode < 16 || mode > 22)))
		return btbb_b200_set_error(BTBB_B200_EINVAL, "decode: bad arguments");
	BT_CUDA_TRY(cudaSetDevice(ctx->device));
	int rc = upload_dec_tables(ctx->device);
	if (rc) return rc;
	if (n == 0) return BTBB_B200_OK;
	int64_t blocks = (n + WARPS - 1) / WARPS;
	int64_t cap = (int64_t)ctx->sm_count * 16;
	if (blocks > cap) blocks = cap;
	decode_kernel<<<(unsigned)blocks, WARPS * 32, 0, (cudaStream_t)cuda_stream>>>(d_stream, stream_length, d_pkts, n, mode, d_out, NULL, NULL, NULL);
	BT_CUDA_TRY(cudaGetLastError());
	return BTBB_B200_OK;
}

/* try_clock + crc_check for CLK1-6 = 0..63 (bluetooth_piconet.c:675-689) of the packets listed in
 * d_idx[0 .. *d_n) (at most n_max of them), results as one 16-bit word per (packet, clock) at
 * d_tc[64 * packet + clock]: the UAP try_clock returned in the low byte, the crc_check class above
 * it (0, 1, 2 as returned; 3 = 10; 4 = 1000) */
int bt_try_clocks_compact(btbb_b200_ctx *ctx, const uint8_t *d_stream, int64_t stream_length,
			  const btbb_b200_pkt_in *d_pkts, const int64_t *d_idx, const unsigned long long *d_n,
			  int64_t n_max, uint16_t *d_tc, cudaStream_t st)
{
	int rc = upload_dec_tables(ctx->device);
	if (rc) return rc;
	if (n_max == 0) return BTBB_B200_OK;
	int64_t blocks = (n_max + WARPS - 1) / WARPS;
	int64_t cap = (int64_t)ctx->sm_count * 16;
	if (blocks > cap) blocks = cap;
	decode_kernel<<<(unsigned)blocks, WARPS * 32, 0, st>>>(d_stream, stream_length, d_pkts, n_max, BTBB_B200_MODE_TRY_CLOCKS,
							      NULL, d_tc, d_idx, d_n);
	BT_CUDA_TRY(cudaGetLastError());
	return BTBB_B200_OK;
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

extern "C" int btbb_b200_decode_host(btbb_b200_ctx *ctx, const char *stream, int64_t stream_length,
				     const btbb_b200_pkt_in *pkts, int64_t n, int mode, btbb_b200_decoded *out)
{
	if (!ctx || n < 0 || (n > 0 && (!stream || !pkts || !out)) || stream_length < 0 || mode < 0 || (mode > 3 && (mode < 16 || mode > 22)))
		return btbb_b200_set_error(BTBB_B200_EINVAL, "decode_host: bad arguments");
	if (n == 0) return BTBB_B200_OK;
	BT_CUDA_TRY(cudaSetDevice(ctx->device));
	uint8_t *d_s = NULL; btbb_b200_pkt_in *d_p = NULL; btbb_b200_decoded *d_o = NULL;
	int64_t nout = mode == BTBB_B200_MODE_TRY_CLOCKS ? n * 64 : n;
	int rc = BTBB_B200_OK;
	cudaError_t e;
	if ((e = cudaMalloc(&d_s, (size_t)stream_length + 1)) != cudaSuccess ||
	    (e = cudaMalloc(&d_p, (size_t)n * sizeof(*d_p))) != cudaSuccess ||
	    (e = cudaMalloc(&d_o, (size_t)nout * sizeof(*d_o))) != cudaSuccess)
		rc = btbb_b200_cuda_fail(e, "cudaMalloc(decode_host)");
	if (!rc && ((e = cudaMemcpy(d_s, stream, (size_t)stream_length, cudaMemcpyHostToDevice)) != cudaSuccess ||
		    (e = cudaMemcpy(d_p, pkts, (size_t)n * sizeof(*d_p), cudaMemcpyHostToDevice)) != cudaSuccess))
		rc = btbb_b200_cuda_fail(e, "cudaMemcpy(decode_host H2D)");
	if (!rc) rc = btbb_b200_decode_dev(ctx, d_s, stream_length, d_p, n, mode, d_o, NULL);
	if (!rc && (e = cudaMemcpy(out, d_o, (size_t)nout * sizeof(*d_o), cudaMemcpyDeviceToHost)) != cudaSuccess)
		rc = btbb_b200_cuda_fail(e, "cudaMemcpy(decode_host D2H)");
	if (d_s) cudaFree(d_s);
	if (d_p) cudaFree(d_p);
	if (d_o) cudaFree(d_o);
	return rc;
}
