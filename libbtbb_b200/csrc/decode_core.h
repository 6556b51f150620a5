/*
 * decode_core.h -- the per-packet chain (bluetooth_packet.c:552-705, 708-1317) as host/device
 * building blocks.  decode.cu runs them one warp per packet; decode_host.cpp runs the same
 * functions serially for the classic single-packet calls (compat.cu).
 *
 * What differs from the reference's bit loops:
 *
 *   * symbols are packed once per packet (bit i of a word array <-> symbol i), FEC 1/3 votes
 *     and every FEC 2/3 block are decoded once, independent of the clock candidate;
 *   * a payload CRC test costs O(1) instead of O(length).  crcgen (:671-690) is the LFSR
 *     reg' = A(reg) ^ bit * 0x8408 (A = shift right, feedback 0x8408), preloaded with
 *     reverse(UAP) << 8, and payload_crc (:772-781) compares it with the next 16 payload bits.
 *     Running those 16 bits through the register as well gives 0 exactly when they match, so
 *     for a payload of L bytes (CRC included) with bits e_i
 *           match  <=>  reverse(UAP) << 8  ==  XOR_{i < 8L} e_i * u_i ,   u_i = A^-(i+1)(0x8408)
 *     -- a prefix sum with fixed weights.  The payload is (data ^ whitening), so the right side
 *     splits into a per-packet table  dp[L] = XOR d_i u_i  (one per bit source: FEC 2/3 output,
 *     raw symbols, the DV alignment, the EV3/EV5 "first eight symbols" quirk) built once per
 *     packet from nibble tables, and a constant table  wp[q][L] = XOR w_{q+i} u_i  over the 127
 *     phases of the whitening sequence.  The length searches of EV3 / EV4 / EV5 (:1028-1040,
 *     :1071-1095, :1114-1126) and the 33 clocks fhs tries (:807-813) become table look-ups that a
 *     warp spreads over its lanes.
 */
#ifndef BTBB_B200_DECODE_CORE_H
#define BTBB_B200_DECODE_CORE_H

#include <stdint.h>
#include "bt_math.h"

#define BTD_LMAX        344     /* payload lengths 0..343 bytes (DH5) */
#define BTD_RAW_WORDS   113     /* 7 x 512 symbols from the 16-byte aligned base, + 1 */
#define BTD_FEC_BLOCKS  183     /* DM5: ceil(228 * 8 / 10) blocks of 15 symbols */
#define BTD_FEC_WORDS   60      /* 1830 bits + spill */
#define BTD_FEC80_BLOCKS 10     /* DV: 12 bytes */

enum { BTD_SRC_NONE = 0, BTD_SRC_FEC0, BTD_SRC_FEC80, BTD_SRC_RAW, BTD_SRC_HV1, BTD_SRC_FIRST8 };
enum { BTD_PEND_NONE = 0, BTD_PEND_FHS, BTD_PEND_EV35, BTD_PEND_EV4 };
/* what is evaluated for a (packet, clock): crc_check (:708-769), btbb_decode_payload (:1223-1297),
 * or one type decoder on its own (BTD_KIND_RAW + 0 fhs, 1 DM, 2 DH, 3 EV3, 4 EV4, 5 EV5, 6 HV) */
enum { BTD_KIND_CRC_CHECK = 0, BTD_KIND_PAYLOAD = 1, BTD_KIND_RAW = 16 };

/* constant tables (built by btd_build_tables, decode_tables.cpp) */
struct btd_small_tables {
	uint32_t wrot[128];       /* wrot[p]: 32 whitening bits starting at sequence position p < 127 */
	uint16_t wp20[64];        /* wp[(phase[c] + 18) % 127][20]: the fhs clock search */
	uint8_t phase[64];        /* sequence position where the LFSR state is 0x40 | clk */
	uint8_t q18[64];          /* (phase[clk] + 18) % 127: where the payload's whitening starts */
	uint8_t fec_col[16];      /* parity column of each FEC 2/3 data bit */
};
struct btd_tables {
	btd_small_tables s;
	uint16_t nib[BTD_LMAX * 2 * 16];     /* nib[(2j + h) * 16 + v] = XOR_{t<4, v_t} u_{8j + 4h + t} */
	uint16_t wp[127 * BTD_LMAX];         /* wp[q * BTD_LMAX + L] */
};

/* what the kernels / the host path see */
struct btd_ctx {
	const btd_small_tables *s;
	const uint16_t *nib;
	const uint16_t *wp;
	int whitened;             /* BTBB_WHITENED of the packet (:663) */
};

/* clock-independent state of one packet */
struct btd_pkt {
	/* CRC prefix tables first: 16-byte aligned, the two that length searches walk padded to 256 entries
	 * so that a warp can fetch eight consecutive entries per lane */
	uint16_t dp_first8[256], dp_fec0[256], dp_raw[BTD_LMAX], dp_fec80[16];
	uint32_t raw[BTD_RAW_WORDS];       /* bit (sh + i) = symbol i, 0 past `length` */
	uint32_t fec0[BTD_FEC_WORDS];      /* corrected data bits of the FEC 2/3 blocks from symbol 122 */
	uint32_t fec80[5];                 /* ... from symbol 202 (DV, :914) */
	uint32_t hv1[4];                   /* FEC 1/3 vote of the 240 symbols at 122 */
	uint32_t f8[2];                    /* the first eight payload symbols, replicated (EV3 / EV5 quirk) */
	uint32_t hdr;                      /* 18 voted header bits */
	int sh, length, hdr_ok, hv1_ok;
	int fail0, fail80;                 /* index of the first uncorrectable block (or a large value) */
};

/* per (packet, clock) result */
struct btd_lane {
	uint32_t uap, type, lt_addr, flags, hec, llid, flow, has_payload;
	int phl, plen, rv;
	int src, pay_clk, wbits;   /* where the payload bytes come from, and how many bits the decoder wrote */
	int pend, s_lo, s_hi;      /* pending search over candidates [s_lo, s_hi) */
	int aux, aux2;
	int clock;
};

/* how much of a packet the decoders of a set of packet types can touch */
struct btd_needs {
	int nblk0, nblk80, hv1;            /* FEC 2/3 blocks per alignment, the HV1 vote */
	int raw_bytes, fec0_bytes, fec80_bytes, first8_bytes;   /* prefix-table lengths */
	int symbols;                       /* symbols from the sync word that may be read */
};

#ifdef __CUDA_ARCH__
#define BTD_FSHR(lo, hi, s) __funnelshift_r((lo), (hi), (s))
#else
#define BTD_FSHR(lo, hi, s) (((s) & 31) ? (((lo) >> ((s) & 31)) | ((hi) << (32 - ((s) & 31)))) : (lo))
#endif

/* n <= 32 bits starting at bit `pos` of a word array (one word of slack is readable) */
BT_HD uint32_t btd_bits(const uint32_t *w, int pos, int n)
{
	const uint32_t v = BTD_FSHR(w[pos >> 5], w[(pos >> 5) + 1], (uint32_t)pos);
	return n >= 32 ? v : v & ((1u << n) - 1u);
}

/* 16 symbols (one byte each, bit 0 significant) -> 16 bits */
BT_HD uint32_t btd_pack4(uint32_t x) { return (((x & 0x01010101u) * 0x01020408u) >> 24) & 0xfu; }
BT_HD uint32_t btd_pack16(uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
	return btd_pack4(a) | (btd_pack4(b) << 4) | (btd_pack4(c) << 8) | (btd_pack4(d) << 12);
}
/* which of the 16 symbols of chunk c (aligned base) belong to the packet: symbol i sits at bit sh + i */
BT_HD uint32_t btd_chunk_mask(int c, int sh, int length)
{
	int lo = sh - 16 * c, hi = length + sh - 16 * c;
	if (lo < 0) lo = 0;
	if (hi > 16) hi = 16;
	if (hi <= lo) return 0;
	return ((1u << hi) - 1u) & ~((1u << lo) - 1u);
}

/* unfec13 (:552-568) of up to 32 triplets starting at symbol `first`; returns the vote bit and
 * whether the triplet disagrees */
BT_HD void btd_vote3(const btd_pkt &p, int first, int i, uint32_t *bit, uint32_t *bad)
{
	const uint32_t t = btd_bits(p.raw, p.sh + first + 3 * i, 3);
	const uint32_t ones = (t & 1u) + ((t >> 1) & 1u) + (t >> 2);
	*bit = ones >= 2;
	*bad = ones == 1 || ones == 2;
}

/* one (15,10) block (unfec23 :585-649): corrected data bits; false where the reference gives up */
BT_HD bool btd_fec23_block(uint32_t cw15, const uint8_t *col, uint32_t *data10)
{
	uint32_t data = cw15 & 0x3ffu, diff = (cw15 >> 10) & 0x1fu;
#ifdef __CUDA_ARCH__
	#pragma unroll
#endif
	for (int i = 0; i < 10; i++)
		if ((data >> i) & 1u) diff ^= col[i];
	if (diff & (diff - 1u)) {
		bool fixed = false;
#ifdef __CUDA_ARCH__
		#pragma unroll
#endif
		for (int i = 0; i < 10; i++)
			if (col[i] == diff) { data ^= 1u << i; fixed = true; }
		if (!fixed) return false;
	}
	*data10 = data;
	return true;
}

/* weight of byte value v at payload byte j in the CRC prefix sums */
BT_HD uint32_t btd_byte_weight(const uint16_t *nib, int j, uint32_t v)
{
	return (uint32_t)nib[(2 * j) * 16 + (v & 15u)] ^ (uint32_t)nib[(2 * j + 1) * 16 + ((v >> 4) & 15u)];
}

/* payload byte j of a bit source, before dewhitening */
BT_HD uint32_t btd_src_byte(const btd_pkt &p, int src, int j)
{
	switch (src) {
	case BTD_SRC_FEC0:   return btd_bits(p.fec0, 8 * j, 8);
	case BTD_SRC_FEC80:  return btd_bits(p.fec80, 8 * j, 8);
	case BTD_SRC_RAW:    return btd_bits(p.raw, p.sh + 122 + 8 * j, 8);
	case BTD_SRC_HV1:    return btd_bits(p.hv1, 8 * j, 8);
	case BTD_SRC_FIRST8: return btd_bits(p.raw, p.sh + 122, 8);      /* EV3 / EV5 quirk, :1036 / :1122 */
	default: return 0;
	}
}

BT_HD int btd_q(const btd_ctx &c, int clock, int pos)
{
	return ((int)c.s->phase[clock & 63] + pos) % 127;
}
BT_HD int btd_q18(const btd_ctx &c, int clock) { return (int)c.s->q18[clock & 63]; }

/* dewhitening bits for payload / header position pos .. pos + n (n <= 32) */
BT_HD uint32_t btd_white(const btd_ctx &c, int clock, int pos, int n)
{
	if (!c.whitened) return 0;
	const uint32_t v = c.s->wrot[btd_q(c, clock, pos)];
	return n >= 32 ? v : v & ((1u << n) - 1u);
}

/* payload_crc (:772-781) for a payload of L >= 2 bytes taken from the source behind dp[] */
BT_HD bool btd_check(const btd_ctx &c, const uint16_t *dp, int q18, int L, uint32_t uap)
{
	uint32_t v = dp[L];
	if (c.whitened) v ^= c.wp[q18 * BTD_LMAX + L];
	return v == bt_crc16_init(uap);
}

/* ---- what a set of packet types needs ---- */
BT_HD void btd_needs_for(uint32_t type_mask, int length, int payload_kind, btd_needs *n)
{
	const int size = length - 122;
	int nblk0 = 0, nblk80 = 0, raw = 0, f0 = 0, f80 = 0, f8 = 0, hv1 = 0;
	(void)payload_kind;
	if (type_mask & (1u << 2)) { if (nblk0 < 16) nblk0 = 16; if (f0 < 20) f0 = 20; }
	if (type_mask & (1u << 3)) { if (nblk0 < 16) nblk0 = 16; if (f0 < 20) f0 = 20; }
	if (type_mask & (1u << 10)) { if (nblk0 < 100) nblk0 = 100; if (f0 < 125) f0 = 125; }
	if (type_mask & (1u << 14)) { nblk0 = BTD_FEC_BLOCKS; f0 = 228; }
	if (type_mask & (1u << 12)) { if (nblk0 < 98) nblk0 = 98; if (f0 < 123) f0 = 123; }
	if (type_mask & (1u << 6)) { if (nblk0 < 16) nblk0 = 16; }
	if (type_mask & (1u << 8)) { nblk80 = BTD_FEC80_BLOCKS; f80 = 12; }
	if (type_mask & ((1u << 4) | (1u << 9))) { if (raw < 30) raw = 30; }
	if (type_mask & (1u << 11)) { if (raw < 187) raw = 187; }
	if (type_mask & (1u << 15)) { raw = 343; }
	if (type_mask & (1u << 5)) hv1 = 1;
	if (type_mask & (1u << 7)) { if (f8 < 32) f8 = 32; if (raw < 30) raw = 30; }
	if (type_mask & (1u << 13)) { f8 = 182; }
	/* nothing looks at payload bits beyond `size` (DM/DH compare bits with symbols, :944/:997) */
	const int cap = size > 0 ? size : 0, cap80 = size - 80 > 0 ? size - 80 : 0;
	if (nblk0 > (cap + 9) / 10) nblk0 = (cap + 9) / 10;
	if (nblk80 > (cap80 + 9) / 10) nblk80 = (cap80 + 9) / 10;
	if (raw > cap / 8) raw = cap / 8;
	if (f0 > cap / 8) f0 = cap / 8;
	if (f80 > cap80 / 8) f80 = cap80 / 8;
	if (f8 > cap / 8) f8 = cap / 8;
	n->nblk0 = nblk0; n->nblk80 = nblk80; n->hv1 = hv1;
	n->raw_bytes = raw; n->fec0_bytes = f0; n->fec80_bytes = f80; n->first8_bytes = f8;
	int sy = 122;
	if (122 + 15 * nblk0 > sy) sy = 122 + 15 * nblk0;
	if (nblk80 && 202 + 15 * nblk80 > sy) sy = 202 + 15 * nblk80;
	if (122 + 8 * raw > sy) sy = 122 + 8 * raw;
	if ((hv1 || (type_mask & (1u << 7))) && sy < 362) sy = 362;
	if (f8 && sy < 130) sy = 130;
	n->symbols = sy;
}

/* The packet types whose state the evaluation of (kind, type) can touch.  crc_check and
 * btbb_decode_payload dispatch on the type; a single type decoder (BTD_KIND_RAW + n) runs whatever the
 * packet's type field says -- fhs(), EV3(), EV4(), EV5() ignore it, DM(), DH() and HV() switch on it and
 * give up on a type that is not theirs (:898-1174). */
BT_HD uint32_t btd_kind_type_mask(int kind, uint32_t type)
{
	const uint32_t own = 1u << (type & 15u);
	if (kind < BTD_KIND_RAW) return own;
	switch (kind - BTD_KIND_RAW) {
	case 0: return 1u << 2;
	case 1: return own & ((1u << 3) | (1u << 8) | (1u << 10) | (1u << 14));
	case 2: return own & ((1u << 4) | (1u << 9) | (1u << 11) | (1u << 15));
	case 3: return 1u << 7;
	case 4: return 1u << 12;
	case 5: return 1u << 13;
	default: return own & ((1u << 5) | (1u << 6) | (1u << 7));
	}
}

/* ---- per (packet, clock) evaluation ---- */
BT_HD void btd_lane_init(btd_lane &s, int clock)
{
	s.uap = s.type = s.lt_addr = s.flags = s.hec = s.llid = s.flow = s.has_payload = 0;
	s.phl = s.plen = s.rv = 0;
	s.src = BTD_SRC_NONE; s.pay_clk = clock; s.wbits = 0;
	s.pend = BTD_PEND_NONE; s.s_lo = s.s_hi = 0; s.aux = s.aux2 = 0;
	s.clock = clock;
}

/* fhs (:783-818) */
BT_HD void btd_fhs_begin(const btd_ctx &c, const btd_pkt &p, btd_lane &s)
{
	const int size = p.length - 122;
	s.plen = 20;
	if (size < 240) { s.rv = 1; return; }
	if (p.fail0 < 16) { s.rv = 0; return; }
	s.src = BTD_SRC_FEC0; s.wbits = 160; s.pay_clk = s.clock;
	if (btd_check(c, p.dp_fec0, btd_q18(c, s.clock), 20, s.uap)) { s.rv = 1000; return; }
	s.pend = BTD_PEND_FHS; s.s_lo = 32; s.s_hi = 64;
}

/* decode_payload_header (:821-895) */
BT_HD int btd_pay_hdr(const btd_ctx &c, const btd_pkt &p, btd_lane &s, int hbytes, int size, int fec, bool dv)
{
	const int nb = hbytes * 8;
	if (size < nb) return 0;
	uint32_t ph;
	if (fec) {
		if (size < (hbytes == 2 ? 30 : 15)) return 0;
		if ((dv ? p.fail80 : p.fail0) < (hbytes == 2 ? 2 : 1)) return 0;
		ph = btd_bits(dv ? p.fec80 : p.fec0, 0, nb);
	} else
		ph = btd_bits(p.raw, p.sh + 122, nb);
	ph ^= btd_white(c, s.clock, 18, nb);
	s.plen = hbytes == 2 ? (int)((ph >> 3) & 0x3ffu) + 4 : (int)((ph >> 3) & 0x1fu) + 3;
	int maxlen;
	switch (s.type) {
	case 3: maxlen = 20; break;
	case 4: maxlen = 30; break;
	case 8: maxlen = 12; break;
	case 10: maxlen = 125; break;
	case 11: maxlen = 187; break;
	case 14: maxlen = 228; break;
	case 15: maxlen = 343; break;
	default: maxlen = 0;      /* AUX1 and everything else (:860-889) */
	}
	if (s.plen > maxlen) s.plen = maxlen;
	s.llid = ph & 3u;
	s.flow = (ph >> 2) & 1u;
	s.phl = hbytes;
	return 1;
}

/* DM (:898-958) */
BT_HD void btd_dm(const btd_ctx &c, const btd_pkt &p, btd_lane &s)
{
	int size = p.length - 122, hbytes = 2, maxlen;
	bool dv = false;
	switch (s.type) {
	case 8: dv = true; size -= 80; hbytes = 1; maxlen = 12; break;
	case 3: hbytes = 1; maxlen = 20; break;
	case 10: maxlen = 125; break;
	case 14: maxlen = 228; break;
	default: s.rv = 0; return;
	}
	if (!btd_pay_hdr(c, p, s, hbytes, size, 1, dv)) { s.rv = 0; return; }
	if (s.plen > maxlen) { s.rv = 1; return; }
	const int nbits = s.plen * 8;
	if (nbits > size) { s.rv = 1; return; }      /* bits against symbols, as the reference (:944) */
	if ((dv ? p.fail80 : p.fail0) < (nbits + 9) / 10) { s.rv = 0; return; }
	s.src = dv ? BTD_SRC_FEC80 : BTD_SRC_FEC0; s.pay_clk = s.clock; s.wbits = nbits;
	s.rv = btd_check(c, dv ? p.dp_fec80 : p.dp_fec0, btd_q18(c, s.clock), s.plen, s.uap) ? 10 : 2;
}

/* DH (:962-1011) */
BT_HD void btd_dh(const btd_ctx &c, const btd_pkt &p, btd_lane &s)
{
	int size = p.length - 122, hbytes = 2, maxlen;
	switch (s.type) {
	case 9: case 4: hbytes = 1; maxlen = 30; break;
	case 11: maxlen = 187; break;
	case 15: maxlen = 343; break;
	default: s.rv = 0; return;
	}
	if (!btd_pay_hdr(c, p, s, hbytes, size, 0, false)) { s.rv = 0; return; }
	if (s.plen > maxlen) { s.rv = 1; return; }
	const int nbits = s.plen * 8;
	if (nbits > size) { s.rv = 1; return; }
	s.src = BTD_SRC_RAW; s.pay_clk = s.clock; s.wbits = nbits;
	if (s.type == 9) { s.rv = 2; return; }
	s.rv = btd_check(c, p.dp_raw, btd_q18(c, s.clock), s.plen, s.uap) ? 10 : 2;
}

/* EV3 (:1013-1042) / EV5 (:1099-1128): every candidate length L in [3, min(maxlength, Lstop)) is one
 * table test; Lstop is the first length whose byte does not fit into the packet (:1030) */
BT_HD void btd_ev35_begin(const btd_pkt &p, btd_lane &s, int maxlength)
{
	const int size = p.length - 122;
	const int lstop = size >= 8 ? (size - 8) / 8 + 1 : 0;
	const int lend = lstop < maxlength ? lstop : maxlength;
	s.src = BTD_SRC_FIRST8; s.pay_clk = s.clock;
	s.aux = lstop; s.aux2 = maxlength;
	s.pend = BTD_PEND_EV35; s.s_lo = 3; s.s_hi = lend;      /* may be empty */
}
BT_HD void btd_ev35_end(btd_lane &s, int found)
{
	if (found >= 0) { s.rv = 10; s.plen = found; }
	else if (s.aux < s.aux2) { s.rv = 1; s.plen = s.aux; }
	else { s.rv = 2; s.plen = s.aux2; }
	s.wbits = 8 * s.plen;
}

/* EV4 (:1044-1097): block k is taken while k < 98, it fits (15k + 15 <= size) and decodes; the CRC
 * test for length L runs once block ceil(8L / 10) is in, lengths in increasing order */
BT_HD void btd_ev4_begin(const btd_pkt &p, btd_lane &s)
{
	const int size = p.length - 122;
	const int ks = size > 0 ? size / 15 : 0;
	int kmax = 98, why = 2;
	if (ks < kmax) { kmax = ks; why = 1; }
	if (p.fail0 < kmax) { kmax = p.fail0; why = kmax < 3 ? 0 : 1; }
	const int lmax = kmax >= 1 ? (10 * (kmax - 1)) / 8 : -1;
	s.src = BTD_SRC_FEC0; s.pay_clk = s.clock; s.wbits = 10 * kmax;
	s.aux = why; s.aux2 = kmax >= 1 ? lmax + 1 : 1;
	s.pend = BTD_PEND_EV4; s.s_lo = 2; s.s_hi = lmax + 1;
}
BT_HD void btd_ev4_end(btd_lane &s, int found)
{
	if (found >= 0) { s.rv = 10; s.plen = found; }
	else { s.rv = s.aux; s.plen = s.aux2; }
}

/* HV (:1131-1174) */
BT_HD void btd_hv(const btd_pkt &p, btd_lane &s)
{
	const int size = p.length - 122;
	s.phl = 0;
	if (size < 240) { s.plen = 0; s.rv = 1; return; }
	switch (s.type) {
	case 5:
		if (!p.hv1_ok) { s.rv = 0; return; }
		s.plen = 10; s.has_payload = 1; s.src = BTD_SRC_HV1; s.pay_clk = s.clock; s.wbits = 80;
		break;
	case 6:
		if (p.fail0 < 16) { s.rv = 0; return; }
		s.plen = 20; s.has_payload = 1; s.src = BTD_SRC_FEC0; s.pay_clk = s.clock; s.wbits = 160;
		break;
	case 7:
		s.plen = 30; s.has_payload = 1; s.src = BTD_SRC_RAW; s.pay_clk = s.clock; s.wbits = 240;
		break;
	}
	s.rv = 2;
}

/* one candidate of a pending search; q18 = btd_q18 of the clock the search belongs to */
BT_HD bool btd_cand_ok(const btd_ctx &c, const btd_pkt &p, int pend, int q18, uint32_t uap, int cand)
{
	if (pend == BTD_PEND_FHS) {
		uint32_t v = p.dp_fec0[20];
		if (c.whitened) v ^= c.s->wp20[cand & 63];
		return v == bt_crc16_init(uap);
	}
	return btd_check(c, pend == BTD_PEND_EV35 ? p.dp_first8 : p.dp_fec0, q18, cand, uap);
}

/* begin: everything up to a pending search (s.pend != 0) or a final s.rv */
BT_HD void btd_eval_begin(const btd_ctx &c, const btd_pkt &p, btd_lane &s, int kind)
{
	s.pend = BTD_PEND_NONE;
	if (kind == BTD_KIND_CRC_CHECK) {
		s.rv = 1;
		switch (s.type) {
		case 2: btd_fhs_begin(c, p, s); break;
		case 8: case 3: case 10: case 14: btd_dm(c, p, s); break;
		case 4: case 11: case 15: btd_dh(c, p, s); break;
		case 7: btd_ev35_begin(p, s, 32); break;
		case 12: btd_ev4_begin(p, s); break;
		case 13: btd_ev35_begin(p, s, 182); break;
		case 5: btd_hv(p, s); break;
		default: break;
		}
	} else if (kind == BTD_KIND_PAYLOAD) {
		s.rv = 0; s.phl = 0;
		switch (s.type) {
		case 0: case 1: s.plen = 0; s.rv = 1; break;
		case 2: btd_fhs_begin(c, p, s); break;
		case 3: case 8: case 10: case 14: btd_dm(c, p, s); break;
		case 4: case 9: case 11: case 15: btd_dh(c, p, s); break;
		case 5: case 6: btd_hv(p, s); break;
		case 7: btd_ev35_begin(p, s, 32); break;
		case 12: btd_ev4_begin(p, s); break;
		case 13: btd_ev35_begin(p, s, 182); break;
		}
	} else {
		s.rv = 0;
		switch (kind - BTD_KIND_RAW) {
		case 0: btd_fhs_begin(c, p, s); break;
		case 1: btd_dm(c, p, s); break;
		case 2: btd_dh(c, p, s); break;
		case 3: btd_ev35_begin(p, s, 32); break;
		case 4: btd_ev4_begin(p, s); break;
		case 5: btd_ev35_begin(p, s, 182); break;
		default: btd_hv(p, s); break;
		}
	}
}

/* end: `found` = first candidate of the pending search that passed, or -1 */
BT_HD void btd_eval_end(const btd_pkt &p, btd_lane &s, int kind, int found)
{
	switch (s.pend) {
	case BTD_PEND_FHS:
		if (found >= 0) { s.rv = 1000; s.pay_clk = found; } else { s.rv = 0; s.pay_clk = 63; }
		break;
	case BTD_PEND_EV35: btd_ev35_end(s, found); break;
	case BTD_PEND_EV4: btd_ev4_end(s, found); break;
	default: break;
	}
	s.pend = BTD_PEND_NONE;
	if (kind == BTD_KIND_CRC_CHECK) {
		/* crc_check's post-filters (:760-766) */
		if (s.rv == 0 && s.type != 2 && s.type != 3 && s.type != 5) s.rv = 1;
		if (s.rv > 1 && (s.type == 7 || s.type == 13)) s.rv = 1;
	} else if (kind == BTD_KIND_PAYLOAD) {
		if (s.type == 7 && s.rv <= 1) btd_hv(p, s);      /* EV3, then HV3 (:1262-1271) */
		s.has_payload = 1;
	}
}

/* try_clock (:1178-1195) for one clock: UAP and type from the voted header */
BT_HD void btd_try_clock(const btd_ctx &c, const btd_pkt &p, btd_lane &s)
{
	if (!p.hdr_ok) return;      /* unfec13 failed: UAP / type stay as they were (0 in a fresh packet) */
	const uint32_t hp = p.hdr ^ (c.whitened ? c.s->wrot[c.s->phase[s.clock & 63]] & 0x3ffffu : 0u);
	s.uap = bt_uap_from_hec(hp & 0x3ffu, hp >> 10);
	s.type = (hp >> 3) & 15u;
}

/* btbb_decode_header (:1198-1221); returns its return value, *hp = the dewhitened header bits */
BT_HD int btd_decode_header(const btd_ctx &c, const btd_pkt &p, btd_lane &s, uint32_t want_uap, uint32_t *hp_out)
{
	*hp_out = 0;
	s.uap = want_uap;
	if (!p.hdr_ok) return 0;
	const uint32_t hp = p.hdr ^ btd_white(c, s.clock, 0, 18);
	*hp_out = hp;
	if (bt_uap_from_hec(hp & 0x3ffu, hp >> 10) != want_uap) return 0;
	s.lt_addr = hp & 7u; s.type = (hp >> 3) & 15u; s.flags = (hp >> 7) & 7u; s.hec = hp >> 10;
	return 1;
}

/* ---- record emission ---- */
/* how many payload bits of the record are taken from the source; raw_payload = what the decoders
 * left in pkt->payload even when they failed (what btbb_pcap_append_packet logs, pcap.c:173-209) */
BT_HD int btd_emit_bits(const btd_lane &s, int raw_payload)
{
	if (s.plen <= 0 || s.plen > 344 || s.src == BTD_SRC_NONE) return 0;
	if (!raw_payload && s.rv < 2) return 0;
	const int nb = 8 * s.plen;
	return s.wbits < nb ? s.wbits : nb;
}

/* where a record's payload bits come from: word array, first bit, bits to advance per 32-bit word
 * (0 for the EV3 / EV5 quirk, where every byte is dewhitened from the same eight symbols) */
BT_HD void btd_src_desc(const btd_pkt &p, int src, const uint32_t **base, int *pos0, int *step)
{
	*step = 32; *pos0 = 0;
	switch (src) {
	case BTD_SRC_FEC0:   *base = p.fec0; break;
	case BTD_SRC_FEC80:  *base = p.fec80; break;
	case BTD_SRC_RAW:    *base = p.raw; *pos0 = p.sh + 122; break;
	case BTD_SRC_HV1:    *base = p.hv1; break;
	case BTD_SRC_FIRST8: *base = p.f8; *step = 0; break;
	default:             *base = p.f8; *step = 0; break;      /* nothing to emit: the caller's bit count is 0 */
	}
}

/* payload word j (bytes 4j .. 4j + 3 of the record's payload[]): pos = pos0 + step * j,
 * q = (q18 of pay_clk + 32 j) % 127, left = payload bits not yet emitted */
BT_HD uint32_t btd_pay_word(const btd_ctx &c, const uint32_t *base, int pos, int q, int left)
{
	if (left <= 0) return 0;      /* (also keeps the read inside the source array) */
	uint32_t d = btd_bits(base, pos, 32);
	if (c.whitened) d ^= c.s->wrot[q];
	return left >= 32 ? d : d & ((1u << left) - 1u);
}

/* the seven header words of a btbb_b200_decoded record */
BT_HD uint32_t btd_record_word(const btd_lane &s, int header_ok, uint32_t header_packed, int w)
{
	switch (w) {
	case 0: return (uint32_t)header_ok;
	case 1: return (uint32_t)s.rv;
	case 2: return (s.uap & 0xffu) | ((s.type & 0xffu) << 8) | ((s.lt_addr & 0xffu) << 16) | ((s.flags & 0xffu) << 24);
	case 3: return (s.hec & 0xffu) | ((s.llid & 0xffu) << 8) | ((s.flow & 0xffu) << 16) | ((s.has_payload & 0xffu) << 24);
	case 4: return (uint32_t)s.phl;
	case 5: return (uint32_t)s.plen;
	default: return header_packed;
	}
}

/* compact class of a crc_check result for the UAP sieve: UAP | class << 8 */
BT_HD uint32_t btd_tc16(const btd_lane &s)
{
	const uint32_t cls = s.rv == 0 ? 0u : s.rv == 1 ? 1u : s.rv == 2 ? 2u : s.rv == 10 ? 3u : 4u;
	return (s.uap & 0xffu) | (cls << 8);
}

#endif
