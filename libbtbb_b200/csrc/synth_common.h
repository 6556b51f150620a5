/*
 * synth_common.h -- deterministic synthetic BR/EDR capture (SURVEY.md 8d "Synthetic input").
 *
 * The reference has no transmitter; this forward encoder follows the air-frame
 * conventions its decoder implies (bluetooth_packet.c: sync word :188-199, trailer
 * :1384-1388, header at symbol 68 :1181, FEC 1/3 :552-568, whitening :653-668 with
 * offset 18 for the payload :800/:833/:950, payload at symbol 122 :906, FEC 2/3 :571-582,
 * CRC :671-690/:772-781).  Every symbol is a pure function of (cfg, global index), so the
 * host and the device produce identical bytes and each rank can generate its own shard.
 */
#ifndef BTBB_B200_SYNTH_COMMON_H
#define BTBB_B200_SYNTH_COMMON_H

#include "bt_math.h"
#include "../../include/btbb_b200.h"

#define SYNTH_WORDS 100   /* 3200 bits >= BT_MAX_SYMBOLS */

BT_HD uint64_t synth_noise_word(uint64_t seed, uint64_t w)
{
	return bt_splitmix64(seed ^ (w * 0x9E3779B97F4A7C15ULL));
}

BT_HD uint32_t synth_noise_symbol(uint64_t seed, int64_t g)
{
	return (uint32_t)(synth_noise_word(seed, (uint64_t)g >> 6) >> (g & 63)) & 1u;
}

/* independent bit flip with probability ber_q32 / 2^32 (second PRNG stream, seed+1) */
BT_HD uint32_t synth_flip(uint64_t seed, uint32_t ber_q32, int64_t g)
{
	if (!ber_q32) return 0;
	uint64_t h = bt_splitmix64((seed + 1) ^ (((uint64_t)g >> 1) * 0xD1342543DE82EF95ULL));
	uint32_t v = (g & 1) ? (uint32_t)(h >> 32) : (uint32_t)h;
	return v < ber_q32;
}

BT_HD uint32_t synth_lap(const btbb_b200_synth_cfg *c, uint32_t idx)
{
	if (c->n_laps <= 1) return c->fixed_lap & 0xffffffu;
	return (uint32_t)(bt_splitmix64(c->seed ^ 0x4C41505F5441424CULL ^ idx) & 0xffffffu);
}

/* air type code, max body bytes, payload-header bytes, fec (0 none, 1 = 1/3, 2 = 2/3), crc */
BT_HD void synth_kind_info(int kind, int *type, int *maxbody, int *hbytes, int *fec, int *crc)
{
	switch (kind) {
	case BTBB_B200_KIND_DM1: *type = 3;  *maxbody = 17;  *hbytes = 1; *fec = 2; *crc = 1; break;
	case BTBB_B200_KIND_DH1: *type = 4;  *maxbody = 27;  *hbytes = 1; *fec = 0; *crc = 1; break;
	case BTBB_B200_KIND_DM3: *type = 10; *maxbody = 121; *hbytes = 2; *fec = 2; *crc = 1; break;
	case BTBB_B200_KIND_FHS: *type = 2;  *maxbody = 18;  *hbytes = 0; *fec = 2; *crc = 1; break;
	case BTBB_B200_KIND_HV1: *type = 5;  *maxbody = 10;  *hbytes = 0; *fec = 1; *crc = 0; break;
	case BTBB_B200_KIND_DM5: *type = 14; *maxbody = 224; *hbytes = 2; *fec = 2; *crc = 1; break;
	case BTBB_B200_KIND_DH3: *type = 11; *maxbody = 183; *hbytes = 2; *fec = 0; *crc = 1; break;
	default:                 *type = -1; *maxbody = 0;   *hbytes = 0; *fec = 0; *crc = 0; break;
	}
}

BT_HD int synth_packet_symbols(int kind, int body)
{
	int type, maxbody, hbytes, fec, crc;
	synth_kind_info(kind, &type, &maxbody, &hbytes, &fec, &crc);
	if (type < 0) return 64;                      /* ID: the sync word alone */
	int bits = (hbytes + body + (crc ? 2 : 0)) * 8;
	int syms = fec == 2 ? ((bits + 9) / 10) * 15 : fec == 1 ? bits * 3 : bits;
	return 122 + syms;
}

/* ground truth of the packet planted in `slot` */
BT_HD void synth_params(const btbb_b200_synth_cfg *c, int64_t slot, btbb_b200_planted *p)
{
	uint64_t r0 = bt_splitmix64(c->seed ^ 0x504C414E54ULL ^ ((uint64_t)slot * 0xA0761D6478BD642FULL));
	uint64_t r1 = bt_splitmix64(r0);
	uint32_t mix = c->packet_mix & ((1u << BTBB_B200_KIND_COUNT) - 1);
	if (!mix) mix = 1u << BTBB_B200_KIND_ID;
	int nk = 0;
	for (int i = 0; i < BTBB_B200_KIND_COUNT; i++) nk += (mix >> i) & 1;
	int pick = (int)((r0 & 0xffff) % (uint32_t)nk), kind = 0;
	for (int i = 0; i < BTBB_B200_KIND_COUNT; i++)
		if ((mix >> i) & 1) { if (pick-- == 0) { kind = i; break; } }
	int type, maxbody, hbytes, fec, crc;
	synth_kind_info(kind, &type, &maxbody, &hbytes, &fec, &crc);
	int body = 0;
	if (type >= 0)
		body = (hbytes > 0) ? 1 + (int)(((r0 >> 16) & 0xffff) % (uint32_t)maxbody) : maxbody;
	int n = synth_packet_symbols(kind, body);
	if (n >= c->stride) { kind = BTBB_B200_KIND_ID; body = 0; n = 64; }
	int nl = c->n_laps < 1 ? 1 : (c->n_laps > 64 ? 64 : c->n_laps);
	p->lap = synth_lap(c, (uint32_t)((r0 >> 32) % (uint32_t)nl));
	p->uap = (uint8_t)(r1 & 0xff);
	p->kind = (uint8_t)kind;
	p->clk6 = (uint8_t)((r1 >> 8) & 0x3f);
	if (c->reserved & 1u) {
		/* piconet-coherent capture: one UAP per LAP and a piconet clock that advances by one per
		 * slot (a receiver that stamps packet `slot` with CLKN = slot sees CLK1-6 = CLKN + offset) */
		uint64_t rl = bt_splitmix64(c->seed ^ 0x5049434FULL ^ ((uint64_t)p->lap * 0x9E3779B97F4A7C15ULL));
		p->uap = (uint8_t)(rl & 0xff);
		p->clk6 = (uint8_t)(((rl >> 8) + (uint64_t)slot) & 0x3f);
	}
	p->lt_addr = (uint8_t)(1 + ((r1 >> 16) % 7));
	p->n_symbols = n;
	p->body_bytes = body;
	int64_t room = (int64_t)c->stride - n;
	p->offset = slot * (int64_t)c->stride + (room > 0 ? (int64_t)((r1 >> 24) % (uint64_t)room) : 0);
}

BT_HD void synth_put(uint32_t *bits, int pos, uint32_t b)
{
	bits[pos >> 5] |= (b & 1u) << (pos & 31);
}

/* content byte j of the unwhitened payload (payload header, then body) */
BT_HD uint32_t synth_payload_byte(const btbb_b200_planted *p, uint64_t pseed, int hbytes, int j)
{
	if (j < hbytes) {
		uint32_t llid = 1 + (uint32_t)(pseed % 3), flow = (uint32_t)(pseed >> 7) & 1u;
		uint32_t v = llid | (flow << 2) | ((uint32_t)p->body_bytes << 3);
		return (v >> (8 * j)) & 0xffu;
	}
	int k = j - hbytes;
	return (uint32_t)(bt_splitmix64(pseed + (uint64_t)(k >> 3)) >> (8 * (k & 7))) & 0xffu;
}

/* Encode the planted packet into bits[] (symbol i = bit i); returns the symbol count. */
BT_HD int synth_encode(const btbb_b200_synth_cfg *c, const btbb_b200_planted *p, uint32_t *bits)
{
	for (int i = 0; i < SYNTH_WORDS; i++) bits[i] = 0;
	uint64_t sw = bt_gen_syncword(p->lap);
	bits[0] = (uint32_t)sw; bits[1] = (uint32_t)(sw >> 32);
	int type, maxbody, hbytes, fec, crc;
	synth_kind_info(p->kind, &type, &maxbody, &hbytes, &fec, &crc);
	if (type < 0) return 64;
	uint64_t pseed = bt_splitmix64(c->seed ^ 0x424F4459ULL ^ ((uint64_t)p->offset * 0x8EBC6AF09C88C6E3ULL));
	uint32_t msb = (uint32_t)(sw >> 63);
	int pos = 64;
	for (int i = 0; i < 4; i++) synth_put(bits, pos++, (i & 1) ? msb : !msb);
	/* 18 header bits: LT_ADDR, TYPE, FLOW/ARQN/SEQN, HEC; whitened from position 0, sent 3x */
	uint32_t data10 = (p->lt_addr & 7u) | ((uint32_t)type << 3) | ((uint32_t)((pseed >> 40) & 7u) << 7);
	uint32_t hdr = data10 | (bt_hec(data10, p->uap) << 10);
	uint32_t ws = bt_whiten_seed(p->clk6);
	for (int i = 0; i < 18; i++) {
		uint32_t b = ((hdr >> i) & 1u) ^ bt_whiten_step(&ws);
		synth_put(bits, pos++, b); synth_put(bits, pos++, b); synth_put(bits, pos++, b);
	}
	/* payload: content, CRC (LSB first), whitening continues at position 18, then FEC */
	int nbytes = hbytes + p->body_bytes + (crc ? 2 : 0);
	int nbits = nbytes * 8, crc_from = crc ? (nbytes - 2) * 8 : nbits;
	uint32_t reg = bt_crc16_init(p->uap), blk = 0;
	int nblk = 0;
	uint32_t cur = 0;
	for (int j = 0; j < nbits; j++) {
		uint32_t b;
		if (j < crc_from) {
			if ((j & 7) == 0) cur = synth_payload_byte(p, pseed, hbytes, j >> 3);
			b = (cur >> (j & 7)) & 1u;
			reg = bt_crc16_step(reg, b);
		} else
			b = (reg >> (j - crc_from)) & 1u;
		b ^= bt_whiten_step(&ws);
		if (fec == 2) {
			blk |= b << nblk;
			if (++nblk == 10 || j == nbits - 1) {
				uint32_t cw = blk | (bt_fec23_parity(blk) << 10);
				for (int i = 0; i < 15; i++) synth_put(bits, pos++, (cw >> i) & 1u);
				blk = 0; nblk = 0;
			}
		} else if (fec == 1) {
			synth_put(bits, pos++, b); synth_put(bits, pos++, b); synth_put(bits, pos++, b);
		} else
			synth_put(bits, pos++, b);
	}
	return pos;
}

#endif
