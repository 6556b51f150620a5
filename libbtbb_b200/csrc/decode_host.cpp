/*
 * decode_host.cpp -- decode_core.h evaluated serially on the host, one packet at a time.
 *
 * This is the small-call path of the classic single-packet surface (compat.cu: btbb_decode_header,
 * btbb_decode_payload, try_clock, crc_check, fhs / DM / ..., btbb_header_present): a caller of the
 * reference hands over ONE packet of at most 3125 symbols per call (bluetooth_packet.c:1178-1317),
 * for which a kernel launch plus two PCIe round trips costs 10-100x the arithmetic (SURVEY.md
 * section 7, "Drop-in latency").  It is the same code the kernels run (decode_core.h), compiled for
 * the host; the batch entry points (btbb_b200_decode_dev / _host, the UAP sieve) never come here.
 */
#include <string.h>
#include "decode_core.h"
#include "capi_internal.h"

const btd_tables *btd_host_tables();

namespace {

void host_load(btd_pkt &p, const btd_tables *T, const char *symbols, int length, const btd_needs &n)
{
	memset(p.raw, 0, sizeof(p.raw));
	memset(p.fec0, 0, sizeof(p.fec0));
	memset(p.fec80, 0, sizeof(p.fec80));
	memset(p.hv1, 0, sizeof(p.hv1));
	p.sh = 0; p.length = length;
	int upto = n.symbols < length ? n.symbols : length;
	for (int i = 0; i < upto; i++)
		p.raw[i >> 5] |= (uint32_t)(symbols[i] & 1) << (i & 31);
	/* header: 18 triplets from symbol 68, ok iff fewer than 18 / 4 disagree (:563-567) */
	uint32_t hdr = 0; int bad = 0;
	for (int i = 0; i < 18; i++) {
		uint32_t b, d;
		btd_vote3(p, 68, i, &b, &d);
		hdr |= b << i; bad += (int)d;
	}
	p.hdr = hdr; p.hdr_ok = bad < 18 / 4;
	p.f8[0] = p.f8[1] = btd_bits(p.raw, 122, 8) * 0x01010101u;
	p.hv1_ok = 0;
	if (n.hv1) {
		bad = 0;
		for (int i = 0; i < 80; i++) {
			uint32_t b, d;
			btd_vote3(p, 122, i, &b, &d);
			p.hv1[i >> 5] |= b << (i & 31); bad += (int)d;
		}
		p.hv1_ok = bad < 80 / 4;
	}
	p.fail0 = p.fail80 = 1 << 20;
	for (int al = 0; al < 2; al++) {
		const int nblk = al ? n.nblk80 : n.nblk0, start = al ? 202 : 122;
		uint32_t *dst = al ? p.fec80 : p.fec0;
		int fail = 1 << 20;
		for (int b = 0; b < nblk; b++) {
			uint32_t data;
			if (!btd_fec23_block(btd_bits(p.raw, start + 15 * b, 15), T->s.fec_col, &data)) {
				if (fail == (1 << 20)) fail = b;
				continue;
			}
			const int pos = 10 * b;
			dst[pos >> 5] |= data << (pos & 31);
			if ((pos & 31) > 22) dst[(pos >> 5) + 1] |= data >> (32 - (pos & 31));
		}
		if (al) p.fail80 = fail; else p.fail0 = fail;
	}
	struct { int src, nbytes; uint16_t *dp; } tabs[4] = {
		{BTD_SRC_RAW, n.raw_bytes, p.dp_raw}, {BTD_SRC_FEC0, n.fec0_bytes, p.dp_fec0},
		{BTD_SRC_FEC80, n.fec80_bytes, p.dp_fec80}, {BTD_SRC_FIRST8, n.first8_bytes, p.dp_first8}};
	for (auto &t : tabs) {
		uint32_t acc = 0;
		t.dp[0] = 0;
		for (int j = 0; j < t.nbytes; j++) {
			acc ^= btd_byte_weight(T->nib, j, btd_src_byte(p, t.src, j));
			t.dp[j + 1] = (uint16_t)acc;
		}
	}
}

int run_search(const btd_ctx &c, const btd_pkt &p, const btd_lane &s)
{
	const int q18 = btd_q18(c, s.clock);
	for (int cand = s.s_lo; cand < s.s_hi; cand++)
		if (btd_cand_ok(c, p, s.pend, q18, s.uap, cand)) return cand;
	return -1;
}

void emit(const btd_ctx &c, const btd_pkt &p, const btd_lane &s, int header_ok, uint32_t hp, int raw_payload,
	  btbb_b200_decoded *o)
{
	uint32_t *w = reinterpret_cast<uint32_t *>(o);
	for (int i = 0; i < 7; i++) w[i] = btd_record_word(s, header_ok, hp, i);
	const int nbits = btd_emit_bits(s, raw_payload);
	const uint32_t *base;
	int pos, step, q = btd_q18(c, s.pay_clk);
	btd_src_desc(p, s.src, &base, &pos, &step);
	for (int j = 0; j < 86; j++) {
		w[7 + j] = btd_pay_word(c, base, pos, q, nbits - 32 * j);
		pos += step; q += 32; if (q >= 127) q -= 127;
	}
}

}  // namespace

/* One packet through the chain on the host.  mode as btbb_b200_decode_dev (flags included);
 * out holds 1 record (64 in BTBB_B200_MODE_TRY_CLOCKS). */
int bt_decode_one_cpu(const char *symbols, int length, uint32_t clkn, uint8_t uap, int whitened, uint8_t type,
		      int mode, btbb_b200_decoded *out)
{
	const btd_tables *T = btd_host_tables();
	const int raw_payload = (mode & BTBB_B200_MODE_FLAG_RAW_PAYLOAD) != 0;
	mode &= ~BTBB_B200_MODE_FLAG_RAW_PAYLOAD;
	if (length > BT_MAX_SYMBOLS) length = BT_MAX_SYMBOLS;
	if (length < 0) length = 0;
	btd_ctx c;
	c.s = &T->s; c.nib = T->nib; c.wp = T->wp; c.whitened = whitened;
	static thread_local btd_pkt p;
	btd_needs n;
	/* the header first: it decides which packet types are in play */
	btd_needs_for(0, length, 0, &n);
	host_load(p, T, symbols, length, n);
	uint32_t mask = 0;
	btd_lane s;
	if (mode == BTBB_B200_MODE_TRY_CLOCKS) {
		for (int clk = 0; clk < 64; clk++) {
			btd_lane_init(s, clk);
			btd_try_clock(c, p, s);
			mask |= 1u << s.type;
		}
	} else if (mode == BTBB_B200_MODE_DECODE) {
		uint32_t hp;
		btd_lane_init(s, (int)(clkn & 63));
		if (btd_decode_header(c, p, s, uap, &hp)) mask = 1u << s.type;
	} else
		mask = 1u << (type & 15);
	btd_needs_for(mask, length, 0, &n);
	host_load(p, T, symbols, length, n);
	if (mode == BTBB_B200_MODE_TRY_CLOCKS) {
		for (int clk = 0; clk < 64; clk++) {
			btd_lane_init(s, clk);
			btd_try_clock(c, p, s);
			btd_eval_begin(c, p, s, BTD_KIND_CRC_CHECK);
			const int found = s.pend ? run_search(c, p, s) : -1;
			btd_eval_end(p, s, BTD_KIND_CRC_CHECK, found);
			emit(c, p, s, p.hdr_ok, 0, raw_payload, &out[clk]);
		}
		return 0;
	}
	int header_ok = p.hdr_ok;
	uint32_t hp = 0;
	btd_lane_init(s, (int)(clkn & 63));
	if (mode == BTBB_B200_MODE_DECODE) {
		header_ok = btd_decode_header(c, p, s, uap, &hp);
		if (header_ok) {
			btd_eval_begin(c, p, s, BTD_KIND_PAYLOAD);
			const int found = s.pend ? run_search(c, p, s) : -1;
			btd_eval_end(p, s, BTD_KIND_PAYLOAD, found);
		}
	} else {
		const int kind = mode == BTBB_B200_MODE_PAYLOAD ? BTD_KIND_PAYLOAD
			       : mode == BTBB_B200_MODE_CRC_CHECK ? BTD_KIND_CRC_CHECK
			       : BTD_KIND_RAW + (mode - BTBB_B200_MODE_RAW);
		s.uap = uap; s.type = type & 15;
		btd_eval_begin(c, p, s, kind);
		const int found = s.pend ? run_search(c, p, s) : -1;
		btd_eval_end(p, s, kind, found);
	}
	emit(c, p, s, header_ok, hp, raw_payload, out);
	return 0;
}

/* btbb_header_present (:1371-1408) */
int bt_header_present_cpu(const char *s, int length)
{
	if (length < 122) return 0;
	const int msb = s[63] & 1;
	int be = 0;
	for (int i = 0; i < 4; i++) be += (s[64 + i] & 1) ^ ((i & 1) ? msb : !msb);
	for (int i = 0; i < 18; i++) {
		const int t = (s[68 + 3 * i] & 1) + (s[69 + 3 * i] & 1) + (s[70 + 3 * i] & 1);
		be += (t == 1 || t == 2);
	}
	return be < 5;
}
