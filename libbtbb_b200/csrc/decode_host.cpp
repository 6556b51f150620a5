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

/* The clock-independent state of the packet this thread worked on last (decode_core.h: btd_pkt), kept
 * between calls and grown on demand: a caller of the classic API asks about ONE packet again and again
 * (try_clock and crc_check for each of 64 clocks in UAP_from_header, bluetooth_piconet.c:597-671;
 * btbb_decode_header then btbb_decode_payload), and the bit packing, FEC decoding and CRC prefix sums
 * depend on the symbols only.  The entry is validated by comparing the symbols themselves, so a packet
 * object that was rewritten or reused can never be answered from stale state. */
struct host_pkt {
	btd_pkt p;
	btd_needs have;                 /* how far each part of p has been computed */
	uint32_t acc[4];                /* running CRC prefix value per table (raw, fec0, fec80, first8) */
	int length;
	bool valid;
	char sym[BT_MAX_SYMBOLS];
};

/* symbols [from, upto) into p.raw, eight per step where aligned */
void pack_symbols(btd_pkt &p, const char *sym, int from, int upto)
{
	int i = from;
	for (; i < upto && (i & 7); i++)
		p.raw[i >> 5] |= (uint32_t)(sym[i] & 1) << (i & 31);
	for (; i + 8 <= upto; i += 8) {
		uint64_t x;
		memcpy(&x, sym + i, 8);
		const uint32_t b = (uint32_t)(((x & 0x0101010101010101ull) * 0x0102040810204080ull) >> 56);
		p.raw[i >> 5] |= b << (i & 31);
	}
	for (; i < upto; i++)
		p.raw[i >> 5] |= (uint32_t)(sym[i] & 1) << (i & 31);
}

/* a fresh entry: symbols up to the end of the header and the first payload byte, the header vote */
void host_reset(host_pkt &H, const char *symbols, int length)
{
	btd_pkt &p = H.p;
	memcpy(H.sym, symbols, (size_t)length);
	H.length = length; H.valid = true;
	memset(p.raw, 0, sizeof(p.raw));
	memset(p.fec0, 0, sizeof(p.fec0));
	memset(p.fec80, 0, sizeof(p.fec80));
	memset(p.hv1, 0, sizeof(p.hv1));
	memset(&H.have, 0, sizeof(H.have));
	memset(H.acc, 0, sizeof(H.acc));
	p.dp_raw[0] = p.dp_fec0[0] = p.dp_fec80[0] = p.dp_first8[0] = 0;
	p.sh = 0; p.length = length;
	p.hv1_ok = 0; p.fail0 = p.fail80 = 1 << 20;
	H.have.symbols = length < 130 ? length : 130;
	pack_symbols(p, H.sym, 0, H.have.symbols);
	/* header: 18 triplets from symbol 68, ok iff fewer than 18 / 4 disagree (:563-567) */
	uint32_t hdr = 0; int bad = 0;
	for (int i = 0; i < 18; i++) {
		uint32_t b, d;
		btd_vote3(p, 68, i, &b, &d);
		hdr |= b << i; bad += (int)d;
	}
	p.hdr = hdr; p.hdr_ok = bad < 18 / 4;
	p.f8[0] = p.f8[1] = btd_bits(p.raw, 122, 8) * 0x01010101u;
}

/* grow the entry to what `n` asks for; every part only ever grows, and a prefix table is rewound to the
 * last byte that was complete when the bits under it grow */
void host_extend(host_pkt &H, const btd_tables *T, const btd_needs &n)
{
	btd_pkt &p = H.p;
	btd_needs &h = H.have;
	const int upto = n.symbols < H.length ? n.symbols : H.length;
	if (upto > h.symbols) {
		const int whole = h.symbols > 122 ? (h.symbols - 122) / 8 : 0;
		if (h.raw_bytes > whole) { h.raw_bytes = whole; H.acc[0] = p.dp_raw[whole]; }
		pack_symbols(p, H.sym, h.symbols, upto);
		h.symbols = upto;
	}
	if (n.hv1 && !h.hv1) {
		int bad = 0;
		for (int i = 0; i < 80; i++) {
			uint32_t b, d;
			btd_vote3(p, 122, i, &b, &d);
			p.hv1[i >> 5] |= b << (i & 31); bad += (int)d;
		}
		p.hv1_ok = bad < 80 / 4;
		h.hv1 = 1;
	}
	for (int al = 0; al < 2; al++) {
		const int nblk = al ? n.nblk80 : n.nblk0, start = al ? 202 : 122;
		int &done = al ? h.nblk80 : h.nblk0;
		if (nblk <= done) continue;
		int &bytes = al ? h.fec80_bytes : h.fec0_bytes;
		const int whole = 10 * done / 8;
		if (bytes > whole) { bytes = whole; H.acc[al ? 2 : 1] = (al ? p.dp_fec80 : p.dp_fec0)[whole]; }
		uint32_t *dst = al ? p.fec80 : p.fec0;
		int &fail = al ? p.fail80 : p.fail0;
		for (int b = done; b < nblk; b++) {
			uint32_t data;
			if (!btd_fec23_block(btd_bits(p.raw, start + 15 * b, 15), T->s.fec_col, &data)) {
				if (fail == (1 << 20)) fail = b;
				continue;
			}
			const int pos = 10 * b;
			dst[pos >> 5] |= data << (pos & 31);
			if ((pos & 31) > 22) dst[(pos >> 5) + 1] |= data >> (32 - (pos & 31));
		}
		done = nblk;
	}
	struct { int src, want, *have; uint16_t *dp; uint32_t *acc; } tabs[4] = {
		{BTD_SRC_RAW, n.raw_bytes, &h.raw_bytes, p.dp_raw, &H.acc[0]},
		{BTD_SRC_FEC0, n.fec0_bytes, &h.fec0_bytes, p.dp_fec0, &H.acc[1]},
		{BTD_SRC_FEC80, n.fec80_bytes, &h.fec80_bytes, p.dp_fec80, &H.acc[2]},
		{BTD_SRC_FIRST8, n.first8_bytes, &h.first8_bytes, p.dp_first8, &H.acc[3]}};
	for (auto &t : tabs) {
		uint32_t acc = *t.acc;
		for (int j = *t.have; j < t.want; j++) {
			acc ^= btd_byte_weight(T->nib, j, btd_src_byte(p, t.src, j));
			t.dp[j + 1] = (uint16_t)acc;
		}
		if (t.want > *t.have) { *t.have = t.want; *t.acc = acc; }
	}
}

host_pkt &host_get(const char *symbols, int length)
{
	static thread_local host_pkt H;
	if (!(H.valid && H.length == length && memcmp(H.sym, symbols, (size_t)length) == 0))
		host_reset(H, symbols, length);
	return H;
}

/* The decoder btd_eval_begin runs for (kind, type) when it is DM ('m') or DH ('h'): those two read the
 * payload length from the payload header, so the host decodes just the header's FEC blocks first and
 * then exactly the blocks and prefix sums that length needs, as the reference does (:898-958, :962-1011),
 * instead of the most the packet type could carry. */
int dm_or_dh(int kind, uint32_t type)
{
	if (kind >= BTD_KIND_RAW)
		return kind - BTD_KIND_RAW == 1 ? 'm' : kind - BTD_KIND_RAW == 2 ? 'h' : 0;
	switch (type) {
	case 3: case 8: case 10: case 14: return 'm';
	case 4: case 11: case 15: return 'h';
	case 9: return kind == BTD_KIND_PAYLOAD ? 'h' : 0;
	default: return 0;
	}
}

void needs_by_length(const btd_ctx &c, host_pkt &H, const btd_tables *T, int which, uint32_t type, int clock)
{
	const int size = H.length - 122;
	btd_needs n;
	memset(&n, 0, sizeof(n));
	n.symbols = 130;
	btd_lane t;
	btd_lane_init(t, clock);
	t.type = type;
	if (which == 'm') {
		const bool dv = type == 8;
		if (type != 3 && type != 8 && type != 10 && type != 14) return;
		const int hbytes = (type == 3 || type == 8) ? 1 : 2, first = dv ? 202 : 122, left = dv ? size - 80 : size;
		const int cap = left > 0 ? (left + 9) / 10 : 0;
		int nb = cap < 2 ? cap : 2;
		(dv ? n.nblk80 : n.nblk0) = nb;
		n.symbols = first + 15 * nb;
		host_extend(H, T, n);
		if (!btd_pay_hdr(c, H.p, t, hbytes, left, 1, dv)) return;
		nb = (8 * t.plen + 9) / 10;
		if (nb > cap) nb = cap;
		int bytes = t.plen;
		if (left < 8 * bytes) bytes = left > 0 ? left / 8 : 0;
		(dv ? n.nblk80 : n.nblk0) = nb;
		(dv ? n.fec80_bytes : n.fec0_bytes) = bytes;
		n.symbols = first + 15 * nb;
	} else {
		if (type != 4 && type != 9 && type != 11 && type != 15) return;
		const int hbytes = (type == 4 || type == 9) ? 1 : 2;
		n.symbols = 122 + 16;
		host_extend(H, T, n);
		if (!btd_pay_hdr(c, H.p, t, hbytes, size, 0, false)) return;
		int bytes = t.plen;
		if (size < 8 * bytes) bytes = size > 0 ? size / 8 : 0;
		n.raw_bytes = bytes;
		n.symbols = 122 + 8 * bytes;
	}
	if (n.symbols < 130) n.symbols = 130;
	host_extend(H, T, n);
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
	const int nw = (nbits + 31) / 32;
	for (int j = 0; j < nw; j++) {
		w[7 + j] = btd_pay_word(c, base, pos, q, nbits - 32 * j);
		pos += step; q += 32; if (q >= 127) q -= 127;
	}
	memset(w + 7 + nw, 0, (size_t)(86 - nw) * 4);
}

void grow_for_types(host_pkt &H, const btd_tables *T, uint32_t mask)
{
	btd_needs n;
	btd_needs_for(mask, H.length, 0, &n);
	host_extend(H, T, n);
}

}  // namespace

/* try_clock (:1178-1195) alone: UAP and packet type the header implies for one CLK1-6.  Returns
 * whether the header's FEC 1/3 vote passed (the reference leaves the packet untouched when it did not). */
int bt_try_clock_cpu(const char *symbols, int length, int clock, int whitened, uint8_t *uap, uint8_t *type)
{
	const btd_tables *T = btd_host_tables();
	if (length > BT_MAX_SYMBOLS) length = BT_MAX_SYMBOLS;
	if (length < 0) length = 0;
	btd_ctx c;
	c.s = &T->s; c.nib = T->nib; c.wp = T->wp; c.whitened = whitened;
	host_pkt &H = host_get(symbols, length);
	btd_lane s;
	btd_lane_init(s, clock & 63);
	btd_try_clock(c, H.p, s);
	*uap = (uint8_t)s.uap; *type = (uint8_t)s.type;
	return H.p.hdr_ok;
}

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
	host_pkt &H = host_get(symbols, length);
	const btd_pkt &p = H.p;
	btd_lane s;
	if (mode == BTBB_B200_MODE_TRY_CLOCKS) {
		/* the header decides which packet types are in play */
		uint32_t mask = 0;
		for (int clk = 0; clk < 64; clk++) {
			btd_lane_init(s, clk);
			btd_try_clock(c, p, s);
			mask |= 1u << s.type;
		}
		grow_for_types(H, T, mask);
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
	int header_ok = p.hdr_ok, kind;
	uint32_t hp = 0;
	btd_lane_init(s, (int)(clkn & 63));
	if (mode == BTBB_B200_MODE_DECODE) {
		kind = BTD_KIND_PAYLOAD;
		header_ok = btd_decode_header(c, p, s, uap, &hp);
	} else {
		kind = mode == BTBB_B200_MODE_PAYLOAD ? BTD_KIND_PAYLOAD
		     : mode == BTBB_B200_MODE_CRC_CHECK ? BTD_KIND_CRC_CHECK
		     : BTD_KIND_RAW + (mode - BTBB_B200_MODE_RAW);
		s.uap = uap; s.type = type & 15;
	}
	if (mode != BTBB_B200_MODE_DECODE || header_ok) {
		const int which = dm_or_dh(kind, s.type);
		if (which) needs_by_length(c, H, T, which, s.type, s.clock);
		else grow_for_types(H, T, btd_kind_type_mask(kind, s.type));
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
