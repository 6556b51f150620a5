/*
 * oracle/oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Plain-C restatement of the libbtbb hot path, written from the algorithm
 * (Bluetooth Core spec Vol 2 Part B sections 6.3.3, 7.2, 7.4, 7.5, 7.1) and
 * checked against the reference.  All citations are to
 * /root/reference/lib/src/bluetooth_packet.c unless noted.
 *
 * Parity status: PINNED -- tests/test_oracle_vs_golden.py checks every function
 * here against (a) the golden vectors in the reference's tests/ directory and
 * (b) the tests/golden fixtures generated from the unmodified reference
 * (oracle/_ref/libbtbb_ref.so) by tests/golden/make_golden.py.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <time.h>
#include "oracle.h"

/* ------------------------------------------------------------------ */
/* (64,30) sync-word code: g(x) = 0260534236651 octal (comment at :68-72). */
#define G34   0x585713DA9ULL            /* degree-34 generator, bit i <-> x^i */
#define PN64  0x83848D96BBCC54FCULL     /* PN overlay (:115) */
#define BARKER_A 0x27                   /* 7-bit tail when LAP bit 23 = 0 ... as seen in bits 57..63 */
#define BARKER_B 0x58

/* gen_syndrome (:147-159) is "codeword mod g(x)": bits 0..33 pass through,
 * bits 34..63 are reduced.  The reference does it with four byte tables
 * (sw_check_tables.h); this is the same linear map computed long-hand. */
uint64_t orc_syndrome(uint64_t cw)
{
	int i;
	for (i = 63; i >= 34; i--)
		if ((cw >> i) & 1)
			cw ^= G34 << (i - 34);
	return cw;
}

/* btbb_gen_syncword (:188-199).  Spec construction: LAP + 6 Barker-extension
 * bits, scrambled with PN bits 34..63, systematic parity from g(x), whole word
 * scrambled with PN again.  The reference folds all of that into sw_matrix. */
uint64_t orc_gen_syncword(uint32_t lap)
{
	uint64_t info = lap & 0xffffff;
	uint64_t ext = (lap & 0x800000) ? 0x13 /* a24..a29 = 110010 */ : 0x2c /* 001101 */;
	uint64_t cw;
	info |= ext << 24;
	info ^= PN64 >> 34;
	cw = info << 34;
	cw |= orc_syndrome(cw);
	return cw ^ PN64;
}

static int popcount64(uint64_t v) { return __builtin_popcountll(v); }

/* BARKER_DISTANCE (:55-59): distance of the 7 received bits 57..63 to the nearer
 * of the two legal tails. */
int orc_barker_distance(int b)
{
	int da = popcount64((uint64_t)((b ^ BARKER_A) & 0x7f));
	int db = 7 - da;
	return da < db ? da : db;
}

/* barker_correct (:81-113): nearest legal tail, positioned at bits 57..63. */
uint64_t orc_barker_correct(int b)
{
	int da = popcount64((uint64_t)((b ^ BARKER_A) & 0x7f));
	return (uint64_t)(da <= 3 ? BARKER_A : BARKER_B) << 57;
}

/* ------------------------------------------------------------------ */
/* syndrome -> error table (gen_syndrome_map / cycle / add_syndrome, :129-185):
 * every pattern of 1..k errors confined to bits 0..57.  Sorted array + bsearch
 * instead of uthash; keys are unique because d_min = 14. */
typedef struct { uint64_t syn, err; } tab_ent;
static tab_ent *g_tab;
static long g_tab_n;
static int g_tab_k = -1;
static pthread_mutex_t g_tab_lock = PTHREAD_MUTEX_INITIALIZER;

static uint64_t g_bit_syn[58];

static void enumerate(uint64_t err, uint64_t syn, int start, int depth, tab_ent **w)
{
	int i;
	for (i = start; i < 58; i++) {
		uint64_t e = err | (1ULL << i), s = syn ^ g_bit_syn[i];
		if (depth > 1)
			enumerate(e, s, i + 1, depth - 1, w);
		else {
			(*w)->syn = s; (*w)->err = e; (*w)++;
		}
	}
}

static int cmp_ent(const void *a, const void *b)
{
	uint64_t x = ((const tab_ent *)a)->syn, y = ((const tab_ent *)b)->syn;
	return x < y ? -1 : x > y;
}

/* btbb_init (:279-292).  Like the reference, the table is built once, for the
 * first non-zero k asked for, and never rebuilt. */
int orc_init(int max_ac_errors)
{
	static const long sizes[6] = {0, 58, 1711, 32567, 456837, 5038953};
	int i;
	if (max_ac_errors < 0 || max_ac_errors > 5)
		return -1;
	pthread_mutex_lock(&g_tab_lock);
	if (g_tab == NULL && max_ac_errors) {
		tab_ent *w;
		for (i = 0; i < 58; i++)
			g_bit_syn[i] = orc_syndrome(1ULL << i);
		g_tab = (tab_ent *)malloc(sizeof(tab_ent) * sizes[max_ac_errors]);
		w = g_tab;
		for (i = 1; i <= max_ac_errors; i++)
			enumerate(0, 0, 0, i, &w);
		g_tab_n = (long)(w - g_tab);
		qsort(g_tab, g_tab_n, sizeof(tab_ent), cmp_ent);
		g_tab_k = max_ac_errors;
	}
	if (g_tab_k < 0 && max_ac_errors == 0)
		g_tab_k = 0;
	pthread_mutex_unlock(&g_tab_lock);
	return 0;
}

int orc_table_errors(void) { return g_tab_k < 0 ? 0 : g_tab_k; }
long orc_table_entries(void) { return g_tab_n; }

/* find_syndrome (:139-145) */
int orc_lookup_error(uint64_t syn, uint64_t *error)
{
	long lo = 0, hi = g_tab_n - 1;
	while (lo <= hi) {
		long mid = (lo + hi) / 2;
		if (g_tab[mid].syn == syn) { *error = g_tab[mid].err; return 1; }
		if (g_tab[mid].syn < syn) lo = mid + 1; else hi = mid - 1;
	}
	return 0;
}

/* ------------------------------------------------------------------ */
/* One window test.  known: find_known_lap body (:430-438);
 * promiscuous: promiscuous_packet_search body (:385-416). */
static int test_window(uint64_t w, int known, uint64_t ac, int k, uint32_t *lap, uint8_t *nerr)
{
	if (known) {
		int d = popcount64(w ^ ac);
		*nerr = (uint8_t)d;
		return d <= k;
	} else {
		int b = (int)(w >> 57);
		uint64_t sw, syn, err = 0;
		int e = 0;
		if (orc_barker_distance(b) > 1)
			return 0;
		sw = (w & 0x01ffffffffffffffULL) | orc_barker_correct(b);
		syn = orc_syndrome(sw ^ PN64);
		if (syn) {
			if (g_tab_n && orc_lookup_error(syn, &err)) {
				sw ^= err;
				e = popcount64(err);    /* Barker fixes are not counted (:403-404) */
			} else
				e = 0xff;
		}
		if (e > k)
			return 0;
		*nerr = (uint8_t)e;
		*lap = (uint32_t)((sw >> 34) & 0xffffff);
		return 1;
	}
}

/* All positions btbb_find_ac (:444-464) reports when iterated with restart at
 * offset+1.  The 64-symbol window is kept packed and slid one symbol at a time
 * (air_to_host64, :235-242: symbol i <-> bit i). */
int64_t orc_find_all(const char *stream, int64_t n, uint32_t lap, int k,
		     btbb_b200_hit *hits, int64_t max_hits)
{
	int known = lap != BTBB_B200_LAP_ANY;
	uint64_t ac = known ? orc_gen_syncword(lap) : 0, w = 0;
	int64_t p, found = 0;
	int i;
	if (n <= 0) return 0;
	for (i = 0; i < 64; i++)
		w |= (uint64_t)(stream[i] & 1) << i;
	for (p = 0; p < n; p++) {
		uint32_t l = lap; uint8_t ne = 0;
		if (test_window(w, known, ac, k, &l, &ne)) {
			if (found < max_hits) {
				memset(&hits[found], 0, sizeof(hits[0]));
				hits[found].offset = p; hits[found].lap = l; hits[found].ac_errors = ne;
			}
			found++;
		}
		if (p + 1 < n)
			w = (w >> 1) | ((uint64_t)(stream[p + 64] & 1) << 63);
	}
	return found;
}

typedef struct { const char *s; int64_t b, e; uint32_t lap; int k; int64_t hits; } job_t;
static void *worker(void *a)
{
	job_t *j = (job_t *)a;
	btbb_b200_hit d;
	j->hits = orc_find_all(j->s + j->b, j->e - j->b, j->lap, j->k, &d, 0);
	return NULL;
}

/* contiguous-chunk partition over pthreads; returns elapsed seconds */
double orc_find_all_mt(const char *stream, int64_t n, uint32_t lap, int k, int threads, int64_t *total)
{
	pthread_t tid[256]; job_t job[256]; struct timespec t0, t1; int t;
	if (threads < 1) threads = 1;
	if (threads > 256) threads = 256;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	for (t = 0; t < threads; t++) {
		job[t].s = stream; job[t].b = n * t / threads; job[t].e = n * (t + 1) / threads;
		job[t].lap = lap; job[t].k = k; job[t].hits = 0;
		pthread_create(&tid[t], NULL, worker, &job[t]);
	}
	*total = 0;
	for (t = 0; t < threads; t++) { pthread_join(tid[t], NULL); *total += job[t].hits; }
	clock_gettime(CLOCK_MONOTONIC, &t1);
	return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

/* ------------------------------------------------------------------ */
/* FEC 1/3 (unfec13, :552-568): majority of each triplet; ok iff the number of
 * non-unanimous TRIPLETS is below length/4 (integer division). */
int orc_unfec13(const char *in, char *out, int length)
{
	int i, bad = 0;
	for (i = 0; i < length; i++) {
		int s = in[3 * i] + in[3 * i + 1] + in[3 * i + 2];
		out[i] = (char)(s >= 2);
		bad += (s == 1 || s == 2);
	}
	return bad < length / 4;
}

/* FEC 2/3 encoder (fec23, :571-582): (15,10) shortened Hamming,
 * g(D) = (D+1)(D^4+D+1) = D^5+D^4+D^2+1; parity = data(D)*D^5 mod g(D),
 * data bit i is the coefficient of D^(9-i) (first transmitted = highest power). */
uint16_t orc_fec23(uint16_t data)
{
	unsigned reg = 0; int i;
	for (i = 0; i < 10; i++) {
		unsigned fb = ((reg >> 4) ^ (data >> i)) & 1;
		reg = (reg << 1) & 0x1f;
		if (fb) reg ^= 0x15;   /* D^4 + D^2 + 1 */
	}
	/* register bit 4 is transmitted first -> parity bit 10 */
	{
		unsigned par = 0;
		for (i = 0; i < 5; i++)
			par |= ((reg >> (4 - i)) & 1) << i;
		return (uint16_t)((data & 0x3ff) | (par << 10));
	}
}

/* unfec23 (:585-649).  out receives ceil10(length) bits.  Returns 0 where the
 * reference returns NULL.  Syndrome 0 or a single set bit = clean / parity-bit
 * error; otherwise it must equal the parity column of one data bit. */
static int unfec23_block(const char *in15, char *out10)
{
	unsigned data = 0, chk = 0, diff; int i;
	for (i = 0; i < 10; i++) { out10[i] = in15[i]; data |= (unsigned)(in15[i] & 1) << i; }
	for (i = 0; i < 5; i++) chk |= (unsigned)(in15[10 + i] & 1) << i;
	diff = chk ^ (orc_fec23((uint16_t)data) >> 10);
	if (diff & (diff - 1)) {
		for (i = 0; i < 10; i++)
			if ((unsigned)(orc_fec23((uint16_t)(1u << i)) >> 10) == diff) { out10[i] ^= 1; return 1; }
		return 0;
	}
	return 1;
}

int orc_unfec23(const char *in, int length, char *out)
{
	int blocks = (length + 9) / 10, b;
	for (b = 0; b < blocks; b++)
		if (!unfec23_block(in + 15 * b, out + 10 * b))
			return 0;
	return 1;
}

/* Whitening (unwhiten, :653-668; INDICES :49, WHITENING_DATA :52): 7-bit LFSR
 * x^7+x^4+1 seeded with 1,CLK6..CLK1; bit `position` of the sequence. */
static unsigned wh_step(unsigned *s)
{
	unsigned o = (*s >> 6) & 1;
	*s = (*s << 1) & 0x7f;
	if (o) *s ^= 0x11;
	return o;
}

int orc_whiten_bit(int clk6, int position)
{
	unsigned s = 0x40 | (clk6 & 0x3f), o = 0;
	int i;
	position %= 127;
	for (i = 0; i <= position; i++) o = wh_step(&s);
	return (int)o;
}

void orc_unwhiten(const char *in, char *out, int clock, int length, int skip, int whitened)
{
	unsigned s = 0x40 | (clock & 0x3f);
	int i;
	for (i = 0; i < skip % 127; i++) wh_step(&s);
	for (i = 0; i < length; i++) {
		unsigned w = wh_step(&s);
		out[i] = whitened ? (char)(in[i] ^ w) : in[i];
	}
}

static uint8_t rev8(unsigned b)
{
	unsigned r = 0; int i;
	for (i = 0; i < 8; i++) r |= ((b >> i) & 1) << (7 - i);
	return (uint8_t)r;
}

/* crcgen (:671-690): reflected CRC-16/CCITT (poly 0x8408), register preloaded
 * with the bit-reversed UAP in its high byte. */
uint16_t orc_crc16(const char *bits, int length, int uap)
{
	unsigned reg = (unsigned)rev8((unsigned)uap & 0xff) << 8;
	int i;
	for (i = 0; i < length; i++) {
		unsigned fb = (reg ^ (unsigned)bits[i]) & 1;
		reg >>= 1;
		if (fb) reg ^= 0x8408;
	}
	return (uint16_t)reg;
}

/* HEC: reflected LFSR for g(D)=D^8+D^7+D^5+D^2+D+1 (0xE5 reflected), preloaded
 * with the bit-reversed UAP, fed the 10 header bits LSB first. */
uint8_t orc_hec(uint16_t data, uint8_t uap)
{
	unsigned reg = rev8(uap); int i;
	for (i = 0; i < 10; i++) {
		unsigned fb = (reg ^ (data >> i)) & 1;
		reg >>= 1;
		if (fb) reg ^= 0xE5;
	}
	return (uint8_t)reg;
}

/* uap_from_hec (:693-705): run the same LFSR backwards from the received HEC. */
uint8_t orc_uap_from_hec(uint16_t data, uint8_t hec)
{
	unsigned reg = hec; int i;
	for (i = 9; i >= 0; i--) {
		unsigned fb = reg >> 7;
		if (fb) reg ^= 0xE5;
		reg = ((reg << 1) & 0xff) | (fb ^ ((data >> i) & 1));
	}
	return rev8(reg);
}

/* ------------------------------------------------------------------ */
/* Packet under decode.  Mirrors the fields of struct btbb_packet
 * (bluetooth_packet.h:52-112) the chain touches; symbols past `length` read 0
 * like a freshly calloc'ed packet after btbb_packet_set_data (:467-480). */
typedef struct {
	char sym[3125 + 64];
	int length, whitened;
	uint8_t uap, type, lt_addr, flags, hec, llid, flow, has_payload;
	int phl, plen;
	char hdr[18], ph[16];
	/* the 8 bytes in front of payload[] in the reference struct are
	 * {payload_llid, payload_flow, pad, pad, payload_length (LE int)}; EV4 reads
	 * them through payload_crc() when payload_length == 1 (:1088-1091, :777-778). */
	char pre[8];
	char pay[2744 + 64];
} opkt;

static void opkt_load(opkt *p, const char *symbols, int length, int whitened)
{
	memset(p, 0, sizeof(*p));
	if (length > 3125) length = 3125;
	if (length < 0) length = 0;
	memcpy(p->sym, symbols, (size_t)length);
	p->length = length;
	p->whitened = whitened;
}

static unsigned bits_le(const char *b, int n)
{
	unsigned v = 0; int i;
	for (i = 0; i < n; i++) v |= (unsigned)(b[i] & 1) << i;
	return v;
}

/* payload_crc (:772-781) */
static int pay_crc_ok(opkt *p)
{
	int nbits = (p->plen - 2) * 8;
	unsigned crc, chk;
	if (nbits >= 0) {
		crc = orc_crc16(p->pay, nbits, p->uap);
		chk = bits_le(p->pay + nbits, 16);
	} else {
		/* payload_length == 1: zero-length CRC run and a 16-bit check word that
		 * starts 8 chars before payload[] in the reference struct */
		char tmp[16];
		p->pre[0] = (char)p->llid; p->pre[1] = (char)p->flow; p->pre[2] = p->pre[3] = 0;
		p->pre[4] = (char)(p->plen & 0xff); p->pre[5] = (char)((p->plen >> 8) & 0xff);
		p->pre[6] = (char)((p->plen >> 16) & 0xff); p->pre[7] = (char)((p->plen >> 24) & 0xff);
		memcpy(tmp, p->pre, 8); memcpy(tmp + 8, p->pay, 8);
		crc = orc_crc16(p->pay, 0, p->uap);
		/* air_to_host16 ORs (uint16_t)char << i: a char other than 0/1 spreads over
		 * higher bits exactly like this */
		{ int i; chk = 0; for (i = 0; i < 16; i++) chk |= ((unsigned)(uint16_t)tmp[i] << i) & 0xffff; }
	}
	return crc == chk;
}

/* fhs (:783-818) */
static int dec_fhs(opkt *p, int clock)
{
	char fec[160]; int size = p->length - 122, c;
	p->plen = 20;
	if (size < p->plen * 12) return 1;
	if (!orc_unfec23(p->sym + 122, 160, fec)) return 0;
	orc_unwhiten(fec, p->pay, clock, 160, 18, p->whitened);
	if (pay_crc_ok(p)) return 1000;
	for (c = 32; c < 64; c++) {
		orc_unwhiten(fec, p->pay, c, 160, 18, p->whitened);
		if (pay_crc_ok(p)) return 1000;
	}
	return 0;
}

/* decode_payload_header (:821-895) */
static int dec_pay_hdr(opkt *p, int start, int clock, int hbytes, int size, int fec)
{
	char tmp[20]; int nb = hbytes * 8, maxlen = 0;
	if (size < nb) return 0;
	if (fec) {
		if (size < (hbytes == 2 ? 30 : 15)) return 0;
		if (!orc_unfec23(p->sym + start, nb, tmp)) return 0;
		orc_unwhiten(tmp, p->ph, clock, nb, 18, p->whitened);
	} else
		orc_unwhiten(p->sym + start, p->ph, clock, nb, 18, p->whitened);
	if (hbytes == 2) p->plen = (int)bits_le(p->ph + 3, 10) + 4;
	else p->plen = (int)bits_le(p->ph + 3, 5) + 3;
	switch (p->type) {
	case 3: maxlen = 20; break;    /* DM1 */
	case 4: maxlen = 30; break;    /* DH1 */
	case 8: maxlen = 12; break;    /* DV */
	case 10: maxlen = 125; break;  /* DM3 */
	case 11: maxlen = 187; break;  /* DH3 */
	case 14: maxlen = 228; break;  /* DM5 */
	case 15: maxlen = 343; break;  /* DH5 */
	default: maxlen = 0;           /* AUX1 and everything else fall through to 0 (:860-889) */
	}
	if (p->plen > maxlen) p->plen = maxlen;
	p->llid = (uint8_t)bits_le(p->ph, 2);
	p->flow = (uint8_t)bits_le(p->ph + 2, 1);
	p->phl = hbytes;
	return 1;
}

/* DM (:898-958) */
static int dec_dm(opkt *p, int clock)
{
	int start = 122, size = p->length - 122, hbytes = 2, maxlen, nbits;
	char buf[2750];
	switch (p->type) {
	case 8: start += 80; size -= 80; hbytes = 1; maxlen = 12; break;
	case 3: hbytes = 1; maxlen = 20; break;
	case 10: maxlen = 125; break;
	case 14: maxlen = 228; break;
	default: return 0;
	}
	if (!dec_pay_hdr(p, start, clock, hbytes, size, 1)) return 0;
	if (p->plen > maxlen) return 1;
	nbits = p->plen * 8;
	if (nbits > size) return 1;     /* bits compared with symbols, as in the reference (:944) */
	if (!orc_unfec23(p->sym + start, nbits, buf)) return 0;
	orc_unwhiten(buf, p->pay, clock, nbits, 18, p->whitened);
	return pay_crc_ok(p) ? 10 : 2;
}

/* DH (:962-1011) */
static int dec_dh(opkt *p, int clock)
{
	int start = 122, size = p->length - 122, hbytes = 2, maxlen, nbits;
	switch (p->type) {
	case 9: case 4: hbytes = 1; maxlen = 30; break;
	case 11: maxlen = 187; break;
	case 15: maxlen = 343; break;
	default: return 0;
	}
	if (!dec_pay_hdr(p, start, clock, hbytes, size, 0)) return 0;
	if (p->plen > maxlen) return 1;
	nbits = p->plen * 8;
	if (nbits > size) return 1;
	orc_unwhiten(p->sym + start, p->pay, clock, nbits, 18, p->whitened);
	if (p->type == 9) return 2;
	return pay_crc_ok(p) ? 10 : 2;
}

/* EV3 (:1013-1042) and EV5 (:1099-1128): same loop, maxlength 32 / 182.  Every
 * byte is unwhitened from the FIRST eight payload symbols (the reference passes
 * `stream`, not `stream + bits`, :1036/:1122) with whitening offset 18+bits. */
static int dec_ev35(opkt *p, int clock, int maxlength)
{
	int size = p->length - 122;
	for (p->plen = 0; p->plen < maxlength; p->plen++) {
		int bits = p->plen * 8;
		if (bits + 8 > size) return 1;
		orc_unwhiten(p->sym + 122, p->pay + bits, clock, 8, 18 + bits, p->whitened);
		if (p->plen > 2 && pay_crc_ok(p)) return 10;
	}
	return 2;
}

/* EV4 (:1044-1097) */
static int dec_ev4(opkt *p, int clock)
{
	int size = p->length - 122, syms = 0, bits = 0;
	char blk[10];
	p->plen = 1;
	while (syms < 1470) {
		if (syms + 15 > size) return 1;
		if (!orc_unfec23(p->sym + 122 + syms, 10, blk))
			return syms < 45 ? 0 : 1;
		orc_unwhiten(blk, p->pay + bits, clock, 10, 18 + bits, p->whitened);
		while (p->plen * 8 <= bits) {
			if (pay_crc_ok(p)) return 10;
			p->plen++;
		}
		syms += 15; bits += 10;
	}
	return 2;
}

/* HV (:1131-1174) */
static int dec_hv(opkt *p, int clock)
{
	int size = p->length - 122;
	char tmp[160];
	p->phl = 0;
	if (size < 240) { p->plen = 0; return 1; }
	switch (p->type) {
	case 5:
		if (!orc_unfec13(p->sym + 122, tmp, 80)) return 0;
		p->plen = 10; p->has_payload = 1;
		orc_unwhiten(tmp, p->pay, clock, 80, 18, p->whitened);
		break;
	case 6:
		if (!orc_unfec23(p->sym + 122, 160, tmp)) return 0;
		p->plen = 20; p->has_payload = 1;
		orc_unwhiten(tmp, p->pay, clock, 160, 18, p->whitened);
		break;
	case 7:
		p->plen = 30; p->has_payload = 1;
		orc_unwhiten(p->sym + 122, p->pay, clock, 240, 18, p->whitened);
		break;
	}
	return 2;
}

/* try_clock (:1178-1195); returns unfec13 success */
static int do_try_clock(opkt *p, int clock)
{
	char h[18], u[18];
	if (!orc_unfec13(p->sym + 68, h, 18)) return 0;
	orc_unwhiten(h, u, clock, 18, 0, p->whitened);
	p->uap = orc_uap_from_hec((uint16_t)bits_le(u, 10), (uint8_t)bits_le(u + 10, 8));
	p->type = (uint8_t)bits_le(u + 3, 4);
	return 1;
}

/* crc_check (:708-769) */
static int do_crc_check(opkt *p, int clock)
{
	int rv = 1;
	switch (p->type) {
	case 2: rv = dec_fhs(p, clock); break;
	case 8: case 3: case 10: case 14: rv = dec_dm(p, clock); break;
	case 4: case 11: case 15: rv = dec_dh(p, clock); break;
	case 7: rv = dec_ev35(p, clock, 32); break;
	case 12: rv = dec_ev4(p, clock); break;
	case 13: rv = dec_ev35(p, clock, 182); break;
	case 5: rv = dec_hv(p, clock); break;
	default: break;
	}
	if (rv == 0 && p->type != 2 && p->type != 3 && p->type != 5) return 1;
	if (rv > 1 && (p->type == 7 || p->type == 13)) return 1;
	return rv;
}

/* btbb_decode_payload (:1223-1297) */
static int do_decode_payload(opkt *p, int clock)
{
	int rv = 0;
	p->phl = 0;
	switch (p->type) {
	case 0: case 1: p->plen = 0; rv = 1; break;
	case 2: rv = dec_fhs(p, clock); break;
	case 3: case 8: case 10: case 14: rv = dec_dm(p, clock); break;
	case 4: case 9: case 11: case 15: rv = dec_dh(p, clock); break;
	case 5: case 6: rv = dec_hv(p, clock); break;
	case 7:
		rv = dec_ev35(p, clock, 32);
		if (rv <= 1) rv = dec_hv(p, clock);
		break;
	case 12: rv = dec_ev4(p, clock); break;
	case 13: rv = dec_ev35(p, clock, 182); break;
	}
	p->has_payload = 1;
	return rv;
}

static int g_emit_raw;      /* orc_decode_one_raw: payload bytes as the decoder left them, whatever rv says */

static void emit(const opkt *p, int header_ok, int rv, btbb_b200_decoded *o)
{
	int i;
	memset(o, 0, sizeof(*o));
	o->header_ok = header_ok; o->rv = rv;
	o->uap = p->uap; o->type = p->type; o->lt_addr = p->lt_addr; o->flags = p->flags;
	o->hec = p->hec; o->llid = p->llid; o->flow = p->flow; o->has_payload = p->has_payload;
	o->payload_header_length = p->phl; o->payload_length = p->plen;
	o->header_packed = bits_le(p->hdr, 18);
	if ((rv >= 2 || g_emit_raw) && p->plen > 0 && p->plen <= 344)
		for (i = 0; i < p->plen; i++)
			o->payload[i] = (uint8_t)bits_le(p->pay + 8 * i, 8);
}

/* btbb_decode_header (:1198-1221) + btbb_decode_payload, silent. */
void orc_decode_one(const char *symbols, int length, uint32_t clkn, uint8_t uap,
		    int whitened, btbb_b200_decoded *out)
{
	static __thread opkt p;
	char h[18];
	int ok = 0, rv = 0;
	opkt_load(&p, symbols, length, whitened);
	p.uap = uap;
	if (orc_unfec13(p.sym + 68, h, 18)) {
		unsigned d, hec;
		orc_unwhiten(h, p.hdr, (int)clkn, 18, 0, whitened);
		d = bits_le(p.hdr, 10); hec = bits_le(p.hdr + 10, 8);
		if (orc_uap_from_hec((uint16_t)d, (uint8_t)hec) == uap) {
			p.lt_addr = (uint8_t)bits_le(p.hdr, 3);
			p.type = (uint8_t)bits_le(p.hdr + 3, 4);
			p.flags = (uint8_t)bits_le(p.hdr + 7, 3);
			p.hec = (uint8_t)hec;
			ok = 1;
		}
	}
	if (ok) rv = do_decode_payload(&p, (int)clkn);
	emit(&p, ok, rv, out);
}

/* as orc_decode_one, but payload[] holds what the reference's decoders leave in pkt->payload even
 * when the decode fails (what btbb_pcap_append_packet logs); not thread-safe (test helper) */
void orc_decode_one_raw(const char *symbols, int length, uint32_t clkn, uint8_t uap,
			int whitened, btbb_b200_decoded *out)
{
	g_emit_raw = 1;
	orc_decode_one(symbols, length, clkn, uap, whitened, out);
	g_emit_raw = 0;
}

void orc_try_clock_one(const char *symbols, int length, int clock, int whitened,
		       btbb_b200_decoded *out)
{
	static __thread opkt p;
	int ok, rv;
	opkt_load(&p, symbols, length, whitened);
	ok = do_try_clock(&p, clock);
	rv = do_crc_check(&p, clock);
	emit(&p, ok, rv, out);
}

/* One decoder on a packet whose UAP and type field the caller forces -- what the exported per-type
 * functions (bluetooth_packet.h:115-144) do when they are called directly: fn 0 fhs, 1 DM, 2 DH, 3 EV3,
 * 4 EV4, 5 EV5, 6 HV; -1 btbb_decode_payload, -2 crc_check.  header_ok = the header's unfec13 verdict;
 * raw != 0: payload bytes as the decoder left them, whatever rv says.  Not thread-safe with raw. */
void orc_typed_one(const char *symbols, int length, int clock, uint8_t uap, uint8_t type, int whitened,
		   int fn, int raw, btbb_b200_decoded *out)
{
	static __thread opkt p;
	char h[18];
	int ok, rv;
	opkt_load(&p, symbols, length, whitened);
	p.uap = uap; p.type = type;
	ok = orc_unfec13(p.sym + 68, h, 18);
	switch (fn) {
	case -2: rv = do_crc_check(&p, clock); break;
	case -1: rv = do_decode_payload(&p, clock); break;
	case 0: rv = dec_fhs(&p, clock); break;
	case 1: rv = dec_dm(&p, clock); break;
	case 2: rv = dec_dh(&p, clock); break;
	case 3: rv = dec_ev35(&p, clock, 32); break;
	case 4: rv = dec_ev4(&p, clock); break;
	case 5: rv = dec_ev35(&p, clock, 182); break;
	default: rv = dec_hv(&p, clock); break;
	}
	g_emit_raw = raw;
	emit(&p, ok, rv, out);
	g_emit_raw = 0;
}

/* btbb_header_present (:1371-1408) */
int orc_header_present(const char *s, int length)
{
	int be = 0, i, msb;
	if (length < 122) return 0;
	msb = s[63] & 1;
	for (i = 0; i < 4; i++)
		be += (s[64 + i] & 1) ^ ((i & 1) ? msb : !msb);
	for (i = 0; i < 18; i++) {
		int t = s[68 + 3 * i] + s[69 + 3 * i] + s[70 + 3 * i];
		be += (t == 1 || t == 2);
	}
	return be < 5;
}

/* ---- UAP / CLK1-6 discovery ----
 * btbb_process_packet in survey mode (bluetooth_piconet.c:851-858) around btbb_uap_from_header
 * (:648-750), for packets grouped by piconet.  Sequential, one packet object reused across the
 * 64 candidates exactly as the reference does. */
#define OF_UAP_VALID (1u << 2)
#define OF_CLK6_VALID (1u << 4)
#define OF_CLK27_VALID (1u << 5)
#define OF_HOP_INIT (1u << 9)
#define OF_GOT_FIRST (1u << 10)
#define OF_IS_AFH (1u << 11)
#define OF_LOOKS_AFH (1u << 12)

static void sieve_reset(btbb_b200_sieve *pn)                 /* reset(), :547-568 */
{
	pn->flags &= ~(OF_GOT_FIRST | OF_HOP_INIT | OF_UAP_VALID | OF_CLK6_VALID | OF_CLK27_VALID | OF_IS_AFH);
	if (pn->flags & OF_LOOKS_AFH) pn->flags |= OF_IS_AFH;
	pn->packets_observed = 0;
}

static void sieve_channel_seen(btbb_b200_sieve *pn, unsigned channel)   /* :142-150 */
{
	if (channel >= 80) return;
	if (!(pn->afh_map[channel / 8] & (1u << (channel % 8)))) {
		pn->afh_map[channel / 8] |= (uint8_t)(1u << (channel % 8));
		pn->used_channels++;
	}
}

static int sieve_uap_from_header(opkt *pkt, uint32_t clkn, unsigned channel, btbb_b200_sieve *pn)
{
	int count, remaining = 0, first_clock = 0;
	if (!(pn->flags & OF_GOT_FIRST)) pn->first_pkt_time = clkn;
	sieve_channel_seen(pn, channel);
	if (pn->packets_observed >= 1000) {                     /* MAX_PATTERN_LENGTH, :663-670 */
		sieve_reset(pn);
		return 0;
	}
	pn->packets_observed++;
	pn->total_packets_observed++;
	for (count = 0; count < 64; count++) {
		if (pn->clock6_candidates[count] > -1 || !(pn->flags & OF_GOT_FIRST)) {
			int clock = (int)((count + clkn - pn->first_pkt_time) % 64);
			int uap = do_try_clock(pkt, clock) ? pkt->uap : 0;  /* try_clock returns 0 when the FEC fails */
			int crc_chk = -1;
			if (!(pn->flags & OF_GOT_FIRST) || uap == pn->clock6_candidates[count])
				crc_chk = do_crc_check(pkt, clock);
			if ((pn->flags & OF_UAP_VALID) && uap != pn->uap)
				crc_chk = -1;
			switch (crc_chk) {
			case -1: case 0:
				pn->clock6_candidates[count] = -1;
				break;
			case 1: case 2:
				pn->clock6_candidates[count] = (int16_t)uap;
				first_clock = count;
				remaining++;
				break;
			default:
				pn->clk_offset = (count - (int)(pn->first_pkt_time & 0x3f)) & 0x3f;
				pn->uap = (uint8_t)uap;
				pn->flags |= OF_CLK6_VALID | OF_UAP_VALID;
				pn->total_packets_observed = 0;
				return 1;
			}
		}
	}
	pn->flags |= OF_GOT_FIRST;
	if (remaining == 1) {
		pn->clk_offset = (first_clock - (int)(pn->first_pkt_time & 0x3f)) & 0x3f;
		pn->uap = (uint8_t)pn->clock6_candidates[first_clock];
		pn->flags |= OF_CLK6_VALID | OF_UAP_VALID;
		pn->total_packets_observed = 0;
		return 1;
	}
	if (remaining == 0) sieve_reset(pn);
	return 0;
}

void orc_uap_sieve(const char *stream, int64_t stream_length, const btbb_b200_pkt_in *pkts, int64_t n_pkts,
		   const int64_t *group_start, int64_t n_groups, btbb_b200_sieve *states, int8_t *rv)
{
	static __thread opkt pkt;
	int64_t g, p;
	(void)n_pkts;
	for (g = 0; g < n_groups; g++) {
		btbb_b200_sieve *pn = &states[g];
		for (p = group_start[g]; p < group_start[g + 1]; p++) {
			const btbb_b200_pkt_in *in = &pkts[p];
			int length = in->length > 3125 ? 3125 : in->length, r = BTBB_B200_SIEVE_NOT_CALLED;
			if (in->offset < 0 || in->offset >= stream_length) length = 0;
			else if (in->offset + length > stream_length) length = (int)(stream_length - in->offset);
			if (length < 0) length = 0;
			opkt_load(&pkt, length > 0 ? stream + in->offset : stream, length, in->whitened);
			sieve_channel_seen(pn, in->reserved & 0xff);
			if (orc_header_present(pkt.sym, pkt.length) && !(pn->flags & OF_UAP_VALID))
				r = sieve_uap_from_header(&pkt, in->clkn, in->reserved & 0xff, pn);
			if (rv) rv[p] = (int8_t)r;
		}
	}
}

/* ---- hop sequence: gen_hops (bluetooth_piconet.c:311-362) in closed form,
 * one entry per sequence index (= CLK27-1), so that any range can be produced independently --
 * the shape a one-thread-per-entry kernel needs.  address = (UAP << 24 | LAP) & 0xfffffff
 * (gen_hop_pattern, :371-372); afh_map == NULL means all 79 channels (precalc, :171-193). ---- */
static int hop_perm5(int z, int p_high, int p_low)           /* perm5, :258-290 */
{
	static const int i1[14] = {0, 2, 1, 3, 0, 1, 0, 3, 1, 0, 2, 1, 0, 1};
	static const int i2[14] = {1, 3, 2, 4, 4, 3, 2, 4, 4, 3, 4, 3, 3, 2};
	int p = (p_high << 9) | p_low, i;
	for (i = 13; i >= 0; i--)
		if ((p >> i) & 1) {
			int a = (z >> i1[i]) & 1, b = (z >> i2[i]) & 1;
			if (a != b) z ^= (1 << i1[i]) | (1 << i2[i]);
		}
	return z;
}

void orc_hop_sequence(uint32_t address, const uint8_t *afh_map, int64_t first, int64_t n, uint8_t *out)
{
	int bank[79], used = 0, i;
	int64_t idx;
	const int a1 = (address >> 23) & 0x1f, b = (address >> 19) & 0x0f;                      /* address_precalc, :196-215 */
	const int c1 = ((address >> 4) & 0x10) + ((address >> 3) & 0x08) + ((address >> 2) & 0x04) +
		       ((address >> 1) & 0x02) + (address & 0x01);
	const int d1 = (address >> 10) & 0x1ff;
	const int e = ((address >> 7) & 0x40) + ((address >> 6) & 0x20) + ((address >> 5) & 0x10) +
		      ((address >> 4) & 0x08) + ((address >> 3) & 0x04) + ((address >> 2) & 0x02) + ((address >> 1) & 0x01);
	for (i = 0; i < 79; i++) {
		int chan = (i * 2) % 79;
		if (!afh_map) bank[i] = chan;
		else if (afh_map[chan / 8] & (1 << (chan % 8))) bank[used++] = chan;
	}
	for (idx = first; idx < first + n; idx++) {
		const int y1 = (int)(idx & 1);
		const int64_t pair = idx >> 1;
		const int x = (int)(pair & 31), k = (int)((pair >> 5) & 511), j = (int)((pair >> 14) & 31), ii = (int)((pair >> 19) & 31);
		const int a = a1 ^ ii, c = (c1 ^ j) ^ (y1 ? 0x1f : 0), d = d1 ^ k;
		const uint32_t f = (uint32_t)(16 * (pair >> 5)) % 79;
		const int perm = hop_perm5(((x + a) % 32) ^ b, c, d);
		if (afh_map)          /* note: f_dash = (base_f % 79) % used here, unlike single_hop (:425) */
			out[idx - first] = (uint8_t)bank[(perm + e + (int)(f % (uint32_t)used) + 32 * y1) % used];
		else
			out[idx - first] = (uint8_t)bank[(perm + e + (int)f + 32 * y1) % 79];
	}
}

