/*
 * oracle/ref_shim.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Compiles the UNMODIFIED reference translation unit where it lies
 * (/root/reference/lib/src/bluetooth_packet.c, found through -I) into
 * oracle/_ref/libbtbb_ref.so and adds thin exported wrappers so tests can
 * reach its `static` helpers and iterate btbb_find_ac quickly.  No reference
 * source is copied into this repository: this file only #includes it at
 * build time ("Reaching static functions").
 */
#define RELEASE "2020-12-R1"
#define VERSION "ref-oracle"
#include "bluetooth_packet.c"   /* -I/root/reference/lib/src */

#include <string.h>
#include <pthread.h>
#include <time.h>

uint64_t ref_gen_syndrome(uint64_t cw) { return gen_syndrome(cw); }
uint64_t ref_air_to_host64(const char *s, int bits) { return air_to_host64(s, bits); }
int ref_unfec13(char *in, char *out, int length) { return unfec13(in, out, length); }
uint16_t ref_fec23(uint16_t data) { return fec23(data); }
/* returns 1 and fills out[ceil10(length)] on success, 0 when the reference returns NULL */
int ref_unfec23(char *in, int length, char *out)
{
	char *o = unfec23(in, length);
	int padded = length % 10 ? length + 10 - length % 10 : length;
	if (!o) return 0;
	memcpy(out, o, padded);
	free(o);
	return 1;
}
void ref_unwhiten(char *in, char *out, int clock, int length, int skip, int whitened)
{
	btbb_packet p;
	memset(&p, 0, sizeof(p));
	btbb_packet_set_flag(&p, BTBB_WHITENED, whitened);
	unwhiten(in, out, clock, length, skip, &p);
}
uint16_t ref_crcgen(char *payload, int length, int uap) { return crcgen(payload, length, uap); }
uint8_t ref_uap_from_hec(uint16_t data, uint8_t hec) { return uap_from_hec(data, hec); }
uint8_t ref_reverse(uint8_t b) { return reverse((char)b); }
int ref_sizeof_packet(void) { return (int)sizeof(btbb_packet); }
uint8_t ref_barker_distance(int b) { return BARKER_DISTANCE[b & 127]; }
uint64_t ref_barker_correct(int b) { return barker_correct[b & 127]; }
uint8_t ref_whitening_bit(int i) { return WHITENING_DATA[i % 127]; }
uint8_t ref_whitening_index(int clk) { return INDICES[clk & 63]; }
long ref_syndrome_map_count(void) { return syndrome_map ? (long)HASH_COUNT(syndrome_map) : 0; }

/* One hit record, same layout as btbb_b200_hit (include/btbb_b200.h). */
typedef struct { int64_t offset; uint32_t lap; uint8_t ac_errors; uint8_t pad[3]; } ref_hit;

/*
 * Iterate the reference btbb_find_ac over [0, search_length) restarting at
 * offset+1 after every hit (SURVEY.md 8b "Required extension"): the set of
 * all positions the reference would report.  Calls are chunked below 2^31.
 * stream must hold search_length + 63 symbols.  Returns the total number of
 * hits; at most max_hits are stored.
 */
int64_t ref_find_all(char *stream, int64_t search_length, uint32_t lap,
		     int max_ac_errors, ref_hit *hits, int64_t max_hits)
{
	int64_t pos = 0, n = 0;
	btbb_packet *pkt = btbb_packet_new();
	while (pos < search_length) {
		int64_t left = search_length - pos;
		int chunk = left > (1 << 30) ? (1 << 30) : (int)left;
		int off = btbb_find_ac(stream + pos, chunk, lap, max_ac_errors, &pkt);
		if (off < 0) { pos += chunk; continue; }
		if (n < max_hits) {
			memset(&hits[n], 0, sizeof(ref_hit));
			hits[n].offset = pos + off;
			hits[n].lap = btbb_packet_get_lap(pkt);
			hits[n].ac_errors = btbb_packet_get_ac_errors(pkt);
		}
		n++;
		pos += (int64_t)off + 1;
	}
	btbb_packet_unref(pkt);
	return n;
}

/* ---- multi-threaded timing harness (bench.py --impl reference / cpu_baseline) ---- */
typedef struct {
	char *stream; int64_t begin, end; uint32_t lap; int k; int64_t hits;
} ref_job;

static void *ref_worker(void *arg)
{
	ref_job *j = (ref_job *)arg;
	ref_hit dummy;
	j->hits = ref_find_all(j->stream + j->begin, j->end - j->begin, j->lap, j->k, &dummy, 0);
	return NULL;
}

/*
 * Contiguous-chunk partition of [0, search_length) over `threads` pthreads
 * (the reference is re-entrant after btbb_init: syndrome_map is read-only).
 * Returns elapsed seconds; *total_hits gets the hit count.
 */
double ref_find_all_mt(char *stream, int64_t search_length, uint32_t lap,
		       int max_ac_errors, int threads, int64_t *total_hits)
{
	pthread_t tid[256];
	ref_job job[256];
	struct timespec t0, t1;
	int t;
	if (threads < 1) threads = 1;
	if (threads > 256) threads = 256;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	for (t = 0; t < threads; t++) {
		job[t].stream = stream;
		job[t].begin = search_length * t / threads;
		job[t].end = search_length * (t + 1) / threads;
		job[t].lap = lap; job[t].k = max_ac_errors; job[t].hits = 0;
		pthread_create(&tid[t], NULL, ref_worker, &job[t]);
	}
	*total_hits = 0;
	for (t = 0; t < threads; t++) { pthread_join(tid[t], NULL); *total_hits += job[t].hits; }
	clock_gettime(CLOCK_MONOTONIC, &t1);
	return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

/* the same partition, keeping the hits: thread t's list is that of [begin_t, end_t) with global offsets,
 * and the concatenation in thread order is the ascending list of the whole range (ranges partition the
 * positions; each window is tested on its own).  Returns the total; at most max_hits are stored. */
typedef struct { char *stream; int64_t begin, end; uint32_t lap; int k; ref_hit *buf; int64_t cap, n; } ref_job2;

static void *ref_worker2(void *arg)
{
	ref_job2 *j = (ref_job2 *)arg;
	int64_t i;
	for (;;) {
		j->n = ref_find_all(j->stream + j->begin, j->end - j->begin, j->lap, j->k, j->buf, j->cap);
		if (j->n <= j->cap) break;
		free(j->buf); j->cap = j->n + 16;
		j->buf = (ref_hit *)malloc((size_t)j->cap * sizeof(ref_hit));
	}
	for (i = 0; i < j->n; i++) j->buf[i].offset += j->begin;
	return NULL;
}

int64_t ref_find_all_mt_hits(char *stream, int64_t search_length, uint32_t lap, int max_ac_errors, int threads,
			     ref_hit *hits, int64_t max_hits)
{
	pthread_t tid[256];
	ref_job2 job[256];
	int64_t total = 0;
	int t;
	if (threads < 1) threads = 1;
	if (threads > 256) threads = 256;
	for (t = 0; t < threads; t++) {
		job[t].stream = stream;
		job[t].begin = search_length * t / threads;
		job[t].end = search_length * (t + 1) / threads;
		job[t].lap = lap; job[t].k = max_ac_errors; job[t].n = 0;
		job[t].cap = (job[t].end - job[t].begin) / 2000 + 1024;
		job[t].buf = (ref_hit *)malloc((size_t)job[t].cap * sizeof(ref_hit));
		pthread_create(&tid[t], NULL, ref_worker2, &job[t]);
	}
	for (t = 0; t < threads; t++) {
		int64_t i;
		pthread_join(tid[t], NULL);
		for (i = 0; i < job[t].n; i++, total++)
			if (total < max_hits) hits[total] = job[t].buf[i];
		free(job[t].buf);
	}
	return total;
}

/* ---- per-packet chain, silent (btbb_decode prints; call the two halves) ---- */
typedef struct {
	int32_t header_ok;       /* btbb_decode_header return */
	int32_t rv;              /* btbb_decode_payload return (0 when header fails) */
	uint8_t uap, type, lt_addr, flags, hec, llid, flow, has_payload;
	int32_t payload_header_length;
	int32_t payload_length;
	uint32_t header_packed;
	uint8_t payload[344];    /* packed bytes, payload_length of them */
} ref_decoded;

static void fill_decoded(btbb_packet *p, ref_decoded *out)
{
	int i;
	out->uap = p->UAP; out->type = p->packet_type; out->lt_addr = p->packet_lt_addr;
	out->flags = p->packet_flags; out->hec = p->packet_hec;
	out->llid = p->payload_llid; out->flow = p->payload_flow;
	out->has_payload = (uint8_t)btbb_packet_get_flag(p, BTBB_HAS_PAYLOAD);
	out->payload_header_length = p->payload_header_length;
	out->payload_length = p->payload_length;
	out->header_packed = btbb_packet_get_header_packed(p);
	/* payload is defined (include/btbb_b200.h) as the packed bytes when rv >= 2, else zero */
	if (out->rv >= 2 && p->payload_length > 0 && p->payload_length <= 344)
		for (i = 0; i < p->payload_length; i++)
			out->payload[i] = air_to_host8(&p->payload[i * 8], 8);
}

void ref_decode_one(char *symbols, int length, uint32_t clkn, uint8_t uap, int whitened, ref_decoded *out)
{
	btbb_packet *p = btbb_packet_new();
	int i;
	memset(out, 0, sizeof(*out));
	p->LAP = 0; p->flags = 0;
	btbb_packet_set_flag(p, BTBB_WHITENED, whitened);
	btbb_packet_set_data(p, symbols, length, 0, clkn << 1);
	btbb_packet_set_uap(p, uap);
	btbb_packet_set_flag(p, BTBB_CLK6_VALID, 1);
	btbb_packet_set_flag(p, BTBB_HAS_PAYLOAD, 0);
	out->header_ok = btbb_decode_header(p);
	if (out->header_ok)
		out->rv = btbb_decode_payload(p);
	fill_decoded(p, out);
	btbb_packet_unref(p);
}

/* as ref_decode_one, but payload[] holds pkt->payload as the decoders left it, whatever rv says
 * (what btbb_pcap_append_packet logs) */
void ref_decode_one_raw(char *symbols, int length, uint32_t clkn, uint8_t uap, int whitened, ref_decoded *out)
{
	btbb_packet *p = btbb_packet_new();
	int i;
	memset(out, 0, sizeof(*out));
	p->LAP = 0; p->flags = 0;
	btbb_packet_set_flag(p, BTBB_WHITENED, whitened);
	btbb_packet_set_data(p, symbols, length, 0, clkn << 1);
	btbb_packet_set_uap(p, uap);
	btbb_packet_set_flag(p, BTBB_CLK6_VALID, 1);
	btbb_packet_set_flag(p, BTBB_HAS_PAYLOAD, 0);
	out->header_ok = btbb_decode_header(p);
	if (out->header_ok)
		out->rv = btbb_decode_payload(p);
	fill_decoded(p, out);
	if (p->payload_length > 0 && p->payload_length <= 344)
		for (i = 0; i < p->payload_length; i++)
			out->payload[i] = air_to_host8(&p->payload[i * 8], 8);
	btbb_packet_unref(p);
}

/* one of the exported per-type decoders (bluetooth_packet.h:115-144: fn 0 fhs, 1 DM, 2 DH, 3 EV3, 4 EV4,
 * 5 EV5, 6 HV), btbb_decode_payload (-1) or crc_check (-2) on a packet whose UAP and type the caller forces */
void ref_typed_one(char *symbols, int length, int clock, uint8_t uap, uint8_t type, int whitened, int fn, int raw,
		   ref_decoded *out)
{
	btbb_packet *p = btbb_packet_new();
	char h[18];
	int i;
	memset(out, 0, sizeof(*out));
	p->flags = 0;
	btbb_packet_set_flag(p, BTBB_WHITENED, whitened);
	btbb_packet_set_data(p, symbols, length, 0, (uint32_t)clock << 1);
	btbb_packet_set_uap(p, uap);
	p->packet_type = type;
	out->header_ok = unfec13(p->symbols + 68, h, 18);
	switch (fn) {
	case -2: out->rv = crc_check(clock, p); break;
	case -1: out->rv = btbb_decode_payload(p); break;
	case 0: out->rv = fhs(clock, p); break;
	case 1: out->rv = DM(clock, p); break;
	case 2: out->rv = DH(clock, p); break;
	case 3: out->rv = EV3(clock, p); break;
	case 4: out->rv = EV4(clock, p); break;
	case 5: out->rv = EV5(clock, p); break;
	default: out->rv = HV(clock, p); break;
	}
	fill_decoded(p, out);
	if (raw && p->payload_length > 0 && p->payload_length <= 344)
		for (i = 0; i < p->payload_length; i++)
			out->payload[i] = air_to_host8(&p->payload[i * 8], 8);
	btbb_packet_unref(p);
}

/* try_clock + crc_check for one clock candidate (SURVEY 3.4 inner loop) */
void ref_try_clock_one(char *symbols, int length, int clock, int whitened, ref_decoded *out)
{
	btbb_packet *p = btbb_packet_new();
	char h[18];
	int i;
	memset(out, 0, sizeof(*out));
	p->flags = 0;
	btbb_packet_set_flag(p, BTBB_WHITENED, whitened);
	btbb_packet_set_data(p, symbols, length, 0, 0);
	out->header_ok = unfec13(p->symbols + 68, h, 18);
	try_clock(clock, p);
	out->rv = crc_check(clock, p);
	fill_decoded(p, out);
	btbb_packet_unref(p);
}

int ref_header_present(char *symbols, int length)
{
	btbb_packet *p = btbb_packet_new();
	int r;
	btbb_packet_set_data(p, symbols, length, 0, 0);
	r = btbb_header_present(p);
	btbb_packet_unref(p);
	return r;
}

/* ---- UAP / CLK1-6 discovery: the reference's own btbb_uap_from_header (bluetooth_piconet.c,
 * compiled as a second translation unit by the Makefile), driven the way btbb_process_packet
 * drives it in survey mode (:851-858) but on a piconet object we own, so that its state can be
 * carried in and out.  Record layouts as in include/btbb_b200.h. ---- */
#include <unistd.h>
#include <fcntl.h>
#include "bluetooth_piconet.h"

typedef struct {
	int64_t offset; int32_t length; uint32_t clkn; uint8_t uap, whitened, type, pad; uint32_t reserved;
} ref_pkt_in;
typedef struct {
	uint32_t flags, first_pkt_time; int32_t clk_offset, packets_observed, total_packets_observed;
	uint8_t uap, used_channels, afh_map[10]; int16_t clock6_candidates[64];
} ref_sieve;

void ref_uap_sieve(char *stream, int64_t stream_length, const ref_pkt_in *pkts, int64_t n_pkts,
		   const int64_t *group_start, int64_t n_groups, ref_sieve *states, int8_t *rv)
{
	int64_t g, p;
	int i, saved, devnull;
	static char zero[1];
	(void)n_pkts;
	fflush(stdout);                    /* btbb_uap_from_header prints its findings */
	saved = dup(1);
	devnull = open("/dev/null", O_WRONLY);
	dup2(devnull, 1);
	for (g = 0; g < n_groups; g++) {
		btbb_piconet *pn = btbb_piconet_new();
		ref_sieve *st = &states[g];
		pn->flags = st->flags; pn->first_pkt_time = st->first_pkt_time; pn->clk_offset = st->clk_offset;
		pn->packets_observed = st->packets_observed; pn->total_packets_observed = st->total_packets_observed;
		pn->UAP = st->uap; pn->used_channels = st->used_channels;
		for (i = 0; i < 10; i++) pn->afh_map[i] = st->afh_map[i];
		for (i = 0; i < 64; i++) pn->clock6_candidates[i] = st->clock6_candidates[i];
		for (p = group_start[g]; p < group_start[g + 1]; p++) {
			const ref_pkt_in *in = &pkts[p];
			btbb_packet *pkt = btbb_packet_new();
			int length = in->length > 3125 ? 3125 : in->length, r = -2;
			if (in->offset < 0 || in->offset >= stream_length) length = 0;
			else if (in->offset + length > stream_length) length = (int)(stream_length - in->offset);
			if (length < 0) length = 0;
			btbb_packet_set_flag(pkt, BTBB_WHITENED, in->whitened);
			btbb_packet_set_data(pkt, length > 0 ? stream + in->offset : zero, length, (uint8_t)(in->reserved & 0xff), 0);
			pkt->clkn = in->clkn;              /* as stored, i.e. after btbb_packet_set_data's >> 1 */
			btbb_piconet_set_channel_seen(pn, pkt->channel);
			if (btbb_header_present(pkt) && !btbb_piconet_get_flag(pn, BTBB_UAP_VALID))
				r = btbb_uap_from_header(pkt, pn);
			if (rv) rv[p] = (int8_t)r;
			btbb_packet_unref(pkt);
		}
		st->flags = pn->flags; st->first_pkt_time = pn->first_pkt_time; st->clk_offset = pn->clk_offset;
		st->packets_observed = pn->packets_observed; st->total_packets_observed = pn->total_packets_observed;
		st->uap = pn->UAP; st->used_channels = pn->used_channels;
		for (i = 0; i < 10; i++) st->afh_map[i] = pn->afh_map[i];
		for (i = 0; i < 64; i++) st->clock6_candidates[i] = (int16_t)pn->clock6_candidates[i];
		btbb_piconet_unref(pn);
	}
	fflush(stdout);
	dup2(saved, 1);
	close(saved); close(devnull);
}

/* ---- pcap BR/EDR emission: the reference's own btbb_pcap_create_file / btbb_pcap_append_packet
 * (pcap.c, compiled as a further translation unit) on packets decoded by the reference.
 * For packet i: symbols at stream + offset[i] (length[i]), found with lap[i] / ac_errors[i],
 * decoded with clkn[i] / uap[i], logged with the given timestamps and powers.  rv[i] receives
 * header_ok ? btbb_decode_payload's value : -1. ---- */
typedef struct { uint64_t ns; int8_t sigdbm, noisedbm; uint8_t channel, transport, modulation, pad[3]; } ref_pcap_meta;

int ref_pcap_bredr(const char *path, char *stream, int64_t stream_length, const ref_hit *hits, const ref_pkt_in *pkts,
		   const ref_pcap_meta *meta, int64_t n, uint32_t reflap, uint8_t refuap, int32_t *rv)
{
	btbb_pcap_handle *h = NULL;
	int64_t i;
	static char zero[1];
	if (btbb_pcap_create_file(path, &h)) return -1;
	for (i = 0; i < n; i++) {
		btbb_packet *p = btbb_packet_new();
		int length = pkts[i].length > 3125 ? 3125 : pkts[i].length, ok, r = -1;
		if (pkts[i].offset < 0 || pkts[i].offset >= stream_length) length = 0;
		else if (pkts[i].offset + length > stream_length) length = (int)(stream_length - pkts[i].offset);
		p->LAP = hits[i].lap; p->ac_errors = hits[i].ac_errors; p->flags = 0;
		btbb_packet_set_flag(p, BTBB_WHITENED, pkts[i].whitened);
		btbb_packet_set_data(p, length > 0 ? stream + pkts[i].offset : zero, length, meta[i].channel, pkts[i].clkn << 1);
		btbb_packet_set_uap(p, pkts[i].uap);
		btbb_packet_set_flag(p, BTBB_CLK6_VALID, 1);
		btbb_packet_set_transport(p, meta[i].transport);
		btbb_packet_set_modulation(p, meta[i].modulation);
		ok = btbb_decode_header(p);
		if (ok) r = btbb_decode_payload(p);
		if (rv) rv[i] = r;
		btbb_pcap_append_packet(h, meta[i].ns, meta[i].sigdbm, meta[i].noisedbm, reflap, refuap, p);
		btbb_packet_unref(p);
	}
	btbb_pcap_close(h);
	return 0;
}

/* the same through the reference's pcapng writer (pcapng-bt.c:134-264) */
int ref_pcapng_bredr(const char *path, char *stream, int64_t stream_length, const ref_hit *hits, const ref_pkt_in *pkts,
		   const ref_pcap_meta *meta, int64_t n, uint32_t reflap, uint8_t refuap, int32_t *rv)
{
	btbb_pcapng_handle *h = NULL;
	int64_t i;
	static char zero[1];
	if (btbb_pcapng_create_file(path, "b200 parity", &h)) return -1;
	for (i = 0; i < n; i++) {
		btbb_packet *p = btbb_packet_new();
		int length = pkts[i].length > 3125 ? 3125 : pkts[i].length, ok, r = -1;
		if (pkts[i].offset < 0 || pkts[i].offset >= stream_length) length = 0;
		else if (pkts[i].offset + length > stream_length) length = (int)(stream_length - pkts[i].offset);
		p->LAP = hits[i].lap; p->ac_errors = hits[i].ac_errors; p->flags = 0;
		btbb_packet_set_flag(p, BTBB_WHITENED, pkts[i].whitened);
		btbb_packet_set_data(p, length > 0 ? stream + pkts[i].offset : zero, length, meta[i].channel, pkts[i].clkn << 1);
		btbb_packet_set_uap(p, pkts[i].uap);
		btbb_packet_set_flag(p, BTBB_CLK6_VALID, 1);
		btbb_packet_set_transport(p, meta[i].transport);
		btbb_packet_set_modulation(p, meta[i].modulation);
		ok = btbb_decode_header(p);
		if (ok) r = btbb_decode_payload(p);
		if (rv) rv[i] = r;
		btbb_pcapng_append_packet(h, meta[i].ns, meta[i].sigdbm, meta[i].noisedbm, reflap, refuap, p);
		btbb_packet_unref(p);
	}
	btbb_pcapng_close(h);
	return 0;
}

/* ---- hop sequence: the reference's gen_hop_pattern (bluetooth_piconet.c:365-377) for one address,
 * entries [first, first + n) of its 2^27-entry table copied out.  afh_map == NULL: all channels. ---- */
void ref_hop_sequence(uint32_t address, const uint8_t *afh_map, int64_t first, int64_t n, uint8_t *out)
{
	btbb_piconet *pn = btbb_piconet_new();
	int i, saved, devnull;
	pn->LAP = address & 0xffffff;
	pn->UAP = (address >> 24) & 0xff;
	if (afh_map) {
		btbb_piconet_set_flag(pn, BTBB_IS_AFH, 1);
		for (i = 0; i < 10; i++) pn->afh_map[i] = afh_map[i];
		pn->used_channels = 0;
		for (i = 0; i < 79; i++) pn->used_channels += (afh_map[i / 8] >> (i % 8)) & 1;
	} else
		pn->used_channels = 79;     /* gen_hops divides by it even without AFH (:356); any packet seen makes it > 0 */
	fflush(stdout);
	saved = dup(1); devnull = open("/dev/null", O_WRONLY); dup2(devnull, 1);
	gen_hop_pattern(pn);
	fflush(stdout);
	dup2(saved, 1); close(saved); close(devnull);
	memcpy(out, pn->sequence + first, (size_t)n);
	free(pn->sequence);
	btbb_piconet_unref(pn);
}


/* ---- hop reversal: the reference's own btbb_init_hop_reversal (bluetooth_piconet.c:475-499) and
 * btbb_winnow (:613-645) on a piconet whose pattern log holds the given observations, one more per
 * call.  counts[j] = pn->num_candidates after observation j (-1 once the reference has stopped: it
 * breaks at <= 1 candidates).  cands receives the surviving CLK1-27 values.  Returns the final count. ---- */
int ref_hop_winnow(uint32_t address, const uint8_t *afh_map, int aliased, uint32_t known6, int n_obs,
		   const int32_t *indices, const uint8_t *channels, int32_t *counts, uint32_t *cands, int max)
{
	btbb_piconet *pn = btbb_piconet_new();
	int i, j, saved, devnull, n = 0;
	pn->LAP = address & 0xffffff;
	pn->UAP = (address >> 24) & 0xff;
	if (afh_map) {
		btbb_piconet_set_flag(pn, BTBB_IS_AFH, 1);
		for (i = 0; i < 10; i++) pn->afh_map[i] = afh_map[i];
		pn->used_channels = 0;
		for (i = 0; i < 79; i++) pn->used_channels += (afh_map[i / 8] >> (i % 8)) & 1;
	} else
		pn->used_channels = 79;
	pn->clk_offset = (int)(known6 & 0x3f);
	pn->first_pkt_time = 0;
	/* init_candidates / channel_winnow read pn->aliased (:463, :583), which nothing in the library ever
	 * sets (btbb_init_hop_reversal only records the BTBB_IS_ALIASED flag, :494): a receiver that aliases
	 * has to set the field itself, as done here */
	pn->aliased = aliased;
	for (j = 0; j < n_obs && j < MAX_PATTERN_LENGTH; j++) {
		pn->pattern_indices[j] = indices[j];
		pn->pattern_channels[j] = channels[j];
		counts[j] = -1;
	}
	fflush(stdout);
	saved = dup(1); devnull = open("/dev/null", O_WRONLY); dup2(devnull, 1);
	pn->packets_observed = 1;
	btbb_init_hop_reversal(aliased, pn);
	for (j = 0; j < n_obs; j++) {
		pn->packets_observed = j + 1;
		n = btbb_winnow(pn);
		counts[j] = n;
		if (n <= 1) break;
	}
	fflush(stdout);
	dup2(saved, 1); close(saved); close(devnull);
	if (n >= 1 && btbb_piconet_get_flag(pn, BTBB_HOP_REVERSAL_INIT))
		for (i = 0; i < n && i < max; i++) cands[i] = pn->clock_candidates[i];
	return n;
}
