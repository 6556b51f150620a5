/*
 * compat.cu -- the classic btbb_* surface (include/btbb.h; upstream lib/src/btbb.h:63-151,198
 * and the exported helpers of lib/src/bluetooth_packet.h:115-144) on top of the batch C ABI.
 *
 * struct btbb_packet keeps the upstream field order and sizes (bluetooth_packet.h:52-112,
 * 5952 bytes) because bluetooth_piconet.c and the pcap writers poke at it directly when
 * they are compiled into the same library.  Accessors are plain host C.
 * Routing (btbb_b200_classic_config): a classic call carries ONE packet or one short buffer, so by
 * default single-packet calls and searches of at most 8192 positions are answered by the host
 * small-call path (decode_host.cpp / find_ac_host.cpp: the kernels' own arithmetic from decode_core.h
 * and bt_math.h compiled for the host -- never oracle/), longer searches by the kernels; with
 * BTBB_B200_CLASSIC=gpu every call launches kernels.  The batch entry points never take the host path.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <mutex>
#include <shared_mutex>
#include "../../include/btbb.h"
#include "bt_math.h"
#include "capi_internal.h"

#define MAX_SYMBOLS 3125
#define MAX_PAYLOAD_LENGTH 2744

struct btbb_packet {
	uint32_t refcount;
	uint32_t flags;
	uint8_t channel;
	uint8_t UAP;
	uint16_t NAP;
	uint32_t LAP;
	uint8_t modulation;
	uint8_t transport;
	uint8_t packet_type;
	uint8_t packet_lt_addr;
	uint8_t packet_flags;
	uint8_t packet_hec;
	char packet_header[18];
	int payload_header_length;
	char payload_header[16];
	uint8_t payload_llid;
	uint8_t payload_flow;
	int payload_length;
	char payload[MAX_PAYLOAD_LENGTH];
	uint16_t crc;
	uint32_t clkn;
	uint8_t ac_errors;
	uint16_t length;
	char symbols[MAX_SYMBOLS];
};
static_assert(sizeof(btbb_packet) == 5952, "btbb_packet must keep the upstream layout");

/* ---- library state ----
 * g_state guards the context pointer: every shim call holds it shared for its duration, btbb_init
 * takes it exclusively when it has to (re)build the tables (the reference builds its syndrome map
 * once, for the first non-zero k, :288-289).  GPU-path calls additionally serialise on g_gpu, because
 * they share the context's device scratch.  The host small-call path is re-entrant, like the
 * reference after btbb_init. */
static std::shared_mutex g_state;
static std::mutex g_gpu;
static btbb_b200_ctx *g_ctx;
/* routing of the classic calls (btbb_b200_classic_config): searches of at most this many positions
 * and all single-packet calls are answered on the host (find_ac_host.cpp, decode_host.cpp) */
static std::atomic<int> g_host_below{-2};      /* -2: not configured yet (read BTBB_B200_CLASSIC once) */
static std::atomic<int> g_packet_gpu{0};

static int env_device(void)
{
	const char *e = getenv("BTBB_B200_DEVICE");
	return e ? atoi(e) : 0;
}

static void route_init(void)
{
	if (g_host_below.load() != -2) return;
	const char *e = getenv("BTBB_B200_CLASSIC");      /* "gpu": every classic call launches kernels */
	if (e && !strcmp(e, "gpu")) { g_packet_gpu.store(1); g_host_below.store(-1); }
	else g_host_below.store(8192);
}

/* caller holds g_state (shared); creates the context on first use */
static btbb_b200_ctx *ctx_locked(std::shared_lock<std::shared_mutex> &lk, int k_if_new)
{
	if (g_ctx) return g_ctx;
	lk.unlock();
	{
		std::unique_lock<std::shared_mutex> w(g_state);
		if (!g_ctx && btbb_b200_create(env_device(), k_if_new, &g_ctx) != BTBB_B200_OK) {
			fprintf(stderr, "libbtbb (B200): %s\n", btbb_b200_last_error());
			g_ctx = NULL;
		}
	}
	lk.lock();
	return g_ctx;
}

extern "C" {

/* find_ac_host_below: btbb_find_ac calls searching at most this many positions run on the host
 * (-1: never); packet_calls_on_gpu != 0: btbb_decode* / try_clock / crc_check / btbb_header_present
 * launch kernels instead of taking the host small-call path */
void btbb_b200_classic_config(int find_ac_host_below, int packet_calls_on_gpu)
{
	g_host_below.store(find_ac_host_below < -1 ? -1 : find_ac_host_below);
	g_packet_gpu.store(packet_calls_on_gpu != 0);
}

int btbb_init(int max_ac_errors)
{
	if (max_ac_errors < 0 || max_ac_errors > 5) {
		fprintf(stderr, "%s: max_ac_errors out of range\n", __FUNCTION__);
		return -1;
	}
	route_init();
	std::unique_lock<std::shared_mutex> w(g_state);
	if (g_ctx && btbb_b200_table_errors(g_ctx) == 0 && max_ac_errors > 0) {
		std::lock_guard<std::mutex> g(g_gpu);
		btbb_b200_destroy(g_ctx);
		g_ctx = NULL;
	}
	if (!g_ctx && btbb_b200_create(env_device(), max_ac_errors, &g_ctx) != BTBB_B200_OK) {
		fprintf(stderr, "libbtbb (B200): %s\n", btbb_b200_last_error());
		g_ctx = NULL;
	}
	return g_ctx ? 0 : -2;
}

#ifndef BTBB_B200_RELEASE
#define BTBB_B200_RELEASE "b200-r2"
#endif
const char *btbb_get_release(void) { return BTBB_B200_RELEASE; }
const char *btbb_get_version(void) { return "libbtbb-b200 0.1 (sm_100a)"; }

btbb_packet *btbb_packet_new(void)
{
	btbb_packet *p = (btbb_packet *)calloc(1, sizeof(btbb_packet));
	if (p) p->refcount = 1;
	else fprintf(stderr, "Unable to allocate packet");
	return p;
}
void btbb_packet_ref(btbb_packet *p) { p->refcount++; }
void btbb_packet_unref(btbb_packet *p) { if (--p->refcount == 0) free(p); }

void btbb_packet_set_flag(btbb_packet *p, int flag, int val)
{
	uint32_t m = 1u << flag;
	p->flags = val ? (p->flags | m) : (p->flags & ~m);
}
int btbb_packet_get_flag(const btbb_packet *p, int flag) { return (p->flags >> flag) & 1u; }

uint32_t btbb_packet_get_lap(const btbb_packet *p) { return p->LAP; }
void btbb_packet_set_uap(btbb_packet *p, uint8_t uap) { p->UAP = uap; btbb_packet_set_flag(p, BTBB_UAP_VALID, 1); }
uint8_t btbb_packet_get_uap(const btbb_packet *p) { return p->UAP; }
uint16_t btbb_packet_get_nap(const btbb_packet *p) { return p->NAP; }
uint32_t btbb_packet_get_clkn(const btbb_packet *p) { return p->clkn; }
uint8_t btbb_packet_get_channel(const btbb_packet *p) { return p->channel; }
void btbb_packet_set_modulation(btbb_packet *p, uint8_t m) { p->modulation = m; }
uint8_t btbb_packet_get_modulation(const btbb_packet *p) { return p->modulation; }
void btbb_packet_set_transport(btbb_packet *p, uint8_t t) { p->transport = t; }
uint8_t btbb_packet_get_transport(const btbb_packet *p) { return p->transport; }
uint8_t btbb_packet_get_ac_errors(const btbb_packet *p) { return p->ac_errors; }
const char *btbb_get_symbols(const btbb_packet *p) { return p->symbols; }
int btbb_packet_get_payload_length(const btbb_packet *p) { return p->payload_length; }
const char *btbb_get_payload(const btbb_packet *p) { return p->payload; }
uint8_t btbb_packet_get_type(const btbb_packet *p) { return p->packet_type; }
uint8_t btbb_packet_get_lt_addr(const btbb_packet *p) { return p->packet_lt_addr; }
uint8_t btbb_packet_get_header_flags(const btbb_packet *p) { return p->packet_flags; }
uint8_t btbb_packet_get_hec(const btbb_packet *p) { return p->packet_hec; }

static uint32_t pack_bits(const char *b, int n)
{
	uint32_t v = 0;
	for (int i = 0; i < n; i++) v |= (uint32_t)(b[i] & 1) << i;
	return v;
}

uint32_t btbb_packet_get_header_packed(const btbb_packet *p) { return pack_bits(p->packet_header, 18); }

int btbb_get_payload_packed(const btbb_packet *p, char *dst)
{
	for (int i = 0; i < p->payload_length; i++)
		dst[i] = (char)pack_bits(&p->payload[i * 8], 8);
	return p->payload_length;
}

void btbb_packet_set_data(btbb_packet *p, char *data, int length, uint8_t channel, uint32_t clkn)
{
	if (length > MAX_SYMBOLS) length = MAX_SYMBOLS;
	if (length > 0) memcpy(p->symbols, data, (size_t)length);
	p->length = (uint16_t)length;
	p->channel = channel;
	p->clkn = clkn >> 1;      /* stored as CLK1.. (:479) */
}

uint64_t btbb_gen_syncword(const int LAP) { return bt_gen_syncword((uint32_t)LAP); }

/* ---- access-code search ---- */
static int find_first(char *stream, int search_length, uint32_t lap, int k, uint32_t *lap_out, uint8_t *ac_errors)
{
	route_init();
	std::shared_lock<std::shared_mutex> lk(g_state);
	btbb_b200_ctx *ctx = ctx_locked(lk, 0);
	btbb_b200_hit h;
	int found = 0, rc;
	if (!ctx) return -2;
	if (search_length <= g_host_below.load())
		rc = bt_find_first_cpu(ctx, stream, search_length, lap, k, &h, &found);
	else {
		std::lock_guard<std::mutex> g(g_gpu);
		rc = bt_find_first_host(ctx, stream, search_length, lap, k, &h, &found);
	}
	if (rc != BTBB_B200_OK) {
		fprintf(stderr, "libbtbb (B200): %s\n", btbb_b200_last_error());
		return -2;
	}
	if (!found) return -1;
	*lap_out = h.lap; *ac_errors = h.ac_errors;
	return (int)h.offset;
}

int promiscuous_packet_search(char *stream, int search_length, uint32_t *lap, int max_ac_errors, uint8_t *ac_errors)
{
	return find_first(stream, search_length, BTBB_B200_LAP_ANY, max_ac_errors, lap, ac_errors);
}

int find_known_lap(char *stream, int search_length, uint32_t lap, int max_ac_errors, uint8_t *ac_errors)
{
	uint32_t l;
	return find_first(stream, search_length, lap & 0xffffffu, max_ac_errors, &l, ac_errors);
}

int btbb_find_ac(char *stream, int search_length, uint32_t lap, int max_ac_errors, btbb_packet **pkt_ptr)
{
	uint8_t ne = 0;
	int off;
	if (lap == (uint32_t)LAP_ANY)
		off = promiscuous_packet_search(stream, search_length, &lap, max_ac_errors, &ne);
	else
		off = find_known_lap(stream, search_length, lap, max_ac_errors, &ne);
	if (off >= 0) {
		if (*pkt_ptr == NULL) *pkt_ptr = btbb_packet_new();
		(*pkt_ptr)->LAP = lap;
		(*pkt_ptr)->ac_errors = ne;
		(*pkt_ptr)->flags = 0;
		btbb_packet_set_flag(*pkt_ptr, BTBB_WHITENED, 1);
	}
	return off;
}

/* ---- per-packet chain: the host small-call path, or one GPU launch per call ---- */
static int run_chain(btbb_packet *p, int mode, int clock, btbb_b200_decoded *rec)
{
	route_init();
	const int whitened = btbb_packet_get_flag(p, BTBB_WHITENED);
	if (!g_packet_gpu.load())
		return bt_decode_one_cpu(p->symbols, p->length, (uint32_t)clock, p->UAP, whitened, p->packet_type, mode, rec) ? -1 : 0;
	std::shared_lock<std::shared_mutex> lk(g_state);
	btbb_b200_ctx *ctx = ctx_locked(lk, 0);
	btbb_b200_pkt_in in;
	if (!ctx) return -1;
	memset(&in, 0, sizeof(in));
	in.offset = 0; in.length = p->length; in.clkn = (uint32_t)clock;
	in.uap = p->UAP; in.whitened = (uint8_t)whitened;
	in.type = p->packet_type;
	if (btbb_b200_decode_host(ctx, p->symbols, MAX_SYMBOLS, &in, 1, mode, rec) != BTBB_B200_OK) {
		fprintf(stderr, "libbtbb (B200): %s\n", btbb_b200_last_error());
		return -1;
	}
	return 0;
}

/* unfec13 of the 18 header triplets (:552-568): the header is "there" iff fewer than 18 / 4 disagree */
static int header_fec_ok(const btbb_packet *p)
{
	int bad = 0;
	for (int i = 0; i < 18; i++) {
		const int t = (p->symbols[68 + 3 * i] & 1) + (p->symbols[69 + 3 * i] & 1) + (p->symbols[70 + 3 * i] & 1);
		bad += (t == 1 || t == 2);
	}
	return bad < 18 / 4;
}

static void store_payload(btbb_packet *p, const btbb_b200_decoded *r)
{
	p->payload_header_length = r->payload_header_length;
	p->payload_length = r->payload_length;
	p->payload_llid = r->llid;
	p->payload_flow = r->flow;
	if (r->has_payload) btbb_packet_set_flag(p, BTBB_HAS_PAYLOAD, 1);
	/* the record was asked for with BTBB_B200_MODE_FLAG_RAW_PAYLOAD: it carries what the reference's
	 * decoders leave in pkt->payload whatever rv says (bits they did not write read 0, as in a fresh packet) */
	if (r->payload_length > 0 && r->payload_length <= 344)
		for (int i = 0; i < r->payload_length * 8; i++)
			p->payload[i] = (char)((r->payload[i >> 3] >> (i & 7)) & 1);
}

uint8_t try_clock(int clock, btbb_packet *p)
{
	route_init();
	if (!g_packet_gpu.load()) {
		uint8_t uap, type;
		if (!bt_try_clock_cpu(p->symbols, p->length, clock, btbb_packet_get_flag(p, BTBB_WHITENED), &uap, &type))
			return 0;             /* unfec13 failed: packet untouched (:1186-1187) */
		p->UAP = uap;
		p->packet_type = type;
		return p->UAP;
	}
	static thread_local btbb_b200_decoded rec[64];
	if (run_chain(p, BTBB_B200_MODE_TRY_CLOCKS, 0, rec)) return 0;
	const btbb_b200_decoded *r = &rec[clock & 63];
	if (!r->header_ok) return 0;          /* unfec13 failed: packet untouched (:1186-1187) */
	p->UAP = r->uap;
	p->packet_type = r->type;
	return p->UAP;
}

static int typed(int mode, int clock, btbb_packet *p)
{
	btbb_b200_decoded r;
	if (run_chain(p, mode | BTBB_B200_MODE_FLAG_RAW_PAYLOAD, clock, &r)) return 0;
	store_payload(p, &r);
	return r.rv;
}

int crc_check(int clock, btbb_packet *p) { return typed(BTBB_B200_MODE_CRC_CHECK, clock, p); }
int fhs(int clock, btbb_packet *p) { return typed(BTBB_B200_MODE_RAW + 0, clock, p); }
int DM(int clock, btbb_packet *p)  { return typed(BTBB_B200_MODE_RAW + 1, clock, p); }
int DH(int clock, btbb_packet *p)  { return typed(BTBB_B200_MODE_RAW + 2, clock, p); }
int EV3(int clock, btbb_packet *p) { return typed(BTBB_B200_MODE_RAW + 3, clock, p); }
int EV4(int clock, btbb_packet *p) { return typed(BTBB_B200_MODE_RAW + 4, clock, p); }
int EV5(int clock, btbb_packet *p) { return typed(BTBB_B200_MODE_RAW + 5, clock, p); }
int HV(int clock, btbb_packet *p)  { return typed(BTBB_B200_MODE_RAW + 6, clock, p); }

int btbb_decode_header(btbb_packet *p)
{
	btbb_b200_decoded r;
	if (!btbb_packet_get_flag(p, BTBB_CLK6_VALID)) return 0;
	if (run_chain(p, BTBB_B200_MODE_DECODE, (int)p->clkn, &r)) return 0;
	if (header_fec_ok(p))      /* the reference writes packet_header once unfec13 has passed, whatever the HEC says (:1203-1209) */
		for (int i = 0; i < 18; i++) p->packet_header[i] = (char)((r.header_packed >> i) & 1);
	if (!r.header_ok) return 0;
	p->packet_lt_addr = r.lt_addr; p->packet_type = r.type;
	p->packet_flags = r.flags; p->packet_hec = r.hec;
	return 1;
}

int btbb_decode_payload(btbb_packet *p)
{
	btbb_b200_decoded r;
	if (run_chain(p, BTBB_B200_MODE_PAYLOAD | BTBB_B200_MODE_FLAG_RAW_PAYLOAD, (int)p->clkn, &r)) return 0;
	store_payload(p, &r);
	btbb_packet_set_flag(p, BTBB_HAS_PAYLOAD, 1);
	return r.rv;
}

static const char *const type_names[16] = {
	"NULL", "POLL", "FHS", "DM1", "DH1/2-DH1", "HV1", "HV2/2-EV3", "HV3/EV3/3-EV3",
	"DV/3-DH1", "AUX1", "DM3/2-DH3", "DH3/3-DH3", "EV4/2-EV5", "EV5/3-EV5", "DM5/2-DH5", "DH5/3-DH5"};

void btbb_print_packet(const btbb_packet *p)
{
	if (!btbb_packet_get_flag(p, BTBB_HAS_PAYLOAD)) return;
	printf("  Type: %s\n", type_names[p->packet_type & 15]);
	if (p->payload_header_length > 0) {
		printf("  LT_ADDR: %d\n", p->packet_lt_addr);
		printf("  LLID: %d\n", p->payload_llid);
		printf("  flow: %d\n", p->payload_flow);
		printf("  payload length: %d\n", p->payload_length);
	}
	if (p->payload_length) {
		printf("  Data: ");
		for (int i = 0; i < p->payload_length; i++)
			printf(" %02x", pack_bits(p->payload + 8 * i, 8));
		printf("\n");
	}
}

int btbb_decode(btbb_packet *p)
{
	int rv = 0;
	btbb_packet_set_flag(p, BTBB_HAS_PAYLOAD, 0);
	if (btbb_decode_header(p))
		rv = btbb_decode_payload(p);
	if (rv > 0) {     /* the reference reports on stdout (:1311-1314) */
		printf("Packet decoded with clock 0x%02x (rv=%d)\n", p->clkn & 0x3f, rv);
		btbb_print_packet(p);
	}
	return rv;
}

int btbb_header_present(const btbb_packet *p)
{
	route_init();
	if (!g_packet_gpu.load())
		return bt_header_present_cpu(p->symbols, p->length);
	std::shared_lock<std::shared_mutex> lk(g_state);
	btbb_b200_ctx *ctx = ctx_locked(lk, 0);
	if (!ctx) return 0;
	std::lock_guard<std::mutex> g(g_gpu);
	/* single packet through the batch kernel; the device scratch is the context's own */
	btbb_b200_pkt_in in;
	uint8_t res = 0;
	memset(&in, 0, sizeof(in));
	in.length = p->length;
	return btbb_b200_header_present_host(ctx, p->symbols, MAX_SYMBOLS, &in, 1, &res) == BTBB_B200_OK ? res : 0;
}

char *tun_format(btbb_packet *p)
{
	int length = 9 + p->payload_length;
	char *t = (char *)malloc((size_t)length);
	if (!t) return NULL;
	t[0] = (char)(p->clkn & 0xff); t[1] = (char)((p->clkn >> 8) & 0xff);
	t[2] = (char)((p->clkn >> 16) & 0xff); t[3] = (char)((p->clkn >> 24) & 0xff);
	t[4] = (char)p->channel;
	t[5] = (char)(btbb_packet_get_flag(p, BTBB_CLK27_VALID) | (btbb_packet_get_flag(p, BTBB_NAP_VALID) << 1));
	t[6] = (char)pack_bits(&p->packet_header[0], 7);
	t[7] = (char)pack_bits(&p->packet_header[7], 3);
	t[8] = (char)pack_bits(&p->packet_header[10], 8);
	for (int i = 0; i < p->payload_length; i++)
		t[i + 9] = (char)pack_bits(&p->payload[i * 8], 8);
	return t;
}

uint32_t lap_from_fhs(btbb_packet *p) { return pack_bits(&p->payload[34], 24); }
uint8_t uap_from_fhs(btbb_packet *p) { return (uint8_t)pack_bits(&p->payload[64], 8); }
uint16_t nap_from_fhs(btbb_packet *p) { return (uint16_t)pack_bits(&p->payload[72], 16); }
uint32_t clock_from_fhs(btbb_packet *p) { return pack_bits(&p->payload[115], 26); }

}  /* extern "C" */
