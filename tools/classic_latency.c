/*
 * classic_latency.c -- per-call latency of the classic libbtbb calls (include/btbb.h) for ANY libbtbb
 * build: the unmodified reference (oracle/_ref/libbtbb_ref.so) or the product (lib/libbtbb.so.1).
 * The library is dlopen()ed, so one binary times both through exactly the same calls.
 *
 *   classic_latency <libbtbb.so> <packets.bin> <reps> [find_ac]
 *
 * packets.bin (written by tools/classic_latency.py): int32 count, then per packet int32 length,
 * int32 clk6, int32 uap, int32 ac_offset and 4096 symbols (one byte each; the packet's access code
 * starts at ac_offset, the symbols handed to btbb_packet_set_data start there).
 * Prints one JSON object: microseconds per call (mean over reps x packets) and a checksum of the
 * results, which must be equal for two libraries that compute the same thing.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct btbb_packet btbb_packet;
#define SYMS 4096

static int (*p_init)(int);
static btbb_packet *(*p_new)(void);
static void (*p_unref)(btbb_packet *);
static void (*p_set_flag)(btbb_packet *, int, int);
static void (*p_set_data)(btbb_packet *, char *, int, uint8_t, uint32_t);
static void (*p_set_uap)(btbb_packet *, uint8_t);
static int (*p_decode_header)(btbb_packet *);
static int (*p_decode_payload)(btbb_packet *);
static int (*p_payload_length)(const btbb_packet *);
static uint8_t (*p_try_clock)(int, btbb_packet *);
static int (*p_crc_check)(int, btbb_packet *);
static int (*p_find_ac)(char *, int, uint32_t, int, btbb_packet **);
static uint32_t (*p_get_lap)(const btbb_packet *);

struct pk { int32_t length, clk, uap, ac_offset; char sym[SYMS]; };

static double now_us(void)
{
	struct timespec t;
	clock_gettime(CLOCK_MONOTONIC, &t);
	return t.tv_sec * 1e6 + t.tv_nsec * 1e-3;
}

#define SYM(var, name) do { *(void **)&var = dlsym(h, name); if (!var) { fprintf(stderr, "missing %s\n", name); return 2; } } while (0)

int main(int argc, char **argv)
{
	if (argc < 4) { fprintf(stderr, "usage: %s lib packets.bin reps [find_ac]\n", argv[0]); return 2; }
	void *h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
	if (!h) { fprintf(stderr, "%s\n", dlerror()); return 2; }
	SYM(p_init, "btbb_init"); SYM(p_new, "btbb_packet_new"); SYM(p_unref, "btbb_packet_unref");
	SYM(p_set_flag, "btbb_packet_set_flag"); SYM(p_set_data, "btbb_packet_set_data"); SYM(p_set_uap, "btbb_packet_set_uap");
	SYM(p_decode_header, "btbb_decode_header"); SYM(p_decode_payload, "btbb_decode_payload");
	SYM(p_payload_length, "btbb_packet_get_payload_length"); SYM(p_try_clock, "try_clock"); SYM(p_crc_check, "crc_check");
	SYM(p_find_ac, "btbb_find_ac"); SYM(p_get_lap, "btbb_packet_get_lap");
	const int reps = atoi(argv[3]), with_find = argc > 4 && !strcmp(argv[4], "find_ac");

	FILE *f = fopen(argv[2], "rb");
	int32_t n = 0;
	if (!f || fread(&n, 4, 1, f) != 1 || n <= 0) { fprintf(stderr, "bad packet file\n"); return 2; }
	struct pk *pk = malloc((size_t)n * sizeof *pk);
	if (fread(pk, sizeof *pk, (size_t)n, f) != (size_t)n) { fprintf(stderr, "short packet file\n"); return 2; }
	fclose(f);

	/* 1. btbb_decode_header + btbb_decode_payload with UAP and CLK1-6 known (what btbb_decode does for
	 *    a packet of a followed piconet), including the packet object's set-up and release */
	uint64_t sum_decode = 0;
	double t0 = now_us();
	for (int r = 0; r < reps; r++)
		for (int i = 0; i < n; i++) {
			btbb_packet *p = p_new();
			p_set_flag(p, 0 /* BTBB_WHITENED */, 1);
			p_set_data(p, pk[i].sym + pk[i].ac_offset, pk[i].length, 3, (uint32_t)pk[i].clk << 1);
			p_set_uap(p, (uint8_t)pk[i].uap);
			p_set_flag(p, 4 /* BTBB_CLK6_VALID */, 1);
			int rv = -1;
			if (p_decode_header(p)) rv = p_decode_payload(p);
			sum_decode += (uint64_t)(rv + 7) * 31 + (uint64_t)p_payload_length(p);
			p_unref(p);
		}
	const double us_decode = (now_us() - t0) / ((double)reps * n);

	/* 2. the 64-clock UAP sweep of one packet (UAP_from_header's inner loop, bluetooth_piconet.c:
	 *    try_clock then crc_check for every CLK1-6) */
	uint64_t sum_sweep = 0;
	t0 = now_us();
	for (int r = 0; r < reps; r++)
		for (int i = 0; i < n; i++) {
			btbb_packet *p = p_new();
			p_set_flag(p, 0, 1);
			p_set_data(p, pk[i].sym + pk[i].ac_offset, pk[i].length, 3, 0);
			for (int c = 0; c < 64; c++) {
				const unsigned u = p_try_clock(c, p);
				const int rv = p_crc_check(c, p);
				sum_sweep += (uint64_t)(u * 64 + c) * (uint64_t)(rv + 3);
			}
			p_unref(p);
		}
	const double us_sweep = (now_us() - t0) / ((double)reps * n);

	/* 3. btbb_find_ac over a 4096-symbol buffer (what a receiver hands over per burst), first hit */
	double us_find = -1;
	uint64_t sum_find = 0;
	int init_rc = -99;
	if (with_find && (init_rc = p_init(2)) == 0) {
		t0 = now_us();
		for (int r = 0; r < reps; r++)
			for (int i = 0; i < n; i++) {
				btbb_packet *p = NULL;
				const int off = p_find_ac(pk[i].sym, SYMS - 64, 0xffffffffu, 2, &p);
				sum_find += (uint64_t)(off + 1);
				if (p) { sum_find += p_get_lap(p); p_unref(p); }
			}
		us_find = (now_us() - t0) / ((double)reps * n);
	}
	printf("{\"library\": \"%s\", \"packets\": %d, \"reps\": %d, \"decode_us\": %.3f, \"uap_sweep_64_clocks_us\": %.3f, "
	       "\"find_ac_4096_us\": %s%.3f, \"btbb_init_rc\": %d, \"checksums\": [%llu, %llu, %llu]}\n",
	       argv[1], n, reps, us_decode, us_sweep, us_find < 0 ? "" : "", us_find, init_rc,
	       (unsigned long long)sum_decode, (unsigned long long)sum_sweep, (unsigned long long)sum_find);
	return 0;
}
