/*
 * oracle/oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the libbtbb hot path used as the
 * checker for the CUDA kernels.  Nothing in libbtbb_b200/ may include, link or
 * call this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg do.  Parity of this restatement with the unmodified
 * reference is pinned by tests/test_oracle_vs_golden.py (golden vectors of the
 * reference's own tests + fixtures generated from oracle/_ref/libbtbb_ref.so).
 */
#ifndef BTBB_ORACLE_H
#define BTBB_ORACLE_H
#include <stdint.h>
#include "../include/btbb_b200.h"   /* record layouts only */

uint64_t orc_syndrome(uint64_t codeword);
uint64_t orc_gen_syncword(uint32_t lap);
int      orc_barker_distance(int b7);
uint64_t orc_barker_correct(int b7);
int      orc_init(int max_ac_errors);            /* -1 outside 0..5, like btbb_init */
int      orc_table_errors(void);
long     orc_table_entries(void);
int      orc_lookup_error(uint64_t syndrome, uint64_t *error);

int64_t  orc_find_all(const char *stream, int64_t search_length, uint32_t lap,
		      int max_ac_errors, btbb_b200_hit *hits, int64_t max_hits);
double   orc_find_all_mt(const char *stream, int64_t search_length, uint32_t lap,
			 int max_ac_errors, int threads, int64_t *total_hits);

int      orc_unfec13(const char *in, char *out, int length);
uint16_t orc_fec23(uint16_t data10);
int      orc_unfec23(const char *in, int length, char *out);
int      orc_whiten_bit(int clk6, int position);
void     orc_unwhiten(const char *in, char *out, int clock, int length, int skip, int whitened);
uint16_t orc_crc16(const char *bits, int length, int uap);
uint8_t  orc_hec(uint16_t data10, uint8_t uap);
uint8_t  orc_uap_from_hec(uint16_t data10, uint8_t hec);
int      orc_header_present(const char *symbols, int length);

void     orc_decode_one(const char *symbols, int length, uint32_t clkn, uint8_t uap,
			int whitened, btbb_b200_decoded *out);
void     orc_decode_one_raw(const char *symbols, int length, uint32_t clkn, uint8_t uap,
			    int whitened, btbb_b200_decoded *out);
void     orc_try_clock_one(const char *symbols, int length, int clock, int whitened,
			   btbb_b200_decoded *out);
void     orc_uap_sieve(const char *stream, int64_t stream_length, const btbb_b200_pkt_in *pkts, int64_t n_pkts,
		       const int64_t *group_start, int64_t n_groups, btbb_b200_sieve *states, int8_t *rv);
void     orc_hop_sequence(uint32_t address, const uint8_t *afh_map, int64_t first, int64_t n, uint8_t *out);
#endif
