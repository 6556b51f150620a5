/*
 * btbb.h -- the classic libbtbb BR/EDR packet API, as exported by the B200 build of
 * libbtbb.so.1 (hot-path subset, SURVEY.md section 8b "Must export").
 *
 * Function names, argument order, return conventions and flag numbers follow the
 * upstream header (lib/src/btbb.h:26-151,198) so existing callers (Ubertooth host tools,
 * gr-bluetooth, lib/src/bluetooth_piconet.c, lib/src/pcap*.c) compile and link
 * unchanged.  Piconet, pcap/pcapng and LE entry points (btbb.h:161-281) are not part of
 * this build; INTEGRATION.md shows how to add the reference's own C files for them.
 *
 * Behavioural differences, all deliberate:
 *   - btbb_init() also creates the CUDA context on device $BTBB_B200_DEVICE (default 0);
 *     it returns a negative value when no sm_100 GPU is usable (there is no CPU path).
 *   - btbb_find_ac() and btbb_decode*() launch GPU kernels on every call.
 */
#ifndef BTBB_B200_COMPAT_BTBB_H
#define BTBB_B200_COMPAT_BTBB_H

#include <stdint.h>

/* packet flag numbers (btbb.h:28-36) */
#define BTBB_WHITENED    0
#define BTBB_NAP_VALID   1
#define BTBB_UAP_VALID   2
#define BTBB_LAP_VALID   3
#define BTBB_CLK6_VALID  4
#define BTBB_CLK27_VALID 5
#define BTBB_CRC_CORRECT 6
#define BTBB_HAS_PAYLOAD 7
#define BTBB_IS_EDR      8

#define BTBB_MOD_GFSK            0x00
#define BTBB_MOD_PI_OVER_2_DQPSK 0x01
#define BTBB_MOD_8DPSK           0x02

#define BTBB_TRANSPORT_ANY  0x00
#define BTBB_TRANSPORT_SCO  0x01
#define BTBB_TRANSPORT_ESCO 0x02
#define BTBB_TRANSPORT_ACL  0x03
#define BTBB_TRANSPORT_CSB  0x04

#define LAP_ANY 0xffffffffUL
#define UAP_ANY 0xff

#ifdef __cplusplus
extern "C" {
#endif

typedef struct btbb_packet btbb_packet;

int btbb_init(int max_ac_errors);
const char *btbb_get_release(void);
const char *btbb_get_version(void);

btbb_packet *btbb_packet_new(void);
void btbb_packet_ref(btbb_packet *pkt);
void btbb_packet_unref(btbb_packet *pkt);

/* stream must hold search_length + 72 symbols (one byte each, 0/1); returns the offset of
 * the first access code within max_ac_errors, or a negative number */
int btbb_find_ac(char *stream, int search_length, uint32_t lap, int max_ac_errors, btbb_packet **pkt);

void btbb_packet_set_flag(btbb_packet *pkt, int flag, int val);
int btbb_packet_get_flag(const btbb_packet *pkt, int flag);

uint32_t btbb_packet_get_lap(const btbb_packet *pkt);
void btbb_packet_set_uap(btbb_packet *pkt, uint8_t uap);
uint8_t btbb_packet_get_uap(const btbb_packet *pkt);
uint16_t btbb_packet_get_nap(const btbb_packet *pkt);

void btbb_packet_set_modulation(btbb_packet *pkt, uint8_t modulation);
void btbb_packet_set_transport(btbb_packet *pkt, uint8_t transport);
uint8_t btbb_packet_get_modulation(const btbb_packet *pkt);
uint8_t btbb_packet_get_transport(const btbb_packet *pkt);

uint8_t btbb_packet_get_channel(const btbb_packet *pkt);
uint8_t btbb_packet_get_ac_errors(const btbb_packet *pkt);
uint32_t btbb_packet_get_clkn(const btbb_packet *pkt);
uint32_t btbb_packet_get_header_packed(const btbb_packet *pkt);

void btbb_packet_set_data(btbb_packet *pkt, char *syms, int length, uint8_t channel, uint32_t clkn);

const char *btbb_get_symbols(const btbb_packet *pkt);
int btbb_packet_get_payload_length(const btbb_packet *pkt);
const char *btbb_get_payload(const btbb_packet *pkt);
int btbb_get_payload_packed(const btbb_packet *pkt, char *dst);

uint8_t btbb_packet_get_type(const btbb_packet *pkt);
uint8_t btbb_packet_get_lt_addr(const btbb_packet *pkt);
uint8_t btbb_packet_get_header_flags(const btbb_packet *pkt);
uint8_t btbb_packet_get_hec(const btbb_packet *pkt);

uint64_t btbb_gen_syncword(const int LAP);

int btbb_decode_header(btbb_packet *pkt);
int btbb_decode_payload(btbb_packet *pkt);
int btbb_decode(btbb_packet *pkt);
void btbb_print_packet(const btbb_packet *pkt);
int btbb_header_present(const btbb_packet *pkt);

#ifdef __cplusplus
}
#endif
#endif
