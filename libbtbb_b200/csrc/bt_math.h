/*
 * bt_math.h -- small GF(2) building blocks shared by the host and device halves of the
 * B200 path (sync-word code, HEC, CRC, FEC 2/3 parity, whitening LFSR).  Everything is
 * derived from the generator polynomials of the Bluetooth baseband spec; no table is
 * taken from the reference sources (which ship them precomputed: sw_matrix and
 * sw_check_tables.h, fec23_gen_matrix, WHITENING_DATA/INDICES in
 * lib/src/bluetooth_packet.c:49-119).
 */
#ifndef BTBB_B200_BT_MATH_H
#define BTBB_B200_BT_MATH_H

#include <stdint.h>

#ifdef __CUDACC__
#define BT_HD __host__ __device__ __forceinline__
#else
#define BT_HD static inline
#endif

#define BT_G34      0x585713DA9ULL          /* (64,30) code generator, octal 0260534236651; bit i <-> x^i */
#define BT_PN       0x83848D96BBCC54FCULL   /* PN overlay of the sync word */
#define BT_BARKER_A 0x27u                   /* bits 57..63 of a sync word whose LAP bit 23 is 0 */
#define BT_BARKER_B 0x58u                   /* ... LAP bit 23 is 1 (complement) */
#define BT_MAX_SYMBOLS 3125                 /* bluetooth_packet.h:27 */

/* codeword mod g(x): what gen_syndrome (bluetooth_packet.c:147-159) computes via 4 byte LUTs */
BT_HD uint64_t bt_syndrome_slow(uint64_t cw)
{
	for (int i = 63; i >= 34; i--)
		if ((cw >> i) & 1)
			cw ^= BT_G34 << (i - 34);
	return cw;
}

/* btbb_gen_syncword (bluetooth_packet.c:188-199) from the spec construction */
BT_HD uint64_t bt_gen_syncword(uint32_t lap)
{
	uint64_t info = (uint64_t)(lap & 0xffffffu);
	info |= (uint64_t)((lap & 0x800000u) ? 0x13u : 0x2cu) << 24;
	info ^= BT_PN >> 34;
	uint64_t cw = info << 34;
	cw |= bt_syndrome_slow(cw);
	return cw ^ BT_PN;
}

BT_HD uint32_t bt_rev8(uint32_t b)
{
	b = ((b & 0xf0u) >> 4) | ((b & 0x0fu) << 4);
	b = ((b & 0xccu) >> 2) | ((b & 0x33u) << 2);
	b = ((b & 0xaau) >> 1) | ((b & 0x55u) << 1);
	return b;
}

/* forward HEC: reflected LFSR of D^8+D^7+D^5+D^2+D+1 preloaded with the reversed UAP */
BT_HD uint32_t bt_hec(uint32_t data10, uint32_t uap)
{
	uint32_t reg = bt_rev8(uap & 0xffu);
	for (int i = 0; i < 10; i++) {
		uint32_t fb = (reg ^ (data10 >> i)) & 1u;
		reg >>= 1;
		if (fb) reg ^= 0xE5u;
	}
	return reg;
}

/* uap_from_hec (bluetooth_packet.c:693-705): the same LFSR run backwards */
BT_HD uint32_t bt_uap_from_hec(uint32_t data10, uint32_t hec)
{
	uint32_t reg = hec & 0xffu;
	for (int i = 9; i >= 0; i--) {
		uint32_t fb = reg >> 7;
		if (fb) reg ^= 0xE5u;
		reg = ((reg << 1) & 0xffu) | (fb ^ ((data10 >> i) & 1u));
	}
	return bt_rev8(reg);
}

/* one CRC-16/CCITT (reflected, 0x8408) step; crcgen (bluetooth_packet.c:671-690) */
BT_HD uint32_t bt_crc16_step(uint32_t reg, uint32_t bit)
{
	uint32_t fb = (reg ^ bit) & 1u;
	reg >>= 1;
	return fb ? reg ^ 0x8408u : reg;
}
BT_HD uint32_t bt_crc16_init(uint32_t uap) { return bt_rev8(uap & 0xffu) << 8; }

/* 5 parity bits of the (15,10) shortened Hamming code, g(D)=D^5+D^4+D^2+1
 * (fec23, bluetooth_packet.c:571-582); returned in air order (bit 0 = first sent) */
BT_HD uint32_t bt_fec23_parity(uint32_t data10)
{
	uint32_t reg = 0;
	for (int i = 0; i < 10; i++) {
		uint32_t fb = ((reg >> 4) ^ (data10 >> i)) & 1u;
		reg = (reg << 1) & 0x1fu;
		if (fb) reg ^= 0x15u;
	}
	uint32_t par = 0;
	for (int i = 0; i < 5; i++)
		par |= ((reg >> (4 - i)) & 1u) << i;
	return par;
}

/* whitening LFSR x^7+x^4+1, state seeded 1,CLK6..CLK1 (unwhiten, bluetooth_packet.c:653-668) */
BT_HD uint32_t bt_whiten_seed(uint32_t clk) { return 0x40u | (clk & 0x3fu); }
BT_HD uint32_t bt_whiten_step(uint32_t *s)
{
	uint32_t o = (*s >> 6) & 1u;
	*s = (*s << 1) & 0x7fu;
	if (o) *s ^= 0x11u;
	return o;
}

BT_HD uint64_t bt_splitmix64(uint64_t x)
{
	x += 0x9E3779B97F4A7C15ULL;
	x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
	x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
	return x ^ (x >> 31);
}

#endif
