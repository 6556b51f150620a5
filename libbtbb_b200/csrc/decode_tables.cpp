/*
 * decode_tables.cpp -- constant tables of the per-packet chain (decode_core.h), derived from the
 * generator polynomials in bt_math.h: the whitening sequence (unwhiten, bluetooth_packet.c:653-668)
 * in every rotation, the FEC 2/3 parity columns (fec23 :571-582) and the CRC-16 prefix weights
 * (crcgen :671-690, see decode_core.h for the algebra).  Built once per process.
 */
#include <mutex>
#include <string.h>
#include "decode_core.h"

namespace {

btd_tables g_tables;
std::once_flag g_once;

/* inverse of one zero-input CRC step: A(r) = (r >> 1) ^ (r & 1 ? 0x8408 : 0) */
uint32_t crc_unstep(uint32_t r)
{
	const uint32_t b15 = (r >> 15) & 1u;
	if (b15) r ^= 0x8408u;
	return ((r << 1) | b15) & 0xffffu;
}

void build()
{
	btd_tables &t = g_tables;
	memset(&t, 0, sizeof(t));
	/* whitening: one period from the state 0x40 (clock 0), phase[] = where each seed state occurs */
	uint8_t seq[127];
	uint32_t st = bt_whiten_seed(0);
	for (int i = 0; i < 127; i++) {
		if (st & 0x40u) t.s.phase[st & 0x3fu] = (uint8_t)i;
		seq[i] = (uint8_t)bt_whiten_step(&st);
	}
	for (int p = 0; p < 127; p++)
		for (int b = 0; b < 32; b++)
			t.s.wrot[p] |= (uint32_t)seq[(p + b) % 127] << b;
	t.s.wrot[127] = t.s.wrot[0];
	for (int i = 0; i < 10; i++) t.s.fec_col[i] = (uint8_t)bt_fec23_parity(1u << i);
	/* CRC weights u_i = A^-(i+1)(0x8408) */
	static uint16_t u[BTD_LMAX * 8];
	uint32_t w = 0x8408u;
	for (int i = 0; i < BTD_LMAX * 8; i++) { w = crc_unstep(w); u[i] = (uint16_t)w; }
	for (int j = 0; j < BTD_LMAX; j++)
		for (int h = 0; h < 2; h++)
			for (int v = 0; v < 16; v++) {
				uint16_t x = 0;
				for (int b = 0; b < 4; b++)
					if ((v >> b) & 1) x ^= u[8 * j + 4 * h + b];
				t.nib[(2 * j + h) * 16 + v] = x;
			}
	for (int q = 0; q < 127; q++) {
		uint16_t x = 0;
		t.wp[q * BTD_LMAX] = 0;
		for (int L = 1; L < BTD_LMAX; L++) {
			for (int i = 8 * (L - 1); i < 8 * L; i++)
				if (seq[(q + i) % 127]) x ^= u[i];
			t.wp[q * BTD_LMAX + L] = x;
		}
	}
	for (int c = 0; c < 64; c++) t.s.q18[c] = (uint8_t)((t.s.phase[c] + 18) % 127);
	for (int c = 0; c < 64; c++)
		t.s.wp20[c] = t.wp[((t.s.phase[c] + 18) % 127) * BTD_LMAX + 20];
}

}  // namespace

const btd_tables *btd_host_tables()
{
	std::call_once(g_once, build);
	return &g_tables;
}
