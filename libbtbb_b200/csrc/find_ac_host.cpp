/*
 * find_ac_host.cpp -- btbb_find_ac (bluetooth_packet.c:444-464) for SHORT searches, on the host.
 *
 * The classic callers hand btbb_find_ac a few hundred to a few thousand symbols at a time (one
 * Ubertooth USB transfer); below a few thousand positions a host-to-device copy, a kernel launch
 * and the read-back of the result cost more than the whole search ("Drop-in
 * latency").  compat.cu therefore answers such calls here: one sliding 64-bit window, the Barker
 * tail test, the (64,30) syndrome from bt_math.h and a probe of the host copy of the same
 * open-addressing error table the kernels use (tables.cu).  First hit only.  The batch entry
 * points (btbb_b200_find_ac_dev / _host / _packed_dev / _sharded_*) never come here.
 */
#include <string.h>
#include "bt_math.h"
#include "scan_hash.h"
#include "capi_internal.h"

static inline int popc64(uint64_t v) { return __builtin_popcountll(v); }

int bt_find_first_cpu(const btbb_b200_ctx *ctx, const char *stream, int search_length, uint32_t lap,
		      int max_ac_errors, btbb_b200_hit *hit, int *found)
{
	*found = 0;
	if (search_length <= 0) return BTBB_B200_OK;
	const bool known = lap != BTBB_B200_LAP_ANY;
	const uint64_t ac = known ? bt_gen_syncword(lap) : 0;
	uint64_t w = 0;
	for (int i = 0; i < 64; i++) w |= (uint64_t)(stream[i] & 1) << i;      /* air_to_host64 (:235-242) */
	for (int p = 0; p < search_length; p++) {
		if (known) {                                   /* find_known_lap (:430-438) */
			const int d = popc64(w ^ ac);
			if (d <= max_ac_errors) {
				hit->offset = p; hit->lap = lap; hit->ac_errors = (uint8_t)d;
				hit->pad[0] = hit->pad[1] = hit->pad[2] = 0;
				*found = 1;
				return BTBB_B200_OK;
			}
		} else {                                       /* promiscuous_packet_search (:385-416) */
			const uint32_t tail = (uint32_t)(w >> 57);
			const int da = __builtin_popcount(tail ^ BT_BARKER_A);
			if (da <= 1 || da >= 6) {                  /* BARKER_DISTANCE <= 1 (:55-59, :385) */
				uint64_t sw = (w & 0x01ffffffffffffffULL) | ((uint64_t)(da <= 1 ? BT_BARKER_A : BT_BARKER_B) << 57);
				const uint64_t syn = bt_syndrome_slow(sw ^ BT_PN) & 0x3ffffffffULL;
				int e = 0;
				if (syn) {
					e = 0xff;
					if (ctx->h_err) {
						const uint64_t mask = ((uint64_t)1 << ctx->err_log2) - 1;
						uint64_t h = bt_err_hash(syn, ctx->err_log2);
						for (;;) {
							const bt_err_slot &sl = ctx->h_err[h];
							if (sl.syn == syn) { sw ^= sl.err; e = popc64(sl.err); break; }   /* Barker fixes are not counted */
							if (sl.syn == 0) break;
							h = (h + 1) & mask;
						}
					}
				}
				if (e <= max_ac_errors) {
					hit->offset = p; hit->lap = (uint32_t)(sw >> 34) & 0xffffffu; hit->ac_errors = (uint8_t)e;
					hit->pad[0] = hit->pad[1] = hit->pad[2] = 0;
					*found = 1;
					return BTBB_B200_OK;
				}
			}
		}
		if (p + 1 < search_length) w = (w >> 1) | ((uint64_t)(stream[p + 64] & 1) << 63);
	}
	return BTBB_B200_OK;
}
