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
#include <stdlib.h>
#include <string.h>
#include "bt_math.h"
#include "scan_hash.h"
#include "capi_internal.h"

#include <mutex>

namespace {

inline int popc64(uint64_t v) { return __builtin_popcountll(v); }

/* gen_syndrome (bluetooth_packet.c:147-159) by bytes of the information part: the remainder is linear,
 * so cw mod g(x) = low 34 bits XOR one table entry per byte of cw >> 34.  Also the Barker verdict of
 * every 7-bit tail: 0 = further than 1 from both tails, else the tail to put in its place. */
struct host_tables {
	uint64_t rem[4][256];
	uint8_t tail[128];
};
const host_tables &tables()
{
	static host_tables T;
	static std::once_flag once;
	std::call_once(once, [] {
		for (int j = 0; j < 4; j++)
			for (int v = 0; v < 256; v++)
				T.rem[j][v] = bt_syndrome_slow((uint64_t)v << (34 + 8 * j)) & 0x3ffffffffULL;
		for (int t = 0; t < 128; t++) {
			const int da = __builtin_popcount((unsigned)t ^ BT_BARKER_A);      /* BARKER_DISTANCE <= 1 (:55-59, :385) */
			T.tail[t] = da <= 1 ? (uint8_t)BT_BARKER_A : da >= 6 ? (uint8_t)BT_BARKER_B : 0;
		}
		static_assert(BT_BARKER_A != 0 && BT_BARKER_B != 0, "0 marks a rejected tail");
	});
	return T;
}

const int CHUNK = 4096;      /* positions per packed block */

}  // namespace

static int find_first_core(const bt_err_slot *h_err, int err_log2, const char *stream, int search_length, uint32_t lap,
			   int max_ac_errors, btbb_b200_hit *hit, int *found)
{
	*found = 0;
	if (search_length <= 0) return BTBB_B200_OK;
	const host_tables &T = tables();
	const bool known = lap != BTBB_B200_LAP_ANY;
	const uint64_t ac = known ? bt_gen_syncword(lap) : 0;
	uint64_t bits[CHUNK / 64 + 2];
	for (int first = 0; first < search_length; first += CHUNK) {
		/* air_to_host64 (:235-242) of every window of the block: symbol first + i -> bit i */
		const int npos = search_length - first < CHUNK ? search_length - first : CHUNK, nsym = npos + 63;
		const char *s = stream + first;
		memset(bits, 0, sizeof(bits));
		int i = 0;
		for (; i + 8 <= nsym; i += 8) {
			uint64_t x;
			memcpy(&x, s + i, 8);
			bits[i >> 6] |= (((x & 0x0101010101010101ull) * 0x0102040810204080ull) >> 56) << (i & 63);
		}
		for (; i < nsym; i++) bits[i >> 6] |= (uint64_t)(s[i] & 1) << (i & 63);
		for (int p = 0; p < npos; p++) {
			const int sh = p & 63;
			const uint64_t w = sh ? (bits[p >> 6] >> sh) | (bits[(p >> 6) + 1] << (64 - sh)) : bits[p >> 6];
			if (known) {                                   /* find_known_lap (:430-438) */
				const int d = popc64(w ^ ac);
				if (d > max_ac_errors) continue;
				hit->offset = first + p; hit->lap = lap; hit->ac_errors = (uint8_t)d;
			} else {                                       /* promiscuous_packet_search (:385-416) */
				const uint64_t tail = T.tail[w >> 57];
				if (!tail) continue;
				uint64_t sw = (w & 0x01ffffffffffffffULL) | (tail << 57);
				const uint64_t c = sw ^ BT_PN, info = c >> 34;
				const uint64_t syn = (c & 0x3ffffffffULL) ^ T.rem[0][info & 255] ^ T.rem[1][(info >> 8) & 255] ^
						     T.rem[2][(info >> 16) & 255] ^ T.rem[3][info >> 24];
				int e = 0;
				if (syn) {
					e = 0xff;
					if (h_err) {
						const uint64_t mask = ((uint64_t)1 << err_log2) - 1;
						uint64_t h = bt_err_hash(syn, err_log2);
						for (;;) {
							const bt_err_slot &sl = h_err[h];
							if (sl.syn == syn) { sw ^= sl.err; e = popc64(sl.err); break; }   /* Barker fixes are not counted */
							if (sl.syn == 0) break;
							h = (h + 1) & mask;
						}
					}
				}
				if (e > max_ac_errors) continue;
				hit->offset = first + p; hit->lap = (uint32_t)(sw >> 34) & 0xffffffu; hit->ac_errors = (uint8_t)e;
			}
			hit->pad[0] = hit->pad[1] = hit->pad[2] = 0;
			*found = 1;
			return BTBB_B200_OK;
		}
	}
	return BTBB_B200_OK;
}

int bt_find_first_cpu(const btbb_b200_ctx *ctx, const char *stream, int search_length, uint32_t lap,
		      int max_ac_errors, btbb_b200_hit *hit, int *found)
{
	return find_first_core(ctx->h_err, ctx->err_log2, stream, search_length, lap, max_ac_errors, hit, found);
}

/* The short-search host path without a context: the first hit of a btbb_find_ac call after
 * btbb_init(table_errors), computed by the same routine compat.cu uses for searches of at most 8192
 * positions.  The error table (what tables.cu builds and keeps a host copy of: syndrome -> pattern for
 * up to table_errors errors in bits 0..57, bluetooth_packet.c:161-185, open addressing with
 * bt_err_hash) is built here on first use.  For callers that only ever search short buffers and for
 * the CPU-only tests; table_errors 0..4. */
extern "C" int btbb_b200_find_first_smallcall(const char *stream, int search_length, uint32_t lap, int table_errors,
					      int max_ac_errors, btbb_b200_hit *hit, int *found)
{
	if (!stream || !hit || !found || search_length < 0 || table_errors < 0 || table_errors > 4 || max_ac_errors < 0)
		return btbb_b200_set_error(BTBB_B200_EINVAL, "find_first_smallcall: bad arguments");
	static std::mutex lock;
	static bt_err_slot *tabs[5];
	static int logs[5];
	{
		std::lock_guard<std::mutex> g(lock);
		if (table_errors > 0 && !tabs[table_errors]) {
			uint64_t col[58];
			for (int i = 0; i < 58; i++) col[i] = bt_syndrome_slow(1ULL << i);
			size_t n = 0, c = 1;
			for (int w = 1; w <= table_errors; w++) { c = c * (size_t)(58 - w + 1) / (size_t)w; n += c; }
			int lg = 4;
			while (((size_t)1 << lg) < 2 * n) lg++;
			bt_err_slot *tab = (bt_err_slot *)calloc((size_t)1 << lg, sizeof(bt_err_slot));
			if (!tab) return btbb_b200_set_error(BTBB_B200_ENOMEM, "find_first_smallcall: out of host memory");
			const uint64_t mask = ((uint64_t)1 << lg) - 1;
			/* every pattern of 1..table_errors bits among bits 0..57, by weight and then in index order */
			int idx[4];
			for (int w = 1; w <= table_errors; w++) {
				for (int j = 0; j < w; j++) idx[j] = j;
				for (;;) {
					uint64_t e = 0, sy = 0;
					for (int j = 0; j < w; j++) { e |= 1ULL << idx[j]; sy ^= col[idx[j]]; }
					uint64_t h = bt_err_hash(sy, lg);
					while (tab[h].syn) h = (h + 1) & mask;
					tab[h].syn = sy; tab[h].err = e;
					int j = w - 1;
					while (j >= 0 && idx[j] == 58 - w + j) j--;
					if (j < 0) break;
					idx[j]++;
					for (int t = j + 1; t < w; t++) idx[t] = idx[t - 1] + 1;
				}
			}
			tabs[table_errors] = tab; logs[table_errors] = lg;
		}
	}
	return find_first_core(tabs[table_errors], logs[table_errors], stream, search_length, lap, max_ac_errors, hit, found);
}
