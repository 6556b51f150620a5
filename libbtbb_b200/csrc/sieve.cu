/*
 * sieve.cu -- UAP / CLK1-6 discovery from packet headers for many piconets at once.
 *
 * What the reference does per packet (btbb_process_packet in survey mode,
 * bluetooth_piconet.c:851-858, calling btbb_uap_from_header, :648-750): for each of the 64
 * possible values of CLK1-6 at the first packet that is still a candidate, dewhiten the header
 * with the clock that candidate implies for THIS packet, derive the UAP from the HEC
 * (try_clock, bluetooth_packet.c:1178-1195) and, where it agrees with what the candidate
 * implied before, test the payload CRC (crc_check, :708-769).  Candidates whose UAP changes or
 * whose CRC fails are dropped; a CRC success or a single survivor fixes UAP and CLK1-6.
 *
 * Here the 64 try_clock / crc_check evaluations of EVERY packet run first, as one pass of the
 * decode kernel (decode.cu, one warp per packet, one lane per clock) that leaves a 16-bit word
 * per (packet, clock).  The elimination itself is sequential in a piconet's packets and
 * trivially parallel across piconets: one warp per piconet walks its packets, lane L owning
 * candidates L and L + 32, with ballots for "first success in candidate order" (the reference
 * returns from inside its loop, so later candidates must stay untouched) and for the
 * survivor count.  Because the reference stops looking at a piconet once its UAP is known, the
 * work is done in rounds of 4, 32, 256, ... packets per piconet, and the 64-clock table is only
 * computed for the packets of piconets that are still unresolved.
 */
#include <string.h>
#include "bt_math.h"
#include "capi_internal.h"

namespace {

constexpr uint32_t F_UAP_VALID = 1u << 2, F_CLK6_VALID = 1u << 4, F_CLK27_VALID = 1u << 5, F_HOP_INIT = 1u << 9,
		   F_GOT_FIRST = 1u << 10, F_IS_AFH = 1u << 11, F_LOOKS_AFH = 1u << 12;
constexpr int MAX_PATTERN_LENGTH = 1000;      /* bluetooth_piconet.h:27 */

/* reset() (bluetooth_piconet.c:547-568) */
__device__ __forceinline__ void sieve_reset(uint32_t &flags, int &pobs)
{
	flags &= ~(F_GOT_FIRST | F_HOP_INIT | F_UAP_VALID | F_CLK6_VALID | F_CLK27_VALID | F_IS_AFH);
	if (flags & F_LOOKS_AFH) flags |= F_IS_AFH;
	pobs = 0;
}

/* cur[g] = first packet of piconet g that has not been handled yet */
__global__ void sieve_init_kernel(const int64_t *group_start, int64_t n_groups, int64_t *cur)
{
	const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (g < n_groups) cur[g] = group_start[g];
}

/* The packets the next round needs the 64-clock table for: the next `window` packets of every
 * piconet whose UAP is still unknown.  One warp per piconet; *counter is the list length. */
__global__ void __launch_bounds__(128) sieve_list_kernel(const int64_t *group_start, int64_t n_groups,
							 const btbb_b200_sieve *states, const int64_t *cur, int64_t window,
							 int64_t *idx, unsigned long long *counter)
{
	const int lane = threadIdx.x & 31;
	const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	if (g >= n_groups) return;
	const int64_t c = cur[g], end = group_start[g + 1];
	if (c >= end || (states[g].flags & F_UAP_VALID)) return;
	const int64_t cnt = end - c < window ? end - c : window;
	unsigned long long base = 0;
	if (lane == 0) base = atomicAdd(counter, (unsigned long long)cnt);
	base = __shfl_sync(0xffffffffu, base, 0);
	for (int64_t i = lane; i < cnt; i += 32) idx[base + i] = c + i;
}

/* One round: every piconet with packets left takes up to `window` more of them through
 * btbb_process_packet's survey branch; once its UAP is known (now or on entry) the rest of its
 * packets only mark their channels (btbb_piconet_set_channel_seen) and report "not called". */
__global__ void __launch_bounds__(128) sieve_kernel(const btbb_b200_pkt_in *pkts, const uint8_t *present,
						    const uint16_t *tc, const int64_t *group_start, int64_t n_groups,
						    btbb_b200_sieve *states, int8_t *rv_out, int64_t *cur, int64_t window)
{
	const int lane = threadIdx.x & 31;
	const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	if (g >= n_groups) return;
	const int64_t end = group_start[g + 1];
	int64_t p = cur[g];
	if (p >= end) return;
	const int64_t limit = end - p < window ? end : p + window;
	btbb_b200_sieve *st = &states[g];
	uint32_t flags = st->flags, first = st->first_pkt_time;
	int clk_offset = st->clk_offset, pobs = st->packets_observed, total = st->total_packets_observed;
	uint32_t uap = st->uap, used = st->used_channels;
	uint32_t afh = lane < 10 ? st->afh_map[lane] : 0;
	int c0 = st->clock6_candidates[lane], c1 = st->clock6_candidates[lane + 32];

	while (p < end) {
		if (flags & F_UAP_VALID) {
			/* the rest of the piconet's packets, 32 at a time: channels seen, nothing called */
			uint32_t m0 = 0, m1 = 0, m2 = 0;
			for (int64_t q = p + lane; q < end; q += 32) {
				const uint32_t channel = pkts[q].reserved & 0xffu;
				if (channel < 32) m0 |= 1u << channel;
				else if (channel < 64) m1 |= 1u << (channel - 32);
				else if (channel < 80) m2 |= 1u << (channel - 64);
				if (rv_out) rv_out[q] = (int8_t)BTBB_B200_SIEVE_NOT_CALLED;
			}
			m0 = __reduce_or_sync(0xffffffffu, m0); m1 = __reduce_or_sync(0xffffffffu, m1); m2 = __reduce_or_sync(0xffffffffu, m2);
			const uint32_t mine = lane < 4 ? (m0 >> (8 * lane)) & 0xff : lane < 8 ? (m1 >> (8 * (lane - 4))) & 0xff
					     : lane < 10 ? (m2 >> (8 * (lane - 8))) & 0xff : 0;
			used += __reduce_add_sync(0xffffffffu, (uint32_t)__popc(mine & ~afh));
			afh |= mine;
			p = end;
			break;
		}
		if (p >= limit) break;
		const uint32_t clkn = pkts[p].clkn, channel = pkts[p].reserved & 0xffu;
		/* btbb_piconet_set_channel_seen (:851-855 and :661) */
		if (channel < 80) {
			const uint32_t old = __shfl_sync(0xffffffffu, afh, channel >> 3), bit = 1u << (channel & 7);
			if (!(old & bit)) {
				if (lane == (int)(channel >> 3)) afh |= bit;
				used++;
			}
		}
		int rv = BTBB_B200_SIEVE_NOT_CALLED;
		if (present[p]) {
			const bool got_first = (flags & F_GOT_FIRST) != 0;
			if (!got_first) first = clkn;
			if (pobs >= MAX_PATTERN_LENGTH) {          /* "More hops than we can remember" (:665-671) */
				sieve_reset(flags, pobs);
				rv = 0;
			} else {
				pobs++; total++;
				/* candidate `count` implies clock (count + clkn - first) % 64 for this packet (:681) */
				const uint32_t e0 = tc[p * 64 + ((lane + clkn - first) & 63u)];
				const uint32_t e1 = tc[p * 64 + ((lane + 32 + clkn - first) & 63u)];
				const int u0 = e0 & 0xff, u1 = e1 & 0xff;
				const bool a0 = c0 > -1 || !got_first, a1 = c1 > -1 || !got_first;
				int k0 = -1, k1 = -1;             /* crc_chk: -1 mismatch, else the class of crc_check's value */
				if (!got_first || u0 == c0) k0 = (int)(e0 >> 8);
				if (!got_first || u1 == c1) k1 = (int)(e1 >> 8);
				/* (the reference's "UAP known but different" test, :693-695, cannot fire here:
				 * survey mode stops calling once the UAP is valid) */
				const uint32_t s0 = __ballot_sync(0xffffffffu, a0 && k0 >= 3), s1 = __ballot_sync(0xffffffffu, a1 && k1 >= 3);
				const int winner = s0 ? __ffs(s0) - 1 : s1 ? 32 + __ffs(s1) - 1 : 64;
				/* every candidate below the winner is updated; the winner and all above are left alone */
				if (a0 && lane < winner) c0 = (k0 == 1 || k0 == 2) ? u0 : -1;
				if (a1 && lane + 32 < winner) c1 = (k1 == 1 || k1 == 2) ? u1 : -1;
				if (winner < 64) {                 /* CRC success (:719-733) */
					const int wu = winner < 32 ? __shfl_sync(0xffffffffu, u0, winner) : __shfl_sync(0xffffffffu, u1, winner - 32);
					clk_offset = (winner - (int)(first & 0x3f)) & 0x3f;
					uap = (uint32_t)wu;
					flags |= F_CLK6_VALID | F_UAP_VALID;
					total = 0;
					rv = 1;
				} else {
					flags |= F_GOT_FIRST;
					const uint32_t keep0 = __ballot_sync(0xffffffffu, a0 && (k0 == 1 || k0 == 2));
					const uint32_t keep1 = __ballot_sync(0xffffffffu, a1 && (k1 == 1 || k1 == 2));
					const int remaining = __popc(keep0) + __popc(keep1);
					if (remaining == 1) {      /* single survivor (:741-753) */
						const int fc = keep1 ? 32 + (31 - __clz(keep1)) : 31 - __clz(keep0);
						const int su = fc < 32 ? __shfl_sync(0xffffffffu, c0, fc) : __shfl_sync(0xffffffffu, c1, fc - 32);
						clk_offset = (fc - (int)(first & 0x3f)) & 0x3f;
						uap = (uint32_t)su & 0xff;
						flags |= F_CLK6_VALID | F_UAP_VALID;
						total = 0;
						rv = 1;
					} else {
						if (remaining == 0) sieve_reset(flags, pobs);
						rv = 0;
					}
				}
			}
		}
		if (rv_out && lane == 0) rv_out[p] = (int8_t)rv;
		p++;
	}
	st->clock6_candidates[lane] = (int16_t)c0;
	st->clock6_candidates[lane + 32] = (int16_t)c1;
	if (lane < 10) st->afh_map[lane] = (uint8_t)afh;
	if (lane == 0) {
		st->flags = flags; st->first_pkt_time = first; st->clk_offset = clk_offset;
		st->packets_observed = pobs; st->total_packets_observed = total;
		st->uap = (uint8_t)uap; st->used_channels = (uint8_t)used;
		cur[g] = p;
	}
}

}  // namespace

extern "C" int btbb_b200_uap_sieve_dev(btbb_b200_ctx *ctx, const uint8_t *d_stream, int64_t stream_length,
				       const btbb_b200_pkt_in *d_pkts, int64_t n_pkts,
				       const int64_t *d_group_start, int64_t n_groups,
				       btbb_b200_sieve *d_states, int8_t *d_rv, void *cuda_stream)
{
	if (!ctx || n_pkts < 0 || n_groups < 0 || stream_length < 0 || (n_pkts > 0 && (!d_stream || !d_pkts)) ||
	    (n_groups > 0 && (!d_group_start || !d_states)))
		return btbb_b200_set_error(BTBB_B200_EINVAL, "uap_sieve: bad arguments");
	BT_CUDA_TRY(cudaSetDevice(ctx->device));
	if (n_groups == 0) return BTBB_B200_OK;
	cudaStream_t st = (cudaStream_t)cuda_stream;
	if (n_pkts > ctx->sieve_cap) {
		if (ctx->d_sieve_tc) cudaFree(ctx->d_sieve_tc);
		if (ctx->d_sieve_present) cudaFree(ctx->d_sieve_present);
		if (ctx->d_sieve_idx) cudaFree(ctx->d_sieve_idx);
		ctx->d_sieve_tc = NULL; ctx->d_sieve_present = NULL; ctx->d_sieve_idx = NULL; ctx->sieve_cap = 0;
		const int64_t cap = n_pkts < 4096 ? 4096 : n_pkts;
		BT_CUDA_TRY(cudaMalloc(&ctx->d_sieve_tc, (size_t)cap * 64 * sizeof(uint16_t)));
		BT_CUDA_TRY(cudaMalloc(&ctx->d_sieve_present, (size_t)cap));
		BT_CUDA_TRY(cudaMalloc(&ctx->d_sieve_idx, (size_t)cap * sizeof(int64_t)));
		ctx->sieve_cap = cap;
	}
	if (n_groups > ctx->sieve_groups_cap) {
		if (ctx->d_sieve_cur) cudaFree(ctx->d_sieve_cur);
		ctx->d_sieve_cur = NULL; ctx->sieve_groups_cap = 0;
		const int64_t cap = n_groups < 1024 ? 1024 : n_groups;
		BT_CUDA_TRY(cudaMalloc(&ctx->d_sieve_cur, (size_t)(cap + 1) * sizeof(int64_t)));
		ctx->sieve_groups_cap = cap;
	}
	int64_t *cur = ctx->d_sieve_cur;
	unsigned long long *counter = reinterpret_cast<unsigned long long *>(ctx->d_sieve_cur + ctx->sieve_groups_cap);
	int rc = btbb_b200_header_present_dev(ctx, d_stream, stream_length, d_pkts, n_pkts, ctx->d_sieve_present, cuda_stream);
	if (rc) return rc;
	const int wpb = 4;
	const unsigned gblocks = (unsigned)((n_groups + wpb - 1) / wpb);
	sieve_init_kernel<<<(unsigned)((n_groups + 255) / 256), 256, 0, st>>>(d_group_start, n_groups, cur);
	/* Rounds of 4, 32, 256, ... packets per piconet: the reference stops working on a piconet as
	 * soon as its UAP is known (typically after one or two packets with a CRC), so the 64-clock
	 * table is only computed for the packets a round can still need.  No host round trip: the
	 * list length stays on the device, and a round with nothing left is three empty launches. */
	int64_t done = 0;
	for (int64_t window = 4; done < n_pkts || window == 4; window *= 8) {
		BT_CUDA_TRY(cudaMemsetAsync(counter, 0, sizeof(*counter), st));
		sieve_list_kernel<<<gblocks, wpb * 32, 0, st>>>(d_group_start, n_groups, d_states, cur, window, ctx->d_sieve_idx, counter);
		int64_t n_max = n_pkts - done;
		if (n_groups < ((int64_t)1 << 40) / window && n_groups * window < n_max) n_max = n_groups * window;
		rc = bt_try_clocks_compact(ctx, d_stream, stream_length, d_pkts, ctx->d_sieve_idx, counter, n_max, ctx->d_sieve_tc, st);
		if (rc) return rc;
		sieve_kernel<<<gblocks, wpb * 32, 0, st>>>(d_pkts, ctx->d_sieve_present, ctx->d_sieve_tc, d_group_start, n_groups,
							  d_states, d_rv, cur, window);
		done += window;        /* every piconet with packets left has now handled at least this many more */
		if (window > ((int64_t)1 << 40)) break;
	}
	BT_CUDA_TRY(cudaGetLastError());
	BT_CUDA_TRY(cudaStreamSynchronize(st));
	return BTBB_B200_OK;
}

extern "C" int btbb_b200_uap_sieve_host(btbb_b200_ctx *ctx, const char *stream, int64_t stream_length,
					const btbb_b200_pkt_in *pkts, int64_t n_pkts,
					const int64_t *group_start, int64_t n_groups,
					btbb_b200_sieve *states, int8_t *rv)
{
	if (!ctx || n_pkts < 0 || n_groups < 0 || stream_length < 0 || (n_pkts > 0 && (!stream || !pkts)) ||
	    (n_groups > 0 && (!group_start || !states)))
		return btbb_b200_set_error(BTBB_B200_EINVAL, "uap_sieve_host: bad arguments");
	if (n_groups == 0) return BTBB_B200_OK;
	BT_CUDA_TRY(cudaSetDevice(ctx->device));
	uint8_t *d_s = NULL; btbb_b200_pkt_in *d_p = NULL; int64_t *d_g = NULL; btbb_b200_sieve *d_st = NULL; int8_t *d_rv = NULL;
	int rc = BTBB_B200_OK;
	cudaError_t e;
	if ((e = cudaMalloc(&d_s, (size_t)stream_length + 1)) != cudaSuccess ||
	    (e = cudaMalloc(&d_p, (size_t)(n_pkts + 1) * sizeof(*d_p))) != cudaSuccess ||
	    (e = cudaMalloc(&d_g, (size_t)(n_groups + 1) * sizeof(*d_g))) != cudaSuccess ||
	    (e = cudaMalloc(&d_st, (size_t)n_groups * sizeof(*d_st))) != cudaSuccess ||
	    (e = cudaMalloc(&d_rv, (size_t)n_pkts + 1)) != cudaSuccess)
		rc = btbb_b200_cuda_fail(e, "cudaMalloc(uap_sieve_host)");
	if (!rc && ((e = cudaMemcpy(d_s, stream, (size_t)stream_length, cudaMemcpyHostToDevice)) != cudaSuccess ||
		    (e = cudaMemcpy(d_p, pkts, (size_t)n_pkts * sizeof(*d_p), cudaMemcpyHostToDevice)) != cudaSuccess ||
		    (e = cudaMemcpy(d_g, group_start, (size_t)(n_groups + 1) * sizeof(*d_g), cudaMemcpyHostToDevice)) != cudaSuccess ||
		    (e = cudaMemcpy(d_st, states, (size_t)n_groups * sizeof(*d_st), cudaMemcpyHostToDevice)) != cudaSuccess))
		rc = btbb_b200_cuda_fail(e, "cudaMemcpy(uap_sieve_host H2D)");
	if (!rc) rc = btbb_b200_uap_sieve_dev(ctx, d_s, stream_length, d_p, n_pkts, d_g, n_groups, d_st, d_rv, NULL);
	if (!rc && ((e = cudaMemcpy(states, d_st, (size_t)n_groups * sizeof(*d_st), cudaMemcpyDeviceToHost)) != cudaSuccess ||
		    (rv && n_pkts > 0 && (e = cudaMemcpy(rv, d_rv, (size_t)n_pkts, cudaMemcpyDeviceToHost)) != cudaSuccess)))
		rc = btbb_b200_cuda_fail(e, "cudaMemcpy(uap_sieve_host D2H)");
	if (d_s) cudaFree(d_s);
	if (d_p) cudaFree(d_p);
	if (d_g) cudaFree(d_g);
	if (d_st) cudaFree(d_st);
	if (d_rv) cudaFree(d_rv);
	return rc;
}
