/*
 * hops.cu -- K9: hop-sequence generation and CLK1-27 winnowing.
 *
 * What the reference does (bluetooth_piconet.c): gen_hops (:311-362) fills a 2^27-entry table
 * with the channel of every CLK1-27 value of one piconet address, five nested loops around a
 * 512 KiB permutation table (perm5 :258-290 through perm_table); btbb_init_hop_reversal (:475-499)
 * scans that table for the entries that agree with the first observed hop (init_candidates
 * :455-472, one in 64 x 79 of 2^27), and channel_winnow / btbb_winnow (:575-645) filter the
 * candidate list once per further observed (time, channel) pair.
 *
 * Here:
 *   hop_sequence_kernel   the table itself, any index range, one thread per 16 consecutive entries
 *       (one 128-bit store).  The butterfly network factors by its control bits: the first five
 *       stages are steered by c (5 bits), the last nine by d (9 bits), so perm5(z, c, d) =
 *       D[d][C[c][z]] with a 1 KiB and a 16 KiB table (the reference's perm_table is 512 KiB).  Both
 *       are address-independent, built once per process and staged into shared memory per CTA; an
 *       entry then costs two table reads, the register-bank read and ~10 integer operations.
 *   hop_winnow_*          hop reversal without the table: all 2^21 CLK1-27 values that agree with the
 *       known CLK1-6 are candidates; each thread walks the observations for one candidate until the
 *       first disagreement (78 of 79 fall at the first), evaluating the hop selection kernel
 *       directly.  The per-observation survivor counts reproduce the reference's num_candidates
 *       after every channel_winnow call; the survivors are compacted in ascending order (the order
 *       of the reference's candidate list).
 */
#include <cuda_runtime.h>
#include <string.h>
#include "capi_internal.h"

namespace {

constexpr uint32_t SEQ_MASK = (1u << 27) - 1u;
constexpr int WIN_BLOCK = 256;
constexpr int N_CAND = 1 << 21;                      /* CLK1-27 values with given CLK1-6 */
constexpr int WIN_BLOCKS = N_CAND / WIN_BLOCK;

struct hop_consts {
	int a1, b, c1, d1, e;            /* address_precalc (:197-215) */
	int used;                        /* channels in the bank: 79, or the AFH map's population */
	uint32_t inv;                    /* ceil(2^16 / used): v % used for v < 512 as v - used * ((v * inv) >> 16) */
	int afh, aliased;
	uint8_t bank[80];                /* precalc (:171-193) */
};

__host__ __device__ __forceinline__ uint32_t mod_small(uint32_t v, uint32_t m, uint32_t inv)
{
	return v - m * ((v * inv) >> 16);
}

/* stage i of perm5 swaps bits i1[i], i2[i] when control bit i is set; stages run 13 .. 0,
 * control = c (5 bits) << 9 | d (9 bits) */

/* perm5 (:258-290) for one input, stages [hi, lo] of the butterfly network */
__host__ __device__ __forceinline__ uint32_t perm5_stages(uint32_t z, uint32_t ctrl, int hi, int lo)
{
	const int i1[14] = {0, 2, 1, 3, 0, 1, 0, 3, 1, 0, 2, 1, 0, 1};
	const int i2[14] = {1, 3, 2, 4, 4, 3, 2, 4, 4, 3, 4, 3, 3, 2};
	for (int i = hi; i >= lo; i--) {
		const uint32_t t = ((z >> i1[i]) ^ (z >> i2[i])) & (ctrl >> i) & 1u;
		z ^= (t << i1[i]) | (t << i2[i]);
	}
	return z;
}

/* C[c][z]: stages 13..9 under control c; D[d][z]: stages 8..0 under control d */
struct perm_tables { uint8_t c[32 * 32]; uint8_t d[512 * 32]; };
constexpr int PERM_BYTES = (int)sizeof(perm_tables);
static_assert(PERM_BYTES % 16 == 0, "perm tables are copied as 128-bit words");

/* sequence entries [16 q, 16 q + 16): x = 8 (q & 3) .. + 7, both values of clock bit 1 */
__device__ __forceinline__ void hop16(const hop_consts &h, const uint8_t *s_bank, const perm_tables *pt, uint32_t q, uint8_t out[16])
{
	const uint32_t kk = q >> 2;                       /* index >> 6: the (h, i, j, k) tuple */
	const uint32_t a = (uint32_t)h.a1 ^ ((q >> 16) & 31u);
	const uint32_t c = (uint32_t)h.c1 ^ ((q >> 11) & 31u);
	const uint32_t d = (uint32_t)h.d1 ^ (kk & 511u);
	const uint32_t f = (16u * (kk & 0x1fffffu)) % 79u;
	const uint8_t *c0 = pt->c + 32 * c, *c1 = pt->c + 32 * (c ^ 31u), *dd = pt->d + 32 * d;
	const uint32_t base = (uint32_t)h.e + (h.afh ? mod_small(f, (uint32_t)h.used, h.inv) : f);
	#pragma unroll
	for (int t = 0; t < 8; t++) {
		const uint32_t x = 8u * (q & 3u) + (uint32_t)t;
		const uint32_t z = ((x + a) & 31u) ^ (uint32_t)h.b;
		out[2 * t] = s_bank[mod_small(dd[c0[z]] + base, (uint32_t)h.used, h.inv)];
		out[2 * t + 1] = s_bank[mod_small(dd[c1[z]] + base + 32u, (uint32_t)h.used, h.inv)];
	}
}

__global__ void __launch_bounds__(256) hop_sequence_kernel(hop_consts h, const perm_tables *g_pt, int64_t first, int64_t n, uint8_t *out)
{
	__shared__ uint8_t s_bank[80];
	__shared__ __align__(16) perm_tables s_pt;
	if (threadIdx.x < 80) s_bank[threadIdx.x] = h.bank[threadIdx.x];
	for (int i = threadIdx.x; i < PERM_BYTES / 16; i += blockDim.x)
		reinterpret_cast<uint4 *>(&s_pt)[i] = reinterpret_cast<const uint4 *>(g_pt)[i];
	__syncthreads();
	const int64_t q0 = first >> 4, q1 = (first + n + 15) >> 4;
	const bool aligned = ((reinterpret_cast<uintptr_t>(out) - (uintptr_t)first) & 15) == 0;
	for (int64_t q = q0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < q1; q += (int64_t)gridDim.x * blockDim.x) {
		union { uint8_t b[16]; uint4 v; } u;
		hop16(h, s_bank, &s_pt, (uint32_t)q & (SEQ_MASK >> 4), u.b);
		const int64_t i0 = 16 * q;
		if (aligned && i0 >= first && i0 + 16 <= first + n)
			*reinterpret_cast<uint4 *>(out + (i0 - first)) = u.v;
		else
			for (int t = 0; t < 16; t++)
				if (i0 + t >= first && i0 + t < first + n) out[i0 + t - first] = u.b[t];
	}
}

/* one hop (single entry of the sequence): what the table lookup sequence[i] returns */
__device__ __forceinline__ uint32_t hop_one(const hop_consts &h, const uint8_t *s_bank, uint32_t idx)
{
	const uint32_t y1 = idx & 1u, pair = idx >> 1, kk = pair >> 5;
	const uint32_t x = pair & 31u;
	const uint32_t a = (uint32_t)h.a1 ^ ((pair >> 19) & 31u);
	const uint32_t c = ((uint32_t)h.c1 ^ ((pair >> 14) & 31u)) ^ (y1 ? 31u : 0u);
	const uint32_t d = (uint32_t)h.d1 ^ (kk & 511u);
	const uint32_t ctrl = (c << 9) | d;
	const uint32_t z = perm5_stages(((x + a) & 31u) ^ (uint32_t)h.b, ctrl, 13, 0);
	const uint32_t f = (16u * kk) % 79u;
	const uint32_t v = z + (uint32_t)h.e + (h.afh ? mod_small(f, (uint32_t)h.used, h.inv) : f) + 32u * y1;
	uint32_t ch = s_bank[mod_small(v, (uint32_t)h.used, h.inv)];
	if (h.aliased) ch = ((ch + 24u) % 25u) + 26u;         /* aliased_channel (:449-452) */
	return ch;
}

/* candidate t = CLK1-27 value known6 + 64 t: the index of the first observation it disagrees with
 * (n_obs if none); per-observation elimination counts; survivors per block */
__global__ void __launch_bounds__(WIN_BLOCK) hop_winnow_kernel(hop_consts h, uint32_t known6, int n_obs, const int32_t *idx,
							       const uint8_t *chan, uint16_t *first_fail, unsigned int *fail_hist,
							       unsigned int *block_count)
{
	__shared__ uint8_t s_bank[80];
	__shared__ unsigned int s_cnt, s_hist[64];      /* nearly every candidate falls at one of the first observations */
	if (threadIdx.x < 80) s_bank[threadIdx.x] = h.bank[threadIdx.x];
	if (threadIdx.x == 0) s_cnt = 0;
	if (threadIdx.x < 64) s_hist[threadIdx.x] = 0;
	__syncthreads();
	const uint32_t t = blockIdx.x * WIN_BLOCK + threadIdx.x;
	const uint32_t cand = known6 + 64u * t;
	int m = 0;
	for (; m < n_obs; m++)
		if (hop_one(h, s_bank, (cand + (uint32_t)idx[m]) & SEQ_MASK) != chan[m]) break;
	first_fail[t] = (uint16_t)m;
	if (m >= n_obs) atomicAdd(&s_cnt, 1u);
	else if (m < 64) atomicAdd(&s_hist[m], 1u);
	else atomicAdd(&fail_hist[m], 1u);
	__syncthreads();
	if (threadIdx.x == 0) block_count[blockIdx.x] = s_cnt;
	if (threadIdx.x < 64 && threadIdx.x < n_obs && s_hist[threadIdx.x]) atomicAdd(&fail_hist[threadIdx.x], s_hist[threadIdx.x]);
}

/* exclusive scan of the per-block survivor counts (one block), total in *total */
__global__ void __launch_bounds__(1024) hop_scan_kernel(unsigned int *block_count, int nblocks, unsigned int *total)
{
	__shared__ unsigned int wsum[32];
	__shared__ unsigned int carry;
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	if (threadIdx.x == 0) carry = 0;
	__syncthreads();
	for (int b0 = 0; b0 < nblocks; b0 += 1024) {
		const int i = b0 + threadIdx.x;
		const unsigned int x = i < nblocks ? block_count[i] : 0;
		unsigned int inc = x;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const unsigned int u = __shfl_up_sync(0xffffffffu, inc, d);
			if (lane >= d) inc += u;
		}
		if (lane == 31) wsum[w] = inc;
		__syncthreads();
		if (w == 0) {
			const unsigned int y = wsum[lane];
			unsigned int z = y;
			#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const unsigned int u = __shfl_up_sync(0xffffffffu, z, d);
				if (lane >= d) z += u;
			}
			wsum[lane] = z - y;
		}
		__syncthreads();
		const unsigned int excl = carry + wsum[w] + inc - x;
		if (i < nblocks) block_count[i] = excl;
		__syncthreads();
		if (threadIdx.x == 1023) carry = excl + x;
		__syncthreads();
	}
	if (threadIdx.x == 0) *total = carry;
}

/* survivors, ascending: block b writes its survivors from block_count[b] (now the exclusive base) */
__global__ void __launch_bounds__(WIN_BLOCK) hop_compact_kernel(uint32_t known6, int n_obs, const uint16_t *first_fail,
								const unsigned int *block_base, uint32_t *cands, unsigned int max_out)
{
	__shared__ unsigned int wbase[WIN_BLOCK / 32];
	const uint32_t t = blockIdx.x * WIN_BLOCK + threadIdx.x;
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const bool live = first_fail[t] == (uint16_t)n_obs;
	const unsigned int m = __ballot_sync(0xffffffffu, live);
	if (lane == 0) wbase[w] = __popc(m);
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned int run = block_base[blockIdx.x];
		for (int i = 0; i < WIN_BLOCK / 32; i++) { const unsigned int c = wbase[i]; wbase[i] = run; run += c; }
	}
	__syncthreads();
	if (live) {
		const unsigned int at = wbase[w] + __popc(m & ((1u << lane) - 1u));
		if (at < max_out) cands[at] = known6 + 64u * t;
	}
}

int make_consts(const btbb_b200_hop_cfg *cfg, hop_consts *h)
{
	memset(h, 0, sizeof(*h));
	const uint32_t address = cfg->address & 0xfffffffu;
	h->a1 = (address >> 23) & 0x1f;
	h->b = (address >> 19) & 0x0f;
	h->c1 = ((address >> 4) & 0x10) + ((address >> 3) & 0x08) + ((address >> 2) & 0x04) + ((address >> 1) & 0x02) + (address & 0x01);
	h->d1 = (address >> 10) & 0x1ff;
	h->e = ((address >> 7) & 0x40) + ((address >> 6) & 0x20) + ((address >> 5) & 0x10) + ((address >> 4) & 0x08) +
	       ((address >> 3) & 0x04) + ((address >> 2) & 0x02) + ((address >> 1) & 0x01);
	h->afh = cfg->afh != 0; h->aliased = cfg->aliased != 0;
	int used = 0;
	for (int i = 0; i < 79; i++) {
		const int chan = (i * 2) % 79;
		if (!h->afh) h->bank[i] = (uint8_t)chan;
		else if (cfg->afh_map[chan / 8] & (1 << (chan % 8))) h->bank[used++] = (uint8_t)chan;
	}
	h->used = h->afh ? used : 79;
	if (h->used < 1) return btbb_b200_set_error(BTBB_B200_EINVAL, "hop: empty AFH channel map");
	h->inv = 65536u / (uint32_t)h->used + 1u;
	for (uint32_t v = 0; v < 512; v++)      /* every value the kernels reduce is below 512 */
		if (mod_small(v, (uint32_t)h->used, h->inv) != v % (uint32_t)h->used)
			return btbb_b200_set_error(BTBB_B200_EINVAL, "hop: internal reciprocal check failed");
	return BTBB_B200_OK;
}

}  // namespace

extern "C" int btbb_b200_hop_sequence_dev(btbb_b200_ctx *ctx, const btbb_b200_hop_cfg *cfg, int64_t first, int64_t n,
					  uint8_t *d_out, void *cuda_stream)
{
	if (!ctx || !cfg || first < 0 || n < 0 || first + n > ((int64_t)1 << 27) || (n > 0 && !d_out))
		return btbb_b200_set_error(BTBB_B200_EINVAL, "hop_sequence: bad arguments");
	hop_consts h;
	int rc = make_consts(cfg, &h);
	if (rc) return rc;
	if (n == 0) return BTBB_B200_OK;
	BT_CUDA_TRY(cudaSetDevice(ctx->device));
	const int64_t chunks = ((first + n + 15) >> 4) - (first >> 4);
	int64_t blocks = (chunks + 255) / 256;
	const int64_t cap = (int64_t)ctx->sm_count * 8;
	if (blocks > cap) blocks = cap;
	if (!ctx->d_perm_tables) {
		perm_tables *t = new perm_tables;
		for (uint32_t c = 0; c < 32; c++)
			for (uint32_t z = 0; z < 32; z++) t->c[32 * c + z] = (uint8_t)perm5_stages(z, c << 9, 13, 9);
		for (uint32_t d = 0; d < 512; d++)
			for (uint32_t z = 0; z < 32; z++) t->d[32 * d + z] = (uint8_t)perm5_stages(z, d, 8, 0);
		void *dp = NULL;
		cudaError_t e = cudaMalloc(&dp, sizeof(perm_tables));
		if (e == cudaSuccess) e = cudaMemcpy(dp, t, sizeof(perm_tables), cudaMemcpyHostToDevice);
		delete t;
		if (e != cudaSuccess) { if (dp) cudaFree(dp); return btbb_b200_cuda_fail(e, "hop_sequence: permutation tables"); }
		ctx->d_perm_tables = dp;
	}
	hop_sequence_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)cuda_stream>>>(h, static_cast<const perm_tables *>(ctx->d_perm_tables), first, n, d_out);
	BT_CUDA_TRY(cudaGetLastError());
	return BTBB_B200_OK;
}

extern "C" int btbb_b200_hop_winnow(btbb_b200_ctx *ctx, const btbb_b200_hop_cfg *cfg, uint32_t known_clk6, int n_obs,
				    const int32_t *indices, const uint8_t *channels, uint32_t *candidates, int64_t max_candidates,
				    int64_t *n_candidates, int32_t *survivors_after)
{
	if (!ctx || !cfg || n_obs < 1 || n_obs > 65535 || !indices || !channels || !n_candidates || max_candidates < 0 ||
	    (max_candidates > 0 && !candidates))
		return btbb_b200_set_error(BTBB_B200_EINVAL, "hop_winnow: bad arguments");
	hop_consts h;
	int rc = make_consts(cfg, &h);
	if (rc) return rc;
	BT_CUDA_TRY(cudaSetDevice(ctx->device));
	std::lock_guard<std::mutex> guard(*ctx->host_lock);
	/* scratch 3: observations | first-fail table | histogram | block counts | total | survivors */
	const size_t o_idx = 0, o_ch = o_idx + (size_t)n_obs * 4, o_ff = (o_ch + (size_t)n_obs + 15) & ~(size_t)15;
	const size_t o_hist = o_ff + (size_t)N_CAND * 2, o_blk = o_hist + (size_t)n_obs * 4, o_tot = o_blk + (size_t)WIN_BLOCKS * 4;
	const size_t o_out = (o_tot + 4 + 15) & ~(size_t)15;
	const size_t out_cap = (size_t)(max_candidates < N_CAND ? max_candidates : N_CAND);
	const size_t need = o_out + out_cap * 4 + 16;
	if (need > ctx->scratch_cap[3]) {
		if (ctx->d_scratch[3]) cudaFree(ctx->d_scratch[3]);
		ctx->d_scratch[3] = NULL; ctx->scratch_cap[3] = 0;
		BT_CUDA_TRY(cudaMalloc(&ctx->d_scratch[3], need));
		ctx->scratch_cap[3] = need;
	}
	char *d = static_cast<char *>(ctx->d_scratch[3]);
	cudaStream_t st = 0;
	BT_CUDA_TRY(cudaMemcpyAsync(d + o_idx, indices, (size_t)n_obs * 4, cudaMemcpyHostToDevice, st));
	BT_CUDA_TRY(cudaMemcpyAsync(d + o_ch, channels, (size_t)n_obs, cudaMemcpyHostToDevice, st));
	BT_CUDA_TRY(cudaMemsetAsync(d + o_hist, 0, (size_t)n_obs * 4, st));
	hop_winnow_kernel<<<WIN_BLOCKS, WIN_BLOCK, 0, st>>>(h, known_clk6 & 63u, n_obs, reinterpret_cast<const int32_t *>(d + o_idx),
							    reinterpret_cast<const uint8_t *>(d + o_ch), reinterpret_cast<uint16_t *>(d + o_ff),
							    reinterpret_cast<unsigned int *>(d + o_hist), reinterpret_cast<unsigned int *>(d + o_blk));
	hop_scan_kernel<<<1, 1024, 0, st>>>(reinterpret_cast<unsigned int *>(d + o_blk), WIN_BLOCKS, reinterpret_cast<unsigned int *>(d + o_tot));
	hop_compact_kernel<<<WIN_BLOCKS, WIN_BLOCK, 0, st>>>(known_clk6 & 63u, n_obs, reinterpret_cast<const uint16_t *>(d + o_ff),
							     reinterpret_cast<const unsigned int *>(d + o_blk),
							     reinterpret_cast<uint32_t *>(d + o_out), (unsigned int)out_cap);
	BT_CUDA_TRY(cudaGetLastError());
	unsigned int total = 0;
	BT_CUDA_TRY(cudaMemcpyAsync(&total, d + o_tot, 4, cudaMemcpyDeviceToHost, st));
	BT_CUDA_TRY(cudaStreamSynchronize(st));
	*n_candidates = total;
	const size_t take = total < out_cap ? total : out_cap;
	if (take) BT_CUDA_TRY(cudaMemcpy(candidates, d + o_out, take * 4, cudaMemcpyDeviceToHost));
	if (survivors_after) {
		/* candidates left after observation j = all minus those that fell at observations 0..j */
		BT_CUDA_TRY(cudaMemcpy(survivors_after, d + o_hist, (size_t)n_obs * 4, cudaMemcpyDeviceToHost));
		int64_t left = N_CAND;
		for (int j = 0; j < n_obs; j++) { left -= (uint32_t)survivors_after[j]; survivors_after[j] = (int32_t)left; }
	}
	if ((int64_t)total > max_candidates) return btbb_b200_set_error(BTBB_B200_EOVERFLOW, "hop_winnow: candidate buffer too small");
	return BTBB_B200_OK;
}
