/*
 * capi.cu -- context life cycle, error reporting and the host-buffer entry points of the
 * C ABI declared in include/btbb_b200.h.
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "capi_internal.h"

static __thread char g_err[256] = "";

int btbb_b200_set_error(int code, const char *msg)
{
	snprintf(g_err, sizeof(g_err), "%s", msg ? msg : "");
	return code;
}

int btbb_b200_cuda_fail(cudaError_t e, const char *where)
{
	snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), where);
	return e == cudaErrorMemoryAllocation ? BTBB_B200_ENOMEM : BTBB_B200_ECUDA;
}

extern "C" const char *btbb_b200_last_error(void) { return g_err; }

extern "C" int btbb_b200_create(int device, int max_ac_errors, btbb_b200_ctx **out)
{
	if (!out) return btbb_b200_set_error(BTBB_B200_EINVAL, "create: null ctx pointer");
	*out = NULL;
	/* same range check as btbb_init (bluetooth_packet.c:282-286) */
	if (max_ac_errors < 0 || max_ac_errors > 5)
		return btbb_b200_set_error(BTBB_B200_EINVAL, "create: max_ac_errors out of range");
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev == 0)
		return btbb_b200_set_error(BTBB_B200_ECUDA, "create: no CUDA device (this library has no CPU path)");
	if (device < 0 || device >= ndev)
		return btbb_b200_set_error(BTBB_B200_EINVAL, "create: bad device index");
	BT_CUDA_TRY(cudaSetDevice(device));
	cudaDeviceProp prop;
	BT_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
	if (prop.major != 10)
		return btbb_b200_set_error(BTBB_B200_ECUDA, "create: device is not sm_100 (kernels are built for sm_100a only)");
	btbb_b200_ctx *ctx = (btbb_b200_ctx *)calloc(1, sizeof(*ctx));
	if (!ctx) return btbb_b200_set_error(BTBB_B200_ENOMEM, "create: out of host memory");
	ctx->device = device;
	ctx->sm_count = prop.multiProcessorCount;
	int rc = bt_tables_build(ctx, max_ac_errors);
	if (rc == BTBB_B200_OK) {
		e = cudaMalloc(&ctx->d_count, 2 * sizeof(unsigned long long));
		if (e != cudaSuccess) rc = btbb_b200_cuda_fail(e, "cudaMalloc(count)");
	}
	if (rc != BTBB_B200_OK) { btbb_b200_destroy(ctx); return rc; }
	*out = ctx;
	return BTBB_B200_OK;
}

extern "C" void btbb_b200_destroy(btbb_b200_ctx *ctx)
{
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	bt_tables_free(ctx);
	if (ctx->d_count) cudaFree(ctx->d_count);
	if (ctx->d_tmp) cudaFree(ctx->d_tmp);
	if (ctx->d_tmp2) cudaFree(ctx->d_tmp2);
	if (ctx->d_sort_hist) cudaFree(ctx->d_sort_hist);
	if (ctx->d_xp) cudaFree(ctx->d_xp);
	if (ctx->d_dbg) cudaFree(ctx->d_dbg);
	if (ctx->d_slab) cudaFree(ctx->d_slab);
	if (ctx->d_slab_cnt) cudaFree(ctx->d_slab_cnt);
	if (ctx->d_slab_base) cudaFree(ctx->d_slab_base);
	for (int i = 0; i < 2; i++) {
		if (ctx->d_stage[i]) cudaFree(ctx->d_stage[i]);
		if (ctx->copy_stream[i]) cudaStreamDestroy(ctx->copy_stream[i]);
	}
	free(ctx);
}

extern "C" int btbb_b200_device(const btbb_b200_ctx *ctx) { return ctx ? ctx->device : -1; }
extern "C" int btbb_b200_table_errors(const btbb_b200_ctx *ctx) { return ctx ? ctx->table_k : -1; }

static int ensure_stage(btbb_b200_ctx *ctx, int64_t bytes)
{
	for (int i = 0; i < 2; i++)
		if (!ctx->copy_stream[i])
			BT_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->copy_stream[i], cudaStreamNonBlocking));
	if (bytes > ctx->stage_cap) {
		for (int i = 0; i < 2; i++) {
			if (ctx->d_stage[i]) cudaFree(ctx->d_stage[i]);
			ctx->d_stage[i] = NULL;
		}
		ctx->stage_cap = 0;
		for (int i = 0; i < 2; i++)
			BT_CUDA_TRY(cudaMalloc(&ctx->d_stage[i], (size_t)bytes));
		ctx->stage_cap = bytes;
	}
	return BTBB_B200_OK;
}

/*
 * Host-buffer scan.  The stream is cut into chunks; chunk c is copied and scanned on
 * stream c&1, so the copy of one chunk overlaps the scan of the previous one.  Each chunk
 * carries a 63-symbol tail so windows that straddle a seam are seen exactly once.
 * first_key != NULL selects first-hit mode (see push_hit in find_ac.cu).
 */
static int scan_host(btbb_b200_ctx *ctx, const char *stream, int64_t search_length, uint32_t lap,
		     int max_ac_errors, btbb_b200_hit *hits, int64_t max_hits, int64_t *n_hits,
		     unsigned long long *first_key)
{
	BT_CUDA_TRY(cudaSetDevice(ctx->device));
	const int64_t CHUNK = (int64_t)64 << 20;
	int64_t chunk = search_length < CHUNK ? search_length : CHUNK;
	int rc = ensure_stage(ctx, chunk + 64);
	if (rc) return rc;
	if (!first_key) {
		int64_t cap = max_hits > 0 ? max_hits : 1;
		rc = bt_ensure_tmp(ctx, cap);
		if (rc) return rc;
		if (cap > ctx->tmp2_cap) {
			if (ctx->d_tmp2) cudaFree(ctx->d_tmp2);
			ctx->d_tmp2 = NULL; ctx->tmp2_cap = 0;
			BT_CUDA_TRY(cudaMalloc(&ctx->d_tmp2, (size_t)cap * sizeof(btbb_b200_hit)));
			ctx->tmp2_cap = cap;
		}
	}
	BT_CUDA_TRY(cudaMemset(ctx->d_count, first_key ? 0xff : 0, sizeof(unsigned long long)));
	int c = 0;
	for (int64_t pos = 0; pos < search_length; pos += chunk, c ^= 1) {
		int64_t len = search_length - pos < chunk ? search_length - pos : chunk;
		cudaStream_t st = ctx->copy_stream[c];
		BT_CUDA_TRY(cudaMemcpyAsync(ctx->d_stage[c], stream + pos, (size_t)(len + 63), cudaMemcpyHostToDevice, st));
		rc = bt_scan_launch(ctx, ctx->d_stage[c], len, lap, max_ac_errors, ctx->d_tmp,
				    first_key ? -1 : max_hits, ctx->d_count, pos, st);
		if (rc) return rc;
	}
	BT_CUDA_TRY(cudaStreamSynchronize(ctx->copy_stream[0]));
	BT_CUDA_TRY(cudaStreamSynchronize(ctx->copy_stream[1]));
	unsigned long long total = 0;
	BT_CUDA_TRY(cudaMemcpy(&total, ctx->d_count, sizeof(total), cudaMemcpyDeviceToHost));
	if (first_key) { *first_key = total; return BTBB_B200_OK; }
	if (total >> 62)
		return btbb_b200_set_error(BTBB_B200_ECUDA, "find_ac_host: unexpected shared-memory window layout");
	*n_hits = (int64_t)total;
	int64_t have = (int64_t)total < max_hits ? (int64_t)total : max_hits;
	btbb_b200_hit *res = NULL;
	rc = bt_sort_hits(ctx, ctx->d_tmp, ctx->d_tmp2, have, bt_sort_passes(search_length), ctx->copy_stream[0], &res);
	if (rc) return rc;
	if (have > 0)
		BT_CUDA_TRY(cudaMemcpyAsync(hits, res, (size_t)have * sizeof(btbb_b200_hit), cudaMemcpyDeviceToHost, ctx->copy_stream[0]));
	BT_CUDA_TRY(cudaStreamSynchronize(ctx->copy_stream[0]));
	if ((int64_t)total > max_hits)
		return btbb_b200_set_error(BTBB_B200_EOVERFLOW, "find_ac_host: hit buffer too small");
	return BTBB_B200_OK;
}

extern "C" int btbb_b200_find_ac_host(btbb_b200_ctx *ctx, const char *stream, int64_t search_length,
				      uint32_t lap, int max_ac_errors, btbb_b200_hit *hits,
				      int64_t max_hits, int64_t *n_hits)
{
	if (!ctx || !n_hits || (!hits && max_hits > 0) || (!stream && search_length > 0) ||
	    search_length < 0 || max_hits < 0 || (lap != BTBB_B200_LAP_ANY && lap > 0xffffffu))
		return btbb_b200_set_error(BTBB_B200_EINVAL, "find_ac_host: bad arguments");
	*n_hits = 0;
	if (search_length == 0) return BTBB_B200_OK;
	return scan_host(ctx, stream, search_length, lap, max_ac_errors, hits, max_hits, n_hits, NULL);
}

/* first hit only: what the classic btbb_find_ac() needs (bluetooth_packet.c:444-464) */
int bt_find_first_host(btbb_b200_ctx *ctx, const char *stream, int search_length, uint32_t lap,
		       int max_ac_errors, btbb_b200_hit *hit, int *found)
{
	unsigned long long key = ~0ULL;
	*found = 0;
	if (search_length <= 0) return BTBB_B200_OK;
	int rc = scan_host(ctx, stream, search_length, lap, max_ac_errors, NULL, 0, NULL, &key);
	if (rc) return rc;
	if (key != ~0ULL) {
		*found = 1;
		hit->offset = (int64_t)(key >> 32);
		hit->lap = (uint32_t)(key >> 8) & 0xffffffu;
		hit->ac_errors = (uint8_t)(key & 0xff);
		hit->pad[0] = hit->pad[1] = hit->pad[2] = 0;
	}
	return BTBB_B200_OK;
}
