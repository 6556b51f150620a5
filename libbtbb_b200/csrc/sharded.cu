/*
 * sharded.cu -- the multi-GPU form of the access-code scan behind the C ABI (SURVEY.md 8e,
 * BASELINE configs[3]): one process per GPU, contiguous shards of the symbol stream, and the one
 * exchange the path has -- the variable-length all-gather of 16-byte hit records.
 *
 * Window positions are independent (bluetooth_packet.c:381-418, :430-439 carry no state from one
 * position to the next), so rank r scans positions [begin_r, end_r) and reads 63 symbols past
 * end_r (the north star fixes the overlap at 72); the kernels report global offsets
 * (btbb_b200_set_offset_bias), ranges partition the positions, and the rank-order concatenation
 * of the per-rank sorted lists is the globally sorted list the reference's iteration produces.
 *
 * Two forms of the exchange:
 *   peer memory, fused stores (default)  every rank owns a gather buffer of world slots x 2
 *       generations, mapped into every peer with CUDA IPC.  The ordering kernel of a scan
 *       (slab_sort_kernel, find_ac.cu) stores every record it places not only into the local list but
 *       also into this rank's slot on EVERY GPU -- plain stores to peer memory over NVLink -- and
 *       slab_scan_kernel writes the count header the same way: the all-gather is fused into the
 *       ordering pass, costs no launch, no copy descriptor and no host call, and is complete when the
 *       scan's stream is.
 *   peer memory, copy engines (BTBB_B200_SHARD_COPY_ENGINES; also the fallback for scans off the slab
 *       path: known LAP, very dense hits)  [count header | records] pushed into the slot on every peer
 *       with device-to-device copies on a copy stream that never waits for the scan stream, while the
 *       next scan (already queued: two may be pending) runs.  Nothing on a scan's own stream needs a
 *       copy engine (set-up by a kernel, counters through pinned host memory, find_ac.cu).
 *   Measured on 8 B200 at the bench's hit density (10^6 hits = 16 MB per rank and step, 112 MB in and
 *   112 MB out per GPU over NVLink): both forms cost the same, ~0.1 ms per 2.1 ms step.
 *   NCCL allgatherv         one ncclAllGather of the counts, then one group of ncclBroadcasts with
 *       exact sizes (NCCL has no native allgatherv).  Needs SMs, so it runs after the scan.
 * NCCL is loaded with dlopen (libnccl.so.2: the copy already in the process when the caller is a
 * PyTorch program, the system one otherwise) and is needed only for the bootstrap (IPC handle
 * exchange), the end-of-exchange barrier and the allgatherv form; libbtbb.so.1 itself does not
 * depend on it.
 */
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "capi_internal.h"

namespace {

struct nccl_api {
	void *handle;
	ncclResult_t (*GetUniqueId)(ncclUniqueId *);
	ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
	ncclResult_t (*CommDestroy)(ncclComm_t);
	ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
	ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
	ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
	ncclResult_t (*GroupStart)();
	ncclResult_t (*GroupEnd)();
	const char *(*GetErrorString)(ncclResult_t);
};
nccl_api g_nccl;
std::mutex g_nccl_lock;

int load_nccl()
{
	std::lock_guard<std::mutex> g(g_nccl_lock);
	if (g_nccl.handle) return BTBB_B200_OK;
	void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
	if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
	if (!h) return btbb_b200_set_error(BTBB_B200_ECUDA, "shard: libnccl.so.2 not found (needed for the multi-GPU entry points only)");
	nccl_api a;
	memset(&a, 0, sizeof(a));
	a.handle = h;
#define BT_SYM(field, name) *(void **)(&a.field) = dlsym(h, name); if (!a.field) { dlclose(h); return btbb_b200_set_error(BTBB_B200_ECUDA, "shard: " name " missing from libnccl"); }
	BT_SYM(GetUniqueId, "ncclGetUniqueId")
	BT_SYM(CommInitRank, "ncclCommInitRank")
	BT_SYM(CommDestroy, "ncclCommDestroy")
	BT_SYM(AllGather, "ncclAllGather")
	BT_SYM(Broadcast, "ncclBroadcast")
	BT_SYM(AllReduce, "ncclAllReduce")
	BT_SYM(GroupStart, "ncclGroupStart")
	BT_SYM(GroupEnd, "ncclGroupEnd")
	BT_SYM(GetErrorString, "ncclGetErrorString")
#undef BT_SYM
	g_nccl = a;
	return BTBB_B200_OK;
}

int nccl_fail(ncclResult_t r, const char *where)
{
	char msg[200];
	snprintf(msg, sizeof(msg), "NCCL error %d (%s) at %s", (int)r, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?", where);
	return btbb_b200_set_error(BTBB_B200_ECUDA, msg);
}
#define BT_NCCL_TRY(call) do { ncclResult_t r__ = (call); if (r__ != ncclSuccess) return nccl_fail(r__, #call); } while (0)

}  // namespace

constexpr int BT_SHARD_MAX_WORLD = 64;

struct bt_shard {
	int rank, world, peer;               /* peer = 1: peer-memory exchange available */
	int flags, last_fanned;
	int64_t slot;                        /* records per slot (without the header record) */
	ncclComm_t comm;
	cudaStream_t copy;                   /* the exchange's own stream: never waits for the scan stream */
	btbb_b200_hit *gather;               /* [2][world][slot + 1] */
	btbb_b200_hit *peers[BT_SHARD_MAX_WORLD];   /* the same buffer on every rank (IPC-mapped) */
	btbb_b200_hit *local[2];             /* [slot + 1]: header record, then this rank's sorted hits */
	btbb_b200_hit *h_hdr;                /* pinned: one header record per generation */
	unsigned long long *h_counts;        /* pinned: world x 2 words read back from the slot headers */
	cudaEvent_t sent[2];
	int sent_valid[2];
	int gen;                             /* generation the next begin() writes */
	int last;                            /* generation of the most recently ended scan, -1 = none */
	int pending;                         /* scans begun and not ended yet (0..2); the oldest one's generation is pend_g[0] */
	int pend_g[2];
	int64_t n_last;
	btbb_b200_hit **d_fan[2];            /* per generation: this rank's slot (first record) in every rank's gather buffer */
	int *d_flag;                         /* barrier scratch */
	int64_t *d_cnt;                      /* allgatherv: counts */
};

static size_t slot_bytes(const bt_shard *s) { return (size_t)(s->slot + 1) * sizeof(btbb_b200_hit); }
static btbb_b200_hit *slot_ptr(const bt_shard *s, btbb_b200_hit *base, int gen, int r)
{
	return base + ((size_t)gen * s->world + r) * (size_t)(s->slot + 1);
}

extern "C" int btbb_b200_shard_unique_id(void *id)
{
	if (!id) return btbb_b200_set_error(BTBB_B200_EINVAL, "shard_unique_id: null");
	int rc = load_nccl();
	if (rc) return rc;
	static_assert(sizeof(ncclUniqueId) == BTBB_B200_SHARD_ID_BYTES, "ncclUniqueId size");
	BT_NCCL_TRY(g_nccl.GetUniqueId(static_cast<ncclUniqueId *>(id)));
	return BTBB_B200_OK;
}

extern "C" int btbb_b200_shard_destroy(btbb_b200_ctx *ctx)
{
	if (!ctx || !ctx->shard) return BTBB_B200_OK;
	bt_shard *s = ctx->shard;
	cudaSetDevice(ctx->device);
	if (s->copy) cudaStreamSynchronize(s->copy);
	for (int r = 0; r < s->world; r++)
		if (r != s->rank && s->peers[r]) cudaIpcCloseMemHandle(s->peers[r]);
	if (s->gather) cudaFree(s->gather);
	for (int g = 0; g < 2; g++) {
		if (s->local[g]) cudaFree(s->local[g]);
		if (s->sent[g]) cudaEventDestroy(s->sent[g]);
	}
	if (s->h_hdr) cudaFreeHost(s->h_hdr);
	if (s->h_counts) cudaFreeHost(s->h_counts);
	for (int g = 0; g < 2; g++) if (s->d_fan[g]) cudaFree(s->d_fan[g]);
	if (s->d_flag) cudaFree(s->d_flag);
	if (s->d_cnt) cudaFree(s->d_cnt);
	if (s->copy) cudaStreamDestroy(s->copy);
	if (s->comm) g_nccl.CommDestroy(s->comm);
	free(s);
	ctx->shard = NULL;
	return BTBB_B200_OK;
}

extern "C" int btbb_b200_shard_init(btbb_b200_ctx *ctx, const void *id, int rank, int world, int64_t slot_records, int flags)
{
	if (!ctx || !id || world < 1 || world > BT_SHARD_MAX_WORLD || rank < 0 || rank >= world || slot_records < 1)
		return btbb_b200_set_error(BTBB_B200_EINVAL, "shard_init: bad arguments");
	if (ctx->shard) return btbb_b200_set_error(BTBB_B200_EINVAL, "shard_init: already initialised");
	int rc = load_nccl();
	if (rc) return rc;
	BT_CUDA_TRY(cudaSetDevice(ctx->device));
	bt_shard *s = (bt_shard *)calloc(1, sizeof(*s));
	if (!s) return btbb_b200_set_error(BTBB_B200_ENOMEM, "shard_init: out of host memory");
	ctx->shard = s;
	s->rank = rank; s->world = world; s->slot = slot_records; s->last = -1; s->flags = flags;
	ncclUniqueId uid;
	memcpy(&uid, id, sizeof(uid));
	ncclResult_t nr = g_nccl.CommInitRank(&s->comm, world, uid, rank);
	if (nr != ncclSuccess) { rc = nccl_fail(nr, "ncclCommInitRank"); btbb_b200_shard_destroy(ctx); return rc; }
#define BT_TRY_OR_DESTROY(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { \
	rc = btbb_b200_cuda_fail(e__, #call); btbb_b200_shard_destroy(ctx); return rc; } } while (0)
	BT_TRY_OR_DESTROY(cudaStreamCreateWithFlags(&s->copy, cudaStreamNonBlocking));
	BT_TRY_OR_DESTROY(cudaMalloc(&s->gather, 2 * (size_t)world * slot_bytes(s)));
	BT_TRY_OR_DESTROY(cudaMemset(s->gather, 0, 2 * (size_t)world * slot_bytes(s)));
	for (int g = 0; g < 2; g++) {
		BT_TRY_OR_DESTROY(cudaMalloc(&s->local[g], slot_bytes(s)));
		BT_TRY_OR_DESTROY(cudaMemset(s->local[g], 0, sizeof(btbb_b200_hit)));
		BT_TRY_OR_DESTROY(cudaEventCreateWithFlags(&s->sent[g], cudaEventDisableTiming));
	}
	BT_TRY_OR_DESTROY(cudaMallocHost(&s->h_hdr, 2 * sizeof(btbb_b200_hit)));
	BT_TRY_OR_DESTROY(cudaMallocHost(&s->h_counts, (size_t)world * 2 * sizeof(unsigned long long)));
	BT_TRY_OR_DESTROY(cudaMalloc(&s->d_flag, 2 * sizeof(int)));
	BT_TRY_OR_DESTROY(cudaMemset(s->d_flag, 0, 2 * sizeof(int)));
	BT_TRY_OR_DESTROY(cudaMalloc(&s->d_cnt, (size_t)(world + 1) * sizeof(int64_t)));
	s->peers[rank] = s->gather;
	s->peer = 0;
	if (!(flags & BTBB_B200_SHARD_NCCL_ONLY) && world > 1) {
		/* map every rank's gather buffer: all-gather the IPC handles, open the peers' */
		cudaIpcMemHandle_t mine, *all_h = NULL;
		void *d_h = NULL;
		BT_TRY_OR_DESTROY(cudaIpcGetMemHandle(&mine, s->gather));
		BT_TRY_OR_DESTROY(cudaMalloc(&d_h, (size_t)(world + 1) * sizeof(mine)));
		char *d_all = (char *)d_h, *d_mine = d_all + (size_t)world * sizeof(mine);
		BT_TRY_OR_DESTROY(cudaMemcpy(d_mine, &mine, sizeof(mine), cudaMemcpyHostToDevice));
		nr = g_nccl.AllGather(d_mine, d_all, sizeof(mine), ncclChar, s->comm, s->copy);
		if (nr != ncclSuccess) { cudaFree(d_h); rc = nccl_fail(nr, "ncclAllGather(ipc handles)"); btbb_b200_shard_destroy(ctx); return rc; }
		all_h = (cudaIpcMemHandle_t *)malloc((size_t)world * sizeof(mine));
		cudaError_t e = cudaStreamSynchronize(s->copy);
		if (e == cudaSuccess) e = cudaMemcpy(all_h, d_all, (size_t)world * sizeof(mine), cudaMemcpyDeviceToHost);
		cudaFree(d_h);
		int ok = e == cudaSuccess;
		for (int r = 0; ok && r < world; r++) {
			if (r == rank) continue;
			void *p = NULL;
			if (cudaIpcOpenMemHandle(&p, all_h[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); break; }
			s->peers[r] = (btbb_b200_hit *)p;
		}
		free(all_h);
		/* every rank must take the same route */
		int *d_ok = s->d_flag;
		int h_ok = ok;
		BT_TRY_OR_DESTROY(cudaMemcpy(d_ok, &h_ok, sizeof(int), cudaMemcpyHostToDevice));
		nr = g_nccl.AllReduce(d_ok, d_ok + 1, 1, ncclInt, ncclMin, s->comm, s->copy);
		if (nr != ncclSuccess) { rc = nccl_fail(nr, "ncclAllReduce(peer ok)"); btbb_b200_shard_destroy(ctx); return rc; }
		BT_TRY_OR_DESTROY(cudaStreamSynchronize(s->copy));
		BT_TRY_OR_DESTROY(cudaMemcpy(&h_ok, d_ok + 1, sizeof(int), cudaMemcpyDeviceToHost));
		s->peer = h_ok;
	} else if (world == 1)
		s->peer = 1;
	if (s->peer) {
		/* fan-out lists of the ordering kernel: slot [g][rank] on every GPU, records start behind the header */
		for (int g = 0; g < 2; g++) {
			btbb_b200_hit *h_fan[BT_SHARD_MAX_WORLD];
			for (int r = 0; r < world; r++) h_fan[r] = slot_ptr(s, s->peers[r], g, rank) + 1;
			BT_TRY_OR_DESTROY(cudaMalloc(&s->d_fan[g], (size_t)world * sizeof(btbb_b200_hit *)));
			BT_TRY_OR_DESTROY(cudaMemcpy(s->d_fan[g], h_fan, (size_t)world * sizeof(btbb_b200_hit *), cudaMemcpyHostToDevice));
		}
	}
#undef BT_TRY_OR_DESTROY
	return BTBB_B200_OK;
}

extern "C" int btbb_b200_shard_info(const btbb_b200_ctx *ctx, int *rank, int *world, int *peer_memory)
{
	if (!ctx || !ctx->shard) return btbb_b200_set_error(BTBB_B200_EINVAL, "shard: not initialised");
	if (rank) *rank = ctx->shard->rank;
	if (world) *world = ctx->shard->world;
	if (peer_memory) *peer_memory = ctx->shard->peer;
	return BTBB_B200_OK;
}

/* all ranks meet; everything enqueued on every rank's copy stream before the barrier is done after it */
static int shard_barrier(bt_shard *s)
{
	if (s->world > 1)
		BT_NCCL_TRY(g_nccl.AllReduce(s->d_flag, s->d_flag + 1, 1, ncclInt, ncclSum, s->comm, s->copy));
	BT_CUDA_TRY(cudaStreamSynchronize(s->copy));
	return BTBB_B200_OK;
}

extern "C" int btbb_b200_find_ac_sharded_begin(btbb_b200_ctx *ctx, const uint8_t *d_stream, int64_t search_length,
					       int64_t first_position, uint32_t lap, int max_ac_errors, void *cuda_stream)
{
	if (!ctx || !ctx->shard) return btbb_b200_set_error(BTBB_B200_EINVAL, "find_ac_sharded: btbb_b200_shard_init first");
	bt_shard *s = ctx->shard;
	if (s->pending >= 2) return btbb_b200_set_error(BTBB_B200_EINVAL, "find_ac_sharded: two scans are already pending");
	if ((!d_stream && search_length > 0) || search_length < 0 || (lap != BTBB_B200_LAP_ANY && lap > 0xffffffu))
		return btbb_b200_set_error(BTBB_B200_EINVAL, "find_ac_sharded: bad arguments");
	BT_CUDA_TRY(cudaSetDevice(ctx->device));
	const int g = s->gen;
	/* the copies that read this generation's buffer two steps ago have long left; make sure */
	if (s->sent_valid[g]) BT_CUDA_TRY(cudaEventSynchronize(s->sent[g]));
	const int64_t saved = ctx->hit_bias;
	ctx->hit_bias = first_position;
	/* default: the ordering kernel stores every record into this rank's slot on every GPU as it writes the sorted
	 * list (promiscuous bulk path; other paths, and BTBB_B200_SHARD_COPY_ENGINES, push with the copy engines in _end) */
	if (s->peer && !(s->flags & BTBB_B200_SHARD_COPY_ENGINES)) { ctx->d_fan = s->d_fan[g]; ctx->fan_n = s->world; }
	int rc = bt_find_ac_dev_begin(ctx, d_stream, 0, search_length, lap, max_ac_errors, s->local[g] + 1, s->slot, (cudaStream_t)cuda_stream);
	ctx->d_fan = NULL; ctx->fan_n = 0;
	ctx->hit_bias = saved;
	if (rc) return rc;
	s->pend_g[s->pending++] = g;
	s->gen ^= 1;
	return BTBB_B200_OK;
}

/* wait for the pending scan; *g_out = the generation it filled */
static int shard_end_wait(btbb_b200_ctx *ctx, int64_t *n_local, int *g_out)
{
	bt_shard *s = ctx->shard;
	if (!s->pending) return btbb_b200_set_error(BTBB_B200_EINVAL, "find_ac_sharded_end: no scan is pending");
	const int g = s->pend_g[0];
	s->pend_g[0] = s->pend_g[1];
	s->pending--;
	int64_t n = 0;
	int rc = bt_find_ac_dev_end(ctx, &n);      /* waits for the oldest pending scan and its ordering pass */
	*n_local = n;
	if (rc) return rc;                         /* EOVERFLOW included: the slot is too small for this shard */
	*g_out = g;
	s->last = g; s->n_last = n;
	s->last_fanned = ctx->last_fanned;
	return BTBB_B200_OK;
}

/* header record, then one device-to-device copy per rank: all on the copy stream, which does not wait
 * for whatever has been enqueued on the scan stream since */
static int shard_push(btbb_b200_ctx *ctx, int g, int64_t n)
{
	bt_shard *s = ctx->shard;
	if (!s->peer) return BTBB_B200_OK;         /* the NCCL form moves the records in _gather */
	if (s->last_fanned) return BTBB_B200_OK;   /* the ordering kernel has already delivered them */
	memset(&s->h_hdr[g], 0, sizeof(btbb_b200_hit));
	s->h_hdr[g].offset = n;
	BT_CUDA_TRY(cudaMemcpyAsync(s->local[g], &s->h_hdr[g], sizeof(btbb_b200_hit), cudaMemcpyHostToDevice, s->copy));
	for (int r = 0; r < s->world; r++)
		BT_CUDA_TRY(cudaMemcpyAsync(slot_ptr(s, s->peers[r], g, s->rank), s->local[g], (size_t)(n + 1) * sizeof(btbb_b200_hit),
					    cudaMemcpyDeviceToDevice, s->copy));
	BT_CUDA_TRY(cudaEventRecord(s->sent[g], s->copy));
	s->sent_valid[g] = 1;
	return BTBB_B200_OK;
}

extern "C" int btbb_b200_find_ac_sharded_end(btbb_b200_ctx *ctx, int64_t *n_local)
{
	if (!ctx || !ctx->shard || !n_local) return btbb_b200_set_error(BTBB_B200_EINVAL, "find_ac_sharded_end: bad arguments");
	int g = 0;
	int rc = shard_end_wait(ctx, n_local, &g);
	if (rc) return rc;
	return shard_push(ctx, g, *n_local);
}

/* _end of the pending scan and _begin of the next one in one call, in the order that keeps the GPU
 * busy: wait for scan i, enqueue scan i + 1, THEN start pushing the records of scan i (copy
 * engines, underneath scan i + 1) */
extern "C" int btbb_b200_find_ac_sharded_next(btbb_b200_ctx *ctx, const uint8_t *d_stream, int64_t search_length,
					      int64_t first_position, uint32_t lap, int max_ac_errors, void *cuda_stream,
					      int64_t *n_prev)
{
	if (!ctx || !ctx->shard || !n_prev) return btbb_b200_set_error(BTBB_B200_EINVAL, "find_ac_sharded_next: bad arguments");
	int g = 0;
	int rc = shard_end_wait(ctx, n_prev, &g);
	if (rc) return rc;
	rc = btbb_b200_find_ac_sharded_begin(ctx, d_stream, search_length, first_position, lap, max_ac_errors, cuda_stream);
	if (rc) return rc;
	return shard_push(ctx, g, *n_prev);
}

extern "C" int btbb_b200_find_ac_sharded_gather(btbb_b200_ctx *ctx, const btbb_b200_hit **d_slots, int64_t *slot_stride,
						int64_t *counts, int64_t *n_total)
{
	if (!ctx || !ctx->shard || !counts) return btbb_b200_set_error(BTBB_B200_EINVAL, "find_ac_sharded_gather: bad arguments");
	bt_shard *s = ctx->shard;
	if (s->last < 0) return btbb_b200_set_error(BTBB_B200_EINVAL, "find_ac_sharded_gather: nothing to gather");
	BT_CUDA_TRY(cudaSetDevice(ctx->device));
	const int g = s->last;
	btbb_b200_hit *mine = slot_ptr(s, s->gather, g, 0);
	int rc;
	if (s->peer) {
		rc = shard_barrier(s);      /* my pushes are done (stream order), then everybody's */
		if (rc) return rc;
	} else {
		/* NCCL allgatherv: counts, then one group of broadcasts with exact sizes */
		int64_t n = s->n_last;
		BT_CUDA_TRY(cudaMemcpyAsync(s->d_cnt + s->world, &n, sizeof(n), cudaMemcpyHostToDevice, s->copy));
		BT_NCCL_TRY(g_nccl.AllGather(s->d_cnt + s->world, s->d_cnt, 1, ncclInt64, s->comm, s->copy));
		BT_CUDA_TRY(cudaMemcpyAsync(s->h_counts, s->d_cnt, (size_t)s->world * sizeof(int64_t), cudaMemcpyDeviceToHost, s->copy));
		BT_CUDA_TRY(cudaStreamSynchronize(s->copy));
		BT_NCCL_TRY(g_nccl.GroupStart());
		for (int r = 0; r < s->world; r++) {
			const int64_t c = (int64_t)s->h_counts[r];
			if (c > s->slot) { g_nccl.GroupEnd(); return btbb_b200_set_error(BTBB_B200_EOVERFLOW, "find_ac_sharded_gather: a rank found more hits than a slot holds"); }
			if (c == 0) continue;
			ncclResult_t nr = g_nccl.Broadcast(s->local[g] + 1, slot_ptr(s, s->gather, g, r) + 1, (size_t)c * sizeof(btbb_b200_hit),
							   ncclChar, r, s->comm, s->copy);
			if (nr != ncclSuccess) { g_nccl.GroupEnd(); return nccl_fail(nr, "ncclBroadcast(hit records)"); }
		}
		BT_NCCL_TRY(g_nccl.GroupEnd());
		BT_CUDA_TRY(cudaStreamSynchronize(s->copy));
	}
	if (s->peer) {
		/* the counts sit in the header record of every slot */
		BT_CUDA_TRY(cudaMemcpy2DAsync(s->h_counts, sizeof(unsigned long long), mine, slot_bytes(s), sizeof(unsigned long long),
					      (size_t)s->world, cudaMemcpyDeviceToHost, s->copy));
		BT_CUDA_TRY(cudaStreamSynchronize(s->copy));
	}
	int64_t total = 0;
	for (int r = 0; r < s->world; r++) { counts[r] = (int64_t)s->h_counts[r]; total += counts[r]; }
	if (n_total) *n_total = total;
	if (d_slots) *d_slots = mine + 1;                 /* rank r's records: (*d_slots) + r * (*slot_stride) */
	if (slot_stride) *slot_stride = s->slot + 1;
	return BTBB_B200_OK;
}

extern "C" int btbb_b200_find_ac_sharded_dev(btbb_b200_ctx *ctx, const uint8_t *d_stream, int64_t search_length,
					     int64_t first_position, uint32_t lap, int max_ac_errors,
					     btbb_b200_hit *d_all, int64_t max_all, int64_t *counts, int64_t *n_total, void *cuda_stream)
{
	if (!ctx || !ctx->shard || !counts || !n_total || (!d_all && max_all > 0) || max_all < 0)
		return btbb_b200_set_error(BTBB_B200_EINVAL, "find_ac_sharded: bad arguments");
	bt_shard *s = ctx->shard;
	int rc = btbb_b200_find_ac_sharded_begin(ctx, d_stream, search_length, first_position, lap, max_ac_errors, cuda_stream);
	if (rc) return rc;
	int64_t n_local = 0;
	rc = btbb_b200_find_ac_sharded_end(ctx, &n_local);
	/* a rank that failed still has to meet the others in the gather; report its error afterwards */
	const int rc_local = rc;
	if (rc_local) { s->last = s->gen ^ 1; s->n_last = 0; s->last_fanned = 0; }
	const btbb_b200_hit *slots = NULL;
	int64_t stride = 0;
	rc = btbb_b200_find_ac_sharded_gather(ctx, &slots, &stride, counts, n_total);
	if (rc_local) return rc_local;
	if (rc) return rc;
	int64_t at = 0;
	for (int r = 0; r < s->world; r++) {
		const int64_t take = at + counts[r] <= max_all ? counts[r] : (max_all > at ? max_all - at : 0);
		if (take > 0)
			BT_CUDA_TRY(cudaMemcpyAsync(d_all + at, slots + (size_t)r * stride, (size_t)take * sizeof(btbb_b200_hit),
						    cudaMemcpyDeviceToDevice, s->copy));
		at += counts[r];
	}
	BT_CUDA_TRY(cudaStreamSynchronize(s->copy));
	if (*n_total > max_all) return btbb_b200_set_error(BTBB_B200_EOVERFLOW, "find_ac_sharded: hit buffer too small");
	return BTBB_B200_OK;
}
