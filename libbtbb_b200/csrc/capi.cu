/*
 * capi.cu -- context life cycle, error reporting and the host-buffer entry points of the
 * C ABI declared in include/btbb_b200.h.
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <sched.h>
#include <time.h>
#include <atomic>
#include <new>
#include <thread>
#include <vector>
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
	ctx->opt_pack_streams = 4;      /* +3 % on the B200 host (132 -> 136.5 Gbit/s end to end): the 16 pack threads are then DRAM-bound */
	ctx->host_lock = new (std::nothrow) std::mutex();
	if (!ctx->host_lock) { free(ctx); return btbb_b200_set_error(BTBB_B200_ENOMEM, "create: out of host memory"); }
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
	btbb_b200_shard_destroy(ctx);
	{
		const bt_lane &o = ctx->parked;      /* the inactive lane's scratch */
		if (o.d_count) cudaFree(o.d_count);
		if (o.d_tmp) cudaFree(o.d_tmp);
		if (o.d_sort_hist) cudaFree(o.d_sort_hist);
		if (o.d_slab) cudaFree(o.d_slab);
		if (o.d_slab_cnt) cudaFree(o.d_slab_cnt);
		if (o.d_slab_base) cudaFree(o.d_slab_base);
		if (o.h_res) cudaFreeHost(o.h_res);
		if (o.ev_done) cudaEventDestroy(o.ev_done);
		for (int i = 0; i < 2; i++) if (o.prof_ev[i]) cudaEventDestroy(o.prof_ev[i]);
		if (ctx->ev_done) cudaEventDestroy(ctx->ev_done);
	}
	bt_tables_free(ctx);
	if (ctx->d_count) cudaFree(ctx->d_count);
	if (ctx->d_tmp) cudaFree(ctx->d_tmp);
	if (ctx->d_tmp2) cudaFree(ctx->d_tmp2);
	if (ctx->d_sort_hist) cudaFree(ctx->d_sort_hist);
	if (ctx->d_xp) cudaFree(ctx->d_xp);
	if (ctx->d_unpack) cudaFree(ctx->d_unpack);
	if (ctx->d_packed) cudaFree(ctx->d_packed);
	if (ctx->d_sieve_tc) cudaFree(ctx->d_sieve_tc);
	if (ctx->d_sieve_present) cudaFree(ctx->d_sieve_present);
	if (ctx->h_res) cudaFreeHost(ctx->h_res);
	if (ctx->d_sieve_idx) cudaFree(ctx->d_sieve_idx);
	if (ctx->d_sieve_cur) cudaFree(ctx->d_sieve_cur);
	for (int i = 0; i < 2; i++) if (ctx->h_pack[i]) cudaFreeHost(ctx->h_pack[i]);
	if (ctx->d_dec_tables) cudaFree(ctx->d_dec_tables);
	if (ctx->d_perm_tables) cudaFree(ctx->d_perm_tables);
	for (int i = 0; i < 4; i++) if (ctx->d_scratch[i]) cudaFree(ctx->d_scratch[i]);
	delete ctx->host_lock;
	if (ctx->ev_reset) cudaEventDestroy(ctx->ev_reset);
	for (int i = 0; i < 2; i++) if (ctx->prof_ev[i]) cudaEventDestroy(ctx->prof_ev[i]);
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
 * Host-buffer scan in the byte format.  The range is cut into chunks; chunk c is copied and
 * scanned on stream c&1, so the copy of one chunk overlaps the scan of the previous one.
 * Each chunk carries a 63-symbol tail so windows that straddle a seam are seen exactly once.
 * first_key != NULL selects first-hit mode (see push_hit in find_ac.cu).
 */
static int scan_host_prepare(btbb_b200_ctx *ctx, int64_t span, int64_t max_hits, bool first_key, int64_t *chunk_out)
{
	const int64_t CHUNK = (int64_t)64 << 20;
	int64_t chunk = span < CHUNK ? span : CHUNK;
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
	/* the scans run on the two non-blocking copy streams: reset the counter there, and make the
	 * second stream wait for it */
	BT_CUDA_TRY(cudaMemsetAsync(ctx->d_count, first_key ? 0xff : 0, sizeof(unsigned long long), ctx->copy_stream[0]));
	if (!ctx->ev_reset) BT_CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_reset, cudaEventDisableTiming));
	BT_CUDA_TRY(cudaEventRecord(ctx->ev_reset, ctx->copy_stream[0]));
	BT_CUDA_TRY(cudaStreamWaitEvent(ctx->copy_stream[1], ctx->ev_reset, 0));
	*chunk_out = chunk;
	return BTBB_B200_OK;
}

/* copies + scans of positions [0, span) of `stream`; nothing is waited for */
static int scan_host_enqueue(btbb_b200_ctx *ctx, const char *stream, int64_t span, int64_t chunk, uint32_t lap,
			     int max_ac_errors, int64_t max_hits, bool first_key)
{
	int c = 0;
	for (int64_t pos = 0; pos < span; pos += chunk, c ^= 1) {
		int64_t len = span - pos < chunk ? span - pos : chunk;
		cudaStream_t st = ctx->copy_stream[c];
		BT_CUDA_TRY(cudaMemcpyAsync(ctx->d_stage[c], stream + pos, (size_t)(len + 63), cudaMemcpyHostToDevice, st));
		int rc = bt_scan_launch(ctx, ctx->d_stage[c], len, lap, max_ac_errors, ctx->d_tmp,
					first_key ? -1 : max_hits, ctx->d_count, pos, st);
		if (rc) return rc;
	}
	return BTBB_B200_OK;
}

/* wait, order, bring the records back */
static int scan_host_finish(btbb_b200_ctx *ctx, int64_t span, btbb_b200_hit *hits, int64_t max_hits, int64_t *n_hits,
			    unsigned long long *first_key)
{
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
	int rc = bt_sort_hits(ctx, ctx->d_tmp, ctx->d_tmp2, have, bt_sort_passes(span), 0, ctx->copy_stream[0], &res);
	if (rc) return rc;
	if (have > 0)
		BT_CUDA_TRY(cudaMemcpyAsync(hits, res, (size_t)have * sizeof(btbb_b200_hit), cudaMemcpyDeviceToHost, ctx->copy_stream[0]));
	BT_CUDA_TRY(cudaStreamSynchronize(ctx->copy_stream[0]));
	if ((int64_t)total > max_hits)
		return btbb_b200_set_error(BTBB_B200_EOVERFLOW, "find_ac_host: hit buffer too small");
	return BTBB_B200_OK;
}

static int scan_host(btbb_b200_ctx *ctx, const char *stream, int64_t search_length, uint32_t lap,
		     int max_ac_errors, btbb_b200_hit *hits, int64_t max_hits, int64_t *n_hits,
		     unsigned long long *first_key)
{
	BT_CUDA_TRY(cudaSetDevice(ctx->device));
	if (ctx->lane_count) return btbb_b200_set_error(BTBB_B200_EINVAL, "find_ac_host: device scans are pending on this context");
	bt_lane_reset(ctx);
	int64_t chunk = 0;
	int rc = scan_host_prepare(ctx, search_length, max_hits, first_key != NULL, &chunk);
	if (rc) return rc;
	rc = scan_host_enqueue(ctx, stream, search_length, chunk, lap, max_ac_errors, max_hits, first_key != NULL);
	if (rc) return rc;
	return scan_host_finish(ctx, search_length, hits, max_hits, n_hits, first_key);
}

/* ---------------- host pack stage ----------------
 * The reference hands symbols over one per char (btbb.h:82-94); over PCIe that is 8 bits of
 * traffic per bit of information and the copy, not the scan, bounds a host-buffer call.  Large
 * calls therefore pack 32 symbols per word on the host (host_pack.cpp, a few cores), copy the
 * packed words chunk by chunk while the next chunk is being packed, and run the packed
 * variants of the bulk kernels.  This is a change of transfer format only: every decision is
 * still made on the device. */
namespace {

struct pack_job {
	const char *stream;
	int64_t limit;              /* readable symbols */
	int64_t chunk_words, nchunks, total_words;
	uint32_t *stage[2];
	int nthreads, streams;
	std::atomic<int> ready;     /* 0: workers wait, 1: barriers are set up, -1: give up */
	std::atomic<int64_t> next;  /* next block of the current chunk (blocks are handed out dynamically) */
	pthread_barrier_t start, done;
};

struct pack_arg { pack_job *job; int tid; };

void *pack_worker(void *p)
{
	pack_arg *pa = static_cast<pack_arg *>(p);
	pack_job *j = pa->job;
	int r;
	while ((r = j->ready.load(std::memory_order_acquire)) == 0) sched_yield();
	if (r < 0) return NULL;
	for (int64_t c = 0; c < j->nchunks; c++) {
		pthread_barrier_wait(&j->start);
		const int64_t w0 = c * j->chunk_words;
		const int64_t nw = j->total_words - w0 < j->chunk_words ? j->total_words - w0 : j->chunk_words;
		const int64_t BLOCK = 16384;     /* words: 512 Ki symbols per grab */
		for (;;) {
			const int64_t a = j->next.fetch_add(BLOCK, std::memory_order_relaxed);
			if (a >= nw) break;
			const int64_t b = a + BLOCK < nw ? a + BLOCK : nw;
			bt_pack_range_streams(j->stream, 32 * (w0 + a), b - a, j->limit, j->stage[c & 1] + a, j->streams);
		}
		pthread_barrier_wait(&j->done);
	}
	return NULL;
}

}  // namespace

/* CPUs this process may actually use: hardware threads, capped by a cgroup-v2 CPU quota */
static int usable_cpus()
{
	int n = (int)std::thread::hardware_concurrency();
	if (n < 1) n = 1;
	if (FILE *f = fopen("/sys/fs/cgroup/cpu.max", "r")) {
		long long quota = 0, period = 0;
		if (fscanf(f, "%lld %lld", &quota, &period) == 2 && quota > 0 && period > 0) {
			const int q = (int)((quota + period - 1) / period);
			if (q >= 1 && q < n) n = q;
		}
		fclose(f);
	}
	return n;
}

/* pack workers: every usable CPU up to 32 (BTBB_B200_OPT_PACK_THREADS overrides).  Measured on
 * the B200 host (2 x 32 cores, 16-CPU quota): ~5 GB/s of symbols per thread, and threads
 * beyond the quota only add throttling stalls at the per-chunk barriers. */
static int pack_threads(const btbb_b200_ctx *ctx)
{
	int nt = usable_cpus();
	if (nt > 32) nt = 32;
	if (ctx->opt_pack_threads > 0) nt = ctx->opt_pack_threads;
	if (nt < 1) nt = 1;
	if (nt > 128) nt = 128;
	return nt;
}

/* stream[0 .. nsym) -> ctx->d_packed (ceil(nsym / 32) words), copies issued on copy_stream[0] */
static int pack_and_upload(btbb_b200_ctx *ctx, const char *stream, int64_t nsym)
{
	const int64_t total_words = (nsym + 31) / 32;
	const int64_t CHUNK_WORDS = (int64_t)8 << 20;      /* 256 Mi symbols -> 32 MiB packed */
	for (int i = 0; i < 2; i++)
		if (!ctx->copy_stream[i])
			BT_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->copy_stream[i], cudaStreamNonBlocking));
	if (total_words + 2 > ctx->packed_cap) {
		if (ctx->d_packed) cudaFree(ctx->d_packed);
		ctx->d_packed = NULL; ctx->packed_cap = 0;
		BT_CUDA_TRY(cudaMalloc(&ctx->d_packed, (size_t)(total_words + 2) * sizeof(uint32_t)));
		ctx->packed_cap = total_words + 2;
	}
	const int64_t chunk_words = total_words < CHUNK_WORDS ? total_words : CHUNK_WORDS;
	if (chunk_words > ctx->h_pack_cap) {
		for (int i = 0; i < 2; i++) {
			if (ctx->h_pack[i]) cudaFreeHost(ctx->h_pack[i]);
			ctx->h_pack[i] = NULL;
		}
		ctx->h_pack_cap = 0;
		for (int i = 0; i < 2; i++)
			BT_CUDA_TRY(cudaMallocHost(&ctx->h_pack[i], (size_t)chunk_words * sizeof(uint32_t)));
		ctx->h_pack_cap = chunk_words;
	}
	pack_job job;
	job.stream = stream; job.limit = nsym;
	job.chunk_words = chunk_words; job.total_words = total_words;
	job.nchunks = (total_words + chunk_words - 1) / chunk_words;
	job.stage[0] = ctx->h_pack[0]; job.stage[1] = ctx->h_pack[1];
	job.streams = ctx->opt_pack_streams;
	int nt = pack_threads(ctx);
	if ((int64_t)nt > (chunk_words + 65535) / 65536) nt = (int)((chunk_words + 65535) / 65536);
	job.ready.store(0);
	std::vector<pthread_t> th((size_t)nt);
	std::vector<pack_arg> args((size_t)nt);
	int started = 0;
	for (; started < nt; started++) {
		args[(size_t)started].job = &job; args[(size_t)started].tid = started;
		if (pthread_create(&th[(size_t)started], NULL, pack_worker, &args[(size_t)started])) break;
	}
	if (started == 0)
		return btbb_b200_set_error(BTBB_B200_ENOMEM, "find_ac_host: cannot start pack threads");
	nt = started;                 /* fewer threads than asked for is fine */
	job.nthreads = nt;
	pthread_barrier_init(&job.start, NULL, (unsigned)nt + 1);
	pthread_barrier_init(&job.done, NULL, (unsigned)nt + 1);
	job.ready.store(1, std::memory_order_release);
	cudaEvent_t ev[2] = {NULL, NULL};
	cudaError_t e = cudaSuccess;
	for (int i = 0; i < 2 && e == cudaSuccess; i++) e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
	for (int64_t c = 0; c < job.nchunks; c++) {
		/* the staging buffer of chunk c was last used by the copy of chunk c - 2 */
		if (c >= 2 && e == cudaSuccess) e = cudaEventSynchronize(ev[c & 1]);
		job.next.store(0, std::memory_order_relaxed);
		pthread_barrier_wait(&job.start);
		pthread_barrier_wait(&job.done);
		const int64_t w0 = c * chunk_words;
		const int64_t nw = total_words - w0 < chunk_words ? total_words - w0 : chunk_words;
		if (e == cudaSuccess)
			e = cudaMemcpyAsync(ctx->d_packed + w0, job.stage[c & 1], (size_t)nw * sizeof(uint32_t),
					    cudaMemcpyHostToDevice, ctx->copy_stream[0]);
		if (e == cudaSuccess) e = cudaEventRecord(ev[c & 1], ctx->copy_stream[0]);
	}
	for (int i = 0; i < nt; i++) pthread_join(th[(size_t)i], NULL);
	pthread_barrier_destroy(&job.start); pthread_barrier_destroy(&job.done);
	for (int i = 0; i < 2; i++) if (ev[i]) cudaEventDestroy(ev[i]);
	if (e != cudaSuccess) return btbb_b200_cuda_fail(e, "find_ac_host: packed upload");
	return BTBB_B200_OK;
}

static double now_ms()
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

/*
 * Large host-buffer call: the host cores pack the stream while it is copied, the packed bulk
 * kernels scan it.  Optionally (BTBB_B200_OPT_HOST_SPLIT_PERMILLE) the head [0, split) travels
 * in the byte format instead -- DMA straight from a pinned buffer, no CPU work, scanned chunk
 * by chunk as it arrives -- while the cores pack the rest.  On the B200 host this was
 * measured as no gain (profiles/README.md: DMA and pack threads compete for the same host
 * memory bandwidth, ~80 GB/s in total), so the default split is 0; a host with spare memory
 * bandwidth would set it near 0.35.  Both parts partition the positions exactly, so the
 * concatenation of their hit lists is the ascending list.
 */
static int scan_host_packed(btbb_b200_ctx *ctx, const char *stream, int64_t search_length, uint32_t lap,
			    int max_ac_errors, btbb_b200_hit *hits, int64_t max_hits, int64_t *n_hits)
{
	BT_CUDA_TRY(cudaSetDevice(ctx->device));
	if (ctx->lane_count) return btbb_b200_set_error(BTBB_B200_EINVAL, "find_ac_host: device scans are pending on this context");
	bt_lane_reset(ctx);
	const bool trace = ctx->opt_trace != 0;
	const double t0 = trace ? now_ms() : 0;
	int64_t split = 0;
	{
		const double frac = ctx->opt_host_split / 1000.0;
		split = (int64_t)(frac * (double)search_length) & ~(int64_t)4095;
	}
	int rc;
	int64_t chunk = 0;
	if (split > 0) {
		rc = scan_host_prepare(ctx, split, max_hits, false, &chunk);
		if (rc) return rc;
		rc = scan_host_enqueue(ctx, stream, split, chunk, lap, max_ac_errors, max_hits, false);
		if (rc) return rc;
	}
	rc = pack_and_upload(ctx, stream + split, search_length - split + 63);
	if (rc) return rc;
	const double t1 = trace ? now_ms() : 0;
	int64_t n_a = 0;
	bool overflow = false;
	if (split > 0) {
		rc = scan_host_finish(ctx, split, hits, max_hits, &n_a, NULL);
		if (rc == BTBB_B200_EOVERFLOW) overflow = true;
		else if (rc) return rc;
	}
	const double t2 = trace ? now_ms() : 0;
	const int64_t have_a = n_a < max_hits ? n_a : max_hits;
	const int64_t room = max_hits - have_a;
	const int64_t cap = room > 0 ? room : 1;
	if (cap > ctx->tmp2_cap) {
		if (ctx->d_tmp2) cudaFree(ctx->d_tmp2);
		ctx->d_tmp2 = NULL; ctx->tmp2_cap = 0;
		BT_CUDA_TRY(cudaMalloc(&ctx->d_tmp2, (size_t)cap * sizeof(btbb_b200_hit)));
		ctx->tmp2_cap = cap;
	}
	cudaStream_t st = ctx->copy_stream[0];
	int64_t n_b = 0;
	const int64_t saved_bias = ctx->hit_bias;      /* btbb_b200_set_offset_bias is for the device entry points */
	ctx->hit_bias = 0;
	rc = bt_find_ac_dev_impl(ctx, reinterpret_cast<const uint8_t *>(ctx->d_packed), 1, search_length - split, lap,
				 max_ac_errors, ctx->d_tmp2, room, &n_b, st);
	ctx->hit_bias = saved_bias;
	if (rc == BTBB_B200_EOVERFLOW) overflow = true;
	else if (rc) return rc;
	const int64_t have_b = n_b < room ? n_b : room;
	if (have_b > 0)
		BT_CUDA_TRY(cudaMemcpyAsync(hits + have_a, ctx->d_tmp2, (size_t)have_b * sizeof(btbb_b200_hit), cudaMemcpyDeviceToHost, st));
	BT_CUDA_TRY(cudaStreamSynchronize(st));
	if (split > 0)
		for (int64_t i = 0; i < have_b; i++) hits[have_a + i].offset += split;
	*n_hits = n_a + n_b;
	if (trace)
		fprintf(stderr, "[btbb_b200] find_ac_host: %lld symbols as bytes + %lld packed (%d threads): issue+pack %.2f ms, "
			"byte part drained %.2f ms, packed scan+readback %.2f ms\n", (long long)split,
			(long long)(search_length - split), pack_threads(ctx), t1 - t0, t2 - t1, now_ms() - t2);
	if (overflow)
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
	if (search_length >= ((int64_t)4 << 20) && !ctx->opt_host_bytes)      /* BTBB_B200_OPT_HOST_BYTE_ROUTE keeps the byte-format copy */
		return scan_host_packed(ctx, stream, search_length, lap, max_ac_errors, hits, max_hits, n_hits);
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
