/* capi_internal.h -- state shared by the translation units of libbtbb.so.1 (B200 build). */
#ifndef BTBB_B200_CAPI_INTERNAL_H
#define BTBB_B200_CAPI_INTERNAL_H

#include <stdint.h>
#include <mutex>
#include <cuda_runtime.h>
#include "../../include/btbb_b200.h"

/* Tables the access-code scan kernels stage into shared memory. */
struct bt_scan_tables {
	uint32_t t_a[256];     /* low 32 syndrome bits of a byte at codeword bits 32..39 */
	uint32_t t_b[256];     /* ... bits 40..47 */
	uint32_t t_c[256];     /* ... bits 48..55 */
	uint32_t t_56;         /* ... bit 56 */
	uint32_t c_class[2];   /* syndrome of (PN ^ corrected Barker tail) for tail A / tail B */
	uint32_t pad;
};

/* syndrome -> error pattern, open addressing in global memory (replaces the uthash map of
 * bluetooth_packet.c:121-145).  Empty slots have syn == 0 (no error pattern has a zero syndrome). */
struct bt_err_slot { uint64_t syn, err; };

/* one find_ac call between its begin() and end() halves (find_ac.cu) */
enum { BT_PENDING_NONE = 0, BT_PENDING_EMPTY, BT_PENDING_SLAB, BT_PENDING_UNORDERED, BT_PENDING_GENERIC };
struct bt_pending {
	int mode;
	const uint8_t *d_stream;
	int packed, max_ac_errors;
	int64_t search_length, max_hits, bias;
	uint32_t lap;
	btbb_b200_hit *d_hits;
	cudaStream_t st;
	int fanned;          /* the ordering pass also stored the records into ctx->d_fan's lists */
};

/* Per-scan scratch of the device entry points.  Two scans may be pending on one context (the second is
 * enqueued before the first is waited for, so the GPU never idles between them): the context's own
 * fields of the same names are the ACTIVE lane's, the other lane's values are parked here
 * (bt_lane_select, find_ac.cu). */
struct bt_lane {
	unsigned long long *d_count;
	btbb_b200_hit *d_tmp;
	int64_t tmp_cap;
	uint32_t *d_sort_hist;
	int64_t sort_hist_cap;
	btbb_b200_hit *d_slab;
	uint32_t *d_slab_cnt;
	unsigned long long *d_slab_base;
	int slab_n;
	unsigned long long *h_res;
	bt_pending pending;
	cudaEvent_t ev_done;         /* everything begin() enqueued for this scan has finished */
	cudaEvent_t prof_ev[2];
	int prof_valid;
};

struct bt_shard;

struct btbb_b200_ctx {
	int device;
	int table_k;                 /* tables hold every pattern of 1..table_k errors in bits 0..57 */
	int sm_count;
	bt_scan_tables *d_tables;    /* device copy */
	uint32_t *d_bloom;           /* bitmap over the low 32 syndrome bits of table entries (+ zero) */
	int bloom_log2;              /* log2(bits) */
	uint64_t cc[2];              /* 34-bit syndrome of PN ^ (legal tail << 57) */
	uint32_t m32, m33;           /* parity masks over codeword bits 32..56 for syndrome bits 32 / 33 */
	uint32_t m0;                 /* same for syndrome bit 0 */
	uint32_t *d_lut7;            /* bulk kernel v7: field tables over codeword bits 34..40 / 41..48 / 49..56 -> syndrome bits 1..32 */
	uint32_t *d_map7;            /* bulk kernel v7: 2^16-byte first-level map + 2^17-bit second-level map */
	uint32_t *d_map7b;           /* ... with a 2^15-byte first-level map (layout<1>) */
	uint32_t *d_map7g;           /* tables for 3 errors: 2^27-bit second-level map; 4 / 5 errors: 2^27 / 2^29-bit first-level map */
	int map7g_log2;
	int l2_persist_set;          /* cudaLimitPersistingL2CacheSize configured for the global first-level map */
	size_t l2_persist_bytes;
	btbb_b200_hit *d_slab;       /* slab ordering: (warps + 2) x BT_SLAB_CAP records */
	uint32_t *d_slab_cnt;        /* per-slab fill counts (+ two 64-bit edge counters) */
	unsigned long long *d_slab_base;
	int slab_n;
	void *d_xp;                  /* ring of 16 x 128-byte exact-test parameter blocks (bulk kernel) */
	unsigned xp_next;
	bt_err_slot *d_err;          /* hash table, capacity 1 << err_log2 (NULL when table_k == 0) */
	bt_err_slot *h_err;          /* host copy (small-call path of the classic btbb_find_ac) */
	int err_log2;
	long err_entries;
	/* scratch owned by the context */
	unsigned long long *d_count; /* hit counter */
	btbb_b200_hit *d_tmp;        /* unordered hits before the ordering pass */
	btbb_b200_hit *d_tmp2;       /* second buffer for the host-buffer entry point */
	int64_t tmp2_cap;
	int64_t tmp_cap;
	uint32_t *d_sort_hist;       /* radix-sort histograms */
	int64_t sort_hist_cap;
	uint8_t *d_stage[2];         /* device staging for the host-buffer entry points */
	uint8_t *d_unpack;           /* packed input: bytes of the ragged tail (or of the whole stream when no bulk kernel applies) */
	int64_t unpack_cap;
	uint32_t *d_packed;          /* host entry point: the packed stream on the device */
	int64_t packed_cap;          /* words */
	uint32_t *h_pack[2];         /* pinned staging for the host pack stage */
	int64_t h_pack_cap;          /* words each */
	int64_t stage_cap;
	cudaStream_t copy_stream[2];
	bt_pending pending;
	int64_t hit_bias;            /* added to every offset the device entry points report (btbb_b200_set_offset_bias) */
	unsigned long long *h_res;   /* pinned: {hit count, slab-overflow flag} of the pending call */
	uint16_t *d_sieve_tc;        /* UAP sieve: (packet, clock) table, 64 words per packet */
	uint8_t *d_sieve_present;    /* UAP sieve: btbb_header_present per packet */
	int64_t sieve_cap;           /* packets the two buffers hold */
	int64_t *d_sieve_idx;        /* UAP sieve: packets of the current round (sieve_cap entries) */
	int64_t *d_sieve_cur;        /* UAP sieve: per-piconet cursor, then one 64-bit counter */
	int64_t sieve_groups_cap;
	void *d_perm_tables;         /* hop sequence: the factored perm5 tables (hops.cu) */
	void *d_dec_tables;          /* per-packet chain: device copy of btd_tables (decode_core.h) */
	void *d_scratch[4];          /* grow-only device scratch of the host-buffer entry points */
	size_t scratch_cap[4];
	cudaEvent_t ev_reset;        /* host-buffer scan: the counter reset has been enqueued */
	int opt_tile_only, opt_host_bytes, opt_host_split, opt_pack_threads, opt_trace, opt_decode_wide, opt_pack_streams;   /* btbb_b200_set_option */
	cudaEvent_t prof_ev[2];      /* btbb_b200_set_profiling: around the bulk scan kernel */
	int prof_on, prof_valid;
	cudaEvent_t ev_done;         /* active lane: the pending scan's last enqueued operation */
	bt_lane parked;              /* the inactive lane */
	int lane_active, lane_head, lane_count;   /* which lane the fields above belong to; oldest pending lane; pending scans (0..2) */
	btbb_b200_hit *const *d_fan; /* multi-GPU fan-out of the next scan's ordered records (device array of list pointers), see sharded.cu */
	int fan_n;
	int last_fanned;             /* the scan just ended delivered its records through the fan-out */
	bt_shard *shard;             /* multi-GPU state (sharded.cu), NULL until btbb_b200_shard_init */
	std::mutex *host_lock;       /* serialises the host-buffer entry points, which share the scratch above */
};

int btbb_b200_set_error(int code, const char *msg);
int btbb_b200_cuda_fail(cudaError_t e, const char *where);
#define BT_CUDA_TRY(call) do { cudaError_t e__ = (call); \
	if (e__ != cudaSuccess) return btbb_b200_cuda_fail(e__, #call); } while (0)

/* tables.cu */
int bt_tables_build(btbb_b200_ctx *ctx, int max_ac_errors);
void bt_tables_free(btbb_b200_ctx *ctx);

/* find_ac.cu */
#define BT_SLAB_CAP 1024
struct bt_slab_req { int used, nw; };
int bt_ensure_slab(btbb_b200_ctx *ctx, int nslabs);
int bt_scan_launch_ex(btbb_b200_ctx *ctx, const uint8_t *d_stream, int64_t n, uint32_t lap, int k,
		      btbb_b200_hit *d_out, int64_t max_hits, unsigned long long *d_count,
		      int64_t bias, cudaStream_t st, bt_slab_req *slab, int packed = 0);
int bt_find_ac_dev_begin(btbb_b200_ctx *ctx, const uint8_t *d_stream, int packed, int64_t search_length, uint32_t lap,
			 int max_ac_errors, btbb_b200_hit *d_hits, int64_t max_hits, cudaStream_t st);
int bt_find_ac_dev_end(btbb_b200_ctx *ctx, int64_t *n_hits);
int bt_find_ac_dev_impl(btbb_b200_ctx *ctx, const uint8_t *d_stream, int packed, int64_t search_length, uint32_t lap,
			int max_ac_errors, btbb_b200_hit *d_hits, int64_t max_hits, int64_t *n_hits, cudaStream_t st);
int bt_ensure_tmp(btbb_b200_ctx *ctx, int64_t hits);
void bt_lane_reset(btbb_b200_ctx *ctx);
int bt_scan_launch(btbb_b200_ctx *ctx, const uint8_t *d_stream, int64_t n, uint32_t lap, int k,
		   btbb_b200_hit *d_out, int64_t max_hits, unsigned long long *d_count,
		   int64_t bias, cudaStream_t st);
int bt_sort_hits(btbb_b200_ctx *ctx, btbb_b200_hit *a, btbb_b200_hit *b, int64_t have,
		 int passes, int64_t key_bias, cudaStream_t st, btbb_b200_hit **result);
int bt_sort_passes(int64_t span);

/* decode.cu */
int bt_try_clocks_compact(btbb_b200_ctx *ctx, const uint8_t *d_stream, int64_t stream_length,
			  const btbb_b200_pkt_in *d_pkts, const int64_t *d_idx, const unsigned long long *d_n,
			  int64_t n_max, uint16_t *d_tc, cudaStream_t st);

/* host_pack.cpp */
extern "C" void bt_pack_range(const char *stream, int64_t first, int64_t nwords, int64_t limit, uint32_t *out);
extern "C" void bt_pack_range_streams(const char *stream, int64_t first, int64_t nwords, int64_t limit, uint32_t *out, int streams);

/* find_ac_host.cpp, decode_host.cpp: the small-call path of the classic single-packet surface */
int bt_find_first_cpu(const btbb_b200_ctx *ctx, const char *stream, int search_length, uint32_t lap,
		      int max_ac_errors, btbb_b200_hit *hit, int *found);
int bt_decode_one_cpu(const char *symbols, int length, uint32_t clkn, uint8_t uap, int whitened, uint8_t type,
		      int mode, btbb_b200_decoded *out);
int bt_header_present_cpu(const char *symbols, int length);
int bt_try_clock_cpu(const char *symbols, int length, int clock, int whitened, uint8_t *uap, uint8_t *type);

/* capi.cu */
int bt_find_first_host(btbb_b200_ctx *ctx, const char *stream, int search_length, uint32_t lap,
		       int max_ac_errors, btbb_b200_hit *hit, int *found);

#endif
