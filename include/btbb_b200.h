/*
 * btbb_b200.h -- C ABI of the B200-native Bluetooth BR/EDR packet-detection path.
 *
 * This is the drop-in boundary: plain C, pointers and
 * sizes only.  Every entry point names the reference interface it replaces
 * (paths relative to the libbtbb tree, lib/src/).  The classic btbb_* surface
 * of btbb.h (btbb_init / btbb_find_ac / btbb_packet_* / btbb_decode*) is
 * exported by the same shared object (see include/btbb.h); it is implemented
 * on top of the batch entry points declared here.
 *
 * There is NO CPU fallback behind any of these calls: they launch sm_100a
 * kernels and return a negative BTBB_B200_E* code when CUDA is unavailable.
 */
#ifndef BTBB_B200_H
#define BTBB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BTBB_B200_LAP_ANY 0xffffffffu   /* btbb.h:95 LAP_ANY */

/* error codes (all negative; the reference uses -1 for every failure, btbb.h:69-73) */
#define BTBB_B200_OK            0
#define BTBB_B200_EINVAL       -1   /* bad argument (e.g. max_ac_errors outside 0..5, bluetooth_packet.c:282) */
#define BTBB_B200_ECUDA        -2   /* CUDA runtime / driver error, no device, wrong arch */
#define BTBB_B200_ENOMEM       -3
#define BTBB_B200_EOVERFLOW    -4   /* more hits than the caller's buffer holds (count is still returned) */

/* One detected access code.  16 bytes, SURVEY.md 8d "16 B written per hit". */
typedef struct btbb_b200_hit {
	int64_t  offset;      /* symbol index of the first sync-word symbol   (btbb_find_ac return value) */
	uint32_t lap;         /* recovered / requested LAP                     (init_packet, bluetooth_packet.c:201) */
	uint8_t  ac_errors;   /* corrected bit errors in the access code       (btbb_packet_get_ac_errors) */
	uint8_t  pad[3];      /* zero */
} btbb_b200_hit;

/* Result of the per-packet chain for one packet (btbb_decode_header + btbb_decode_payload,
 * bluetooth_packet.c:1198-1297), or of try_clock + crc_check (:1178-1195, :708-769). */
typedef struct btbb_b200_decoded {
	int32_t  header_ok;             /* btbb_decode_header() return; try_clock mode: unfec13 success */
	int32_t  rv;                    /* btbb_decode_payload()/crc_check() return: 0,1,2,10,1000 */
	uint8_t  uap;                   /* pkt->UAP */
	uint8_t  type;                  /* pkt->packet_type */
	uint8_t  lt_addr;               /* pkt->packet_lt_addr */
	uint8_t  flags;                 /* pkt->packet_flags */
	uint8_t  hec;                   /* pkt->packet_hec */
	uint8_t  llid;                  /* pkt->payload_llid */
	uint8_t  flow;                  /* pkt->payload_flow */
	uint8_t  has_payload;           /* BTBB_HAS_PAYLOAD flag */
	int32_t  payload_header_length; /* pkt->payload_header_length */
	int32_t  payload_length;        /* pkt->payload_length (bytes) */
	uint32_t header_packed;         /* btbb_packet_get_header_packed() */
	uint8_t  payload[344];          /* btbb_get_payload_packed() when rv >= 2, else zero */
} btbb_b200_decoded;

/* Per-packet input of the decode chain (what btbb_packet_set_data + set_uap + flags carry). */
typedef struct btbb_b200_pkt_in {
	int64_t  offset;     /* symbol index of the sync word's first symbol in the stream */
	int32_t  length;     /* symbols available from offset (clamped to 3125, bluetooth_packet.c:472) */
	uint32_t clkn;       /* CLK1-27 as stored in pkt->clkn (i.e. already >>1); only the low 6 bits whiten */
	uint8_t  uap;        /* expected UAP for btbb_decode_header (:1211) */
	uint8_t  whitened;   /* BTBB_WHITENED flag (:663) */
	uint8_t  type;       /* pkt->packet_type as set by the caller; used only by modes >= 2 */
	uint8_t  pad;
	uint32_t reserved;   /* keeps the record 24 bytes without implicit padding */
} btbb_b200_pkt_in;

/* decode modes */
#define BTBB_B200_MODE_DECODE      0   /* btbb_decode_header + btbb_decode_payload (:1198-1297) */
#define BTBB_B200_MODE_TRY_CLOCKS  1   /* try_clock + crc_check for CLK1-6 = 0..63 (:1178-1195, :708-769) */
#define BTBB_B200_MODE_PAYLOAD     2   /* btbb_decode_payload alone, packet_type/UAP/clkn as given (:1223-1297) */
#define BTBB_B200_MODE_CRC_CHECK   3   /* crc_check(clkn & 63, pkt) alone, packet_type/UAP as given (:708-769) */
#define BTBB_B200_MODE_RAW         16  /* + n: one type decoder without crc_check's post-filter:
                                          0 fhs, 1 DM, 2 DH, 3 EV3, 4 EV4, 5 EV5, 6 HV (:783-1174) */
/* OR-ed into a mode: payload[] carries what the reference's decoders leave in pkt->payload even
 * when the decode FAILED (rv < 2) -- the bytes btbb_pcap_append_packet logs (pcap.c:173-209) --
 * instead of zeros.  Bits no decoder wrote read 0 (a freshly allocated btbb_packet). */
#define BTBB_B200_MODE_FLAG_RAW_PAYLOAD 0x100

typedef struct btbb_b200_ctx btbb_b200_ctx;

/* Library / device bring-up.  Replaces btbb_init() (bluetooth_packet.c:279-292): builds the
 * syndrome -> error tables for 1..max_ac_errors bit errors on `device` and uploads the
 * constant tables.  max_ac_errors outside [0,5] -> BTBB_B200_EINVAL.  */
int  btbb_b200_create(int device, int max_ac_errors, btbb_b200_ctx **ctx);
void btbb_b200_destroy(btbb_b200_ctx *ctx);
int  btbb_b200_device(const btbb_b200_ctx *ctx);
int  btbb_b200_table_errors(const btbb_b200_ctx *ctx);   /* the k the tables were built for */
const char *btbb_b200_last_error(void);

/*
 * Batch btbb_find_ac (bluetooth_packet.c:444-464) over a DEVICE-resident stream of one byte
 * per symbol (values 0/1), reporting EVERY position in [0, search_length) that the
 * reference would report when iterated with restart at offset+1, in ascending offset order.
 *   lap == BTBB_B200_LAP_ANY -> promiscuous_packet_search semantics (:368-420)
 *   otherwise                -> find_known_lap semantics (:423-441)
 * d_stream must hold search_length + 63 readable symbols (the header asks for +72, btbb.h:82).
 * d_hits has room for max_hits records; *n_hits receives the total found (may exceed max_hits,
 * in which case BTBB_B200_EOVERFLOW is returned and d_hits holds max_hits genuine records in
 * ascending order -- not necessarily the max_hits lowest offsets).
 * All work is enqueued on cuda_stream (a cudaStream_t passed as void*; NULL = default stream);
 * the call synchronises that stream before returning.
 */
int btbb_b200_find_ac_dev(btbb_b200_ctx *ctx, const uint8_t *d_stream, int64_t search_length,
			  uint32_t lap, int max_ac_errors,
			  btbb_b200_hit *d_hits, int64_t max_hits, int64_t *n_hits,
			  void *cuda_stream);

/*
 * btbb_b200_find_ac_dev in two halves, for callers that pipeline: _begin enqueues the scan and
 * the ordering pass on cuda_stream and returns without waiting; _end waits for them and
 * delivers the count and the status (in the uncommon cases -- known-LAP scans, very dense
 * hits -- it still has work to enqueue and wait for).  TWO calls can be pending per context, each
 * with its own d_hits (untouched in between); _end completes the OLDEST one.  A caller that begins
 * scan i + 1 before it ends scan i keeps the GPU busy across its own host work and the library's
 * launch latency.  btbb_b200_find_ac_dev is _begin followed by _end.
 */
int btbb_b200_find_ac_dev_begin(btbb_b200_ctx *ctx, const uint8_t *d_stream, int64_t search_length,
				uint32_t lap, int max_ac_errors,
				btbb_b200_hit *d_hits, int64_t max_hits, void *cuda_stream);
int btbb_b200_find_ac_dev_end(btbb_b200_ctx *ctx, int64_t *n_hits);

/* Added to every offset btbb_b200_find_ac_dev / _packed_dev / _dev_begin report from now on (0
 * after btbb_b200_create): a rank that scans one shard of a longer stream gets global offsets
 * straight from the kernels. */
int btbb_b200_set_offset_bias(btbb_b200_ctx *ctx, int64_t bias);

/* Tunables of one context (defaults in brackets); nothing here changes a result. */
#define BTBB_B200_OPT_TILE_KERNEL_ONLY    1   /* [0] scans skip the bulk kernels and run the tile kernels (A/B check) */
#define BTBB_B200_OPT_HOST_BYTE_ROUTE     2   /* [0] btbb_b200_find_ac_host copies the byte stream instead of packing on the host */
#define BTBB_B200_OPT_HOST_SPLIT_PERMILLE 3   /* [0] share of a large host call that travels as bytes while the rest is packed */
#define BTBB_B200_OPT_PACK_THREADS        4   /* [0 = all usable CPUs, at most 32] host pack threads */
#define BTBB_B200_OPT_TRACE               5   /* [0] timing lines of the host-buffer scan on stderr */
#define BTBB_B200_OPT_DECODE_WIDE_STAGING 6   /* [0] 64-clock sweep: 12 warps per SM staging 32 records instead of 24 x 16 */
#define BTBB_B200_OPT_PACK_STREAMS        7   /* [4] address streams a host pack thread advances in lock step (1, 2, 4 or 8) */
int btbb_b200_set_option(btbb_b200_ctx *ctx, int option, int64_t value);

/* Measurement hook (bench.py's roofline): with profiling on, every scan records CUDA events right
 * before and after its bulk kernel on the scan's own stream; after the scan has been waited for,
 * btbb_b200_last_scan_kernel_ms returns that kernel's duration.  Off by default. */
int btbb_b200_set_profiling(btbb_b200_ctx *ctx, int on);
int btbb_b200_last_scan_kernel_ms(btbb_b200_ctx *ctx, float *ms);

/*
 * btbb_b200_find_ac_dev for a stream that is already PACKED, 32 symbols per word: symbol i of
 * the stream is bit (i & 31) of d_words[i >> 5] -- the reference's own bit order when it packs
 * a window (air_to_host64, bluetooth_packet.c:235-242: symbol i <-> bit i).  This is the
 * packed-symbol ingest of SURVEY.md 8(f) row 2: a producer that keeps demodulated bits packed
 * skips the 8x byte-per-symbol inflation.  d_words must be 4-byte aligned and hold
 * ceil((search_length + 63) / 32) words.  Same results, ordering and error behaviour as
 * btbb_b200_find_ac_dev.  (btbb_b200_find_ac_host uses this path internally for large
 * buffers: it packs on the host cores while copying, because PCIe is the bound there.)
 */
int btbb_b200_find_ac_packed_dev(btbb_b200_ctx *ctx, const uint32_t *d_words, int64_t search_length,
				 uint32_t lap, int max_ac_errors,
				 btbb_b200_hit *d_hits, int64_t max_hits, int64_t *n_hits,
				 void *cuda_stream);

/* Same, but only enqueues the scan kernel(s) (no synchronisation, no ordering pass);
 * used by the benchmark to time the kernel alone.  d_count is a device int64 counter
 * of the hits (unordered), zeroed by the call itself on the same stream. */
int btbb_b200_find_ac_enqueue(btbb_b200_ctx *ctx, const uint8_t *d_stream, int64_t search_length,
			      uint32_t lap, int max_ac_errors,
			      btbb_b200_hit *d_hits, int64_t max_hits, unsigned long long *d_count,
			      void *cuda_stream);

/* Host-buffer form: copies the stream to the device in chunks (double-buffered, pinned
 * staging), scans, and returns sorted hits in host memory.  This is what the classic
 * btbb_find_ac() shim calls. */
int btbb_b200_find_ac_host(btbb_b200_ctx *ctx, const char *stream, int64_t search_length,
			   uint32_t lap, int max_ac_errors,
			   btbb_b200_hit *hits, int64_t max_hits, int64_t *n_hits);

/*
 * Batch per-packet chain.  BTBB_B200_MODE_DECODE uses the packet's own clkn/UAP; d_out
 * holds n records.  BTBB_B200_MODE_TRY_CLOCKS is the inner loop of
 * bluetooth_piconet.c:675-689; d_out holds n*64 records, record [p*64+c] for packet p,
 * clock c.  The other modes serve the single-function entry points of the classic API.  Symbols at index >= length read as 0 (a freshly calloc'ed
 * btbb_packet, bluetooth_packet.c:297).
 */
int btbb_b200_decode_dev(btbb_b200_ctx *ctx, const uint8_t *d_stream, int64_t stream_length,
			 const btbb_b200_pkt_in *d_pkts, int64_t n, int mode,
			 btbb_b200_decoded *d_out, void *cuda_stream);

int btbb_b200_decode_host(btbb_b200_ctx *ctx, const char *stream, int64_t stream_length,
			  const btbb_b200_pkt_in *pkts, int64_t n, int mode,
			  btbb_b200_decoded *out);

/* BTBB_B200_MODE_TRY_CLOCKS with a compact result: one 16-bit word per (packet, clock) at
 * d_tc[64 * packet + clock] -- the UAP try_clock derived (bluetooth_packet.c:1178-1195) in the low
 * byte, the class of crc_check's return value (:708-769) above it: 0, 1, 2 as returned, 3 for 10,
 * 4 for 1000.  This is what the UAP sieve consumes; 128 bytes per packet instead of 23 808. */
int btbb_b200_try_clocks_compact_dev(btbb_b200_ctx *ctx, const uint8_t *d_stream, int64_t stream_length,
				     const btbb_b200_pkt_in *d_pkts, int64_t n, uint16_t *d_tc, void *cuda_stream);

/*
 * One packet through the same chain ON THE HOST (decode_core.h compiled for the CPU): the small-call
 * path of the classic single-packet surface (btbb_decode_header / btbb_decode_payload / try_clock /
 * crc_check ..., bluetooth_packet.c:1178-1317), where a caller hands over at most 3125 symbols per
 * call and a kernel launch would cost far more than the arithmetic ("Drop-in
 * latency").  Same records as btbb_b200_decode_host for one packet: out holds 1 record, 64 in
 * BTBB_B200_MODE_TRY_CLOCKS.  `type` is used by modes >= 2 only.  The batch entry points above
 * never take this route.
 */
int btbb_b200_decode_smallcall(const char *symbols, int length, uint32_t clkn, uint8_t uap,
			       int whitened, uint8_t type, int mode, btbb_b200_decoded *out);

/* btbb_header_present (bluetooth_packet.c:1371-1408) for n packets; d_present[n] gets 0/1. */
int btbb_b200_header_present_dev(btbb_b200_ctx *ctx, const uint8_t *d_stream, int64_t stream_length,
				 const btbb_b200_pkt_in *d_pkts, int64_t n,
				 uint8_t *d_present, void *cuda_stream);

int btbb_b200_header_present_host(btbb_b200_ctx *ctx, const char *stream, int64_t stream_length,
				  const btbb_b200_pkt_in *pkts, int64_t n, uint8_t *present);

/* Routing of the CLASSIC surface (include/btbb.h) only: btbb_find_ac calls that search at most
 * find_ac_host_below positions (default 8192; -1 = never) and, unless packet_calls_on_gpu is set,
 * the single-packet calls (btbb_decode_header / _payload, try_clock, crc_check, fhs / DM / ...,
 * btbb_header_present) are answered by the host small-call path; everything else launches
 * kernels.  The environment variable BTBB_B200_CLASSIC=gpu sets (-1, 1) at first use.  The batch
 * entry points of this header are not affected: they always run on the GPU. */
void btbb_b200_classic_config(int find_ac_host_below, int packet_calls_on_gpu);

/*
 * UAP / CLK1-6 discovery from packet headers: btbb_uap_from_header (bluetooth_piconet.c:648-750)
 * as btbb_process_packet drives it in survey mode (:851-858), for many piconets at once
 * (SURVEY.md 8(f) row 1).  btbb_b200_sieve is the part of struct btbb_piconet
 * (bluetooth_piconet.h:30-105) those two functions read and write; flag bit numbers are the
 * reference's (btbb.h:27-42).  A zeroed record is a freshly calloc'ed piconet.
 */
typedef struct btbb_b200_sieve {
	uint32_t flags;                  /* BTBB_UAP_VALID 2, BTBB_CLK6_VALID 4, BTBB_CLK27_VALID 5, BTBB_HOP_REVERSAL_INIT 9,
	                                    BTBB_GOT_FIRST_PACKET 10, BTBB_IS_AFH 11, BTBB_LOOKS_LIKE_AFH 12 */
	uint32_t first_pkt_time;         /* pn->first_pkt_time */
	int32_t  clk_offset;             /* pn->clk_offset */
	int32_t  packets_observed;       /* pn->packets_observed */
	int32_t  total_packets_observed; /* pn->total_packets_observed */
	uint8_t  uap;                    /* pn->UAP */
	uint8_t  used_channels;          /* pn->used_channels */
	uint8_t  afh_map[10];            /* pn->afh_map (btbb_piconet_set_channel_seen) */
	int16_t  clock6_candidates[64];  /* pn->clock6_candidates: -1 eliminated, else the UAP that clock implies */
} btbb_b200_sieve;

#define BTBB_B200_SIEVE_NOT_CALLED (-2)  /* rv: header absent or UAP already known -> btbb_uap_from_header not called */

/*
 * For every group g (one piconet, i.e. one LAP), packets [group_start[g], group_start[g+1]) of
 * d_pkts are its packets in arrival order.  Each is handled as btbb_process_packet does in survey
 * mode: btbb_piconet_set_channel_seen(channel), then, if btbb_header_present(pkt) and the UAP is
 * not known yet, btbb_uap_from_header(pkt, pn).  The channel is the low byte of
 * btbb_b200_pkt_in.reserved, clkn is the packet's CLKN as btbb_packet_set_data stores it.
 * d_states[g] is read, updated and written back, so a capture can be fed in several calls.
 * d_rv[p] receives btbb_uap_from_header's return value (0 / 1) or BTBB_B200_SIEVE_NOT_CALLED;
 * it may be NULL.  The 64 try_clock / crc_check evaluations per packet run on the GPU in one
 * pass over all packets, the (sequential per piconet) candidate elimination one warp per piconet.
 * The hop-pattern log (pattern_indices / pattern_channels) is not kept: hop reversal is out
 * of scope; the 1000-packet limit that resets a piconet (:665-671) is honoured.
 */
int btbb_b200_uap_sieve_dev(btbb_b200_ctx *ctx, const uint8_t *d_stream, int64_t stream_length,
			    const btbb_b200_pkt_in *d_pkts, int64_t n_pkts,
			    const int64_t *d_group_start, int64_t n_groups,
			    btbb_b200_sieve *d_states, int8_t *d_rv, void *cuda_stream);
/* Host helper: stable grouping of hit records by LAP (what get_piconet(LAP),
 * bluetooth_piconet.c:820-840, does packet by packet in survey mode).  order[0..n) receives the
 * hit indices sorted by (LAP, arrival order); group_start (n + 1 entries, or NULL) the group
 * boundaries in that order; laps (n entries, or NULL) each group's LAP.  Returns the number of
 * groups, or -1. */
int64_t btbb_b200_group_by_lap(const btbb_b200_hit *hits, int64_t n, int64_t *order,
			       int64_t *group_start, uint32_t *laps);
int btbb_b200_uap_sieve_host(btbb_b200_ctx *ctx, const char *stream, int64_t stream_length,
			     const btbb_b200_pkt_in *pkts, int64_t n_pkts,
			     const int64_t *group_start, int64_t n_groups,
			     btbb_b200_sieve *states, int8_t *rv);

/*
 * ---- multi-GPU: one process per GPU, contiguous shards, all-gather of hit records ----
 * Rank r of `world` scans positions [first_position, first_position + search_length) of the global
 * stream from a device buffer holding that range plus 63 more symbols (the north star fixes the
 * seam overlap at 72); the kernels report GLOBAL offsets.  Ranges partition the positions, so the
 * rank-order concatenation of the per-rank sorted lists is the sorted list of the whole stream.
 * The records travel over NVLink peer memory: every rank's gather buffer is mapped into every other
 * rank (CUDA IPC) and the kernel that orders a scan's hits stores each record into this rank's slot on
 * all GPUs as it writes the local list -- the all-gather is fused into the ordering pass (or,
 * BTBB_B200_SHARD_COPY_ENGINES, pushed there by the copy engines afterwards).  With BTBB_B200_SHARD_NCCL_ONLY, or where peer mapping fails, the exchange is an
 * NCCL allgatherv (one all-gather of the counts + one group of exact-size broadcasts).  NCCL (libnccl.so.2) is loaded on first use.
 *
 *   rank 0:      btbb_b200_shard_unique_id(id); hand the 128 bytes to every rank (any transport)
 *   every rank:  btbb_b200_shard_init(ctx, id, rank, world, slot_records, flags)
 *   per scan:    btbb_b200_find_ac_sharded_begin(...)   enqueue the scan of this rank's shard
 *                btbb_b200_find_ac_sharded_end(...)     wait for it, start pushing its records
 *                btbb_b200_find_ac_sharded_next(...)    = _end of this scan + _begin of the next one,
 *                                                       the push overlapping that next scan
 *   finally:     btbb_b200_find_ac_sharded_gather(...)  every rank's records of the latest scan
 * btbb_b200_find_ac_sharded_dev is begin + end + gather + concatenation into one buffer.
 * slot_records bounds the hits ONE rank may report per scan (BTBB_B200_EOVERFLOW beyond).
 */
#define BTBB_B200_SHARD_ID_BYTES 128
#define BTBB_B200_SHARD_NCCL_ONLY 1       /* exchange = NCCL allgatherv after each scan */
#define BTBB_B200_SHARD_COPY_ENGINES 2    /* peer memory, pushed by the copy engines while the next scan runs, instead of stored by the ordering kernel */
int btbb_b200_shard_unique_id(void *id);
int btbb_b200_shard_init(btbb_b200_ctx *ctx, const void *id, int rank, int world, int64_t slot_records, int flags);
int btbb_b200_shard_info(const btbb_b200_ctx *ctx, int *rank, int *world, int *peer_memory);
int btbb_b200_shard_destroy(btbb_b200_ctx *ctx);
int btbb_b200_find_ac_sharded_begin(btbb_b200_ctx *ctx, const uint8_t *d_stream, int64_t search_length,
				    int64_t first_position, uint32_t lap, int max_ac_errors, void *cuda_stream);
int btbb_b200_find_ac_sharded_end(btbb_b200_ctx *ctx, int64_t *n_local);
/* _end of the pending scan + _begin of the next, ordered so that the GPU never idles: wait for
 * scan i, enqueue scan i + 1, then push scan i's records underneath it; *n_prev = scan i's hits */
int btbb_b200_find_ac_sharded_next(btbb_b200_ctx *ctx, const uint8_t *d_stream, int64_t search_length,
				   int64_t first_position, uint32_t lap, int max_ac_errors, void *cuda_stream,
				   int64_t *n_prev);
/* rank r's records of the latest scan: (*d_slots) + r * (*slot_stride), counts[r] of them (device memory
 * owned by the library, valid until the next but one _end); counts holds `world` entries */
int btbb_b200_find_ac_sharded_gather(btbb_b200_ctx *ctx, const btbb_b200_hit **d_slots, int64_t *slot_stride,
				     int64_t *counts, int64_t *n_total);
int btbb_b200_find_ac_sharded_dev(btbb_b200_ctx *ctx, const uint8_t *d_stream, int64_t search_length,
				  int64_t first_position, uint32_t lap, int max_ac_errors,
				  btbb_b200_hit *d_all, int64_t max_all, int64_t *counts, int64_t *n_total, void *cuda_stream);

/*
 * ---- hop sequence and CLK1-27 discovery (bluetooth_piconet.c:311-362, :455-499, :575-645) ----
 */
typedef struct btbb_b200_hop_cfg {
	uint32_t address;       /* (UAP << 24 | LAP) & 0xfffffff, as gen_hop_pattern passes it (:371-372) */
	uint8_t  afh;           /* BTBB_IS_AFH: only the channels of afh_map are in the register bank (precalc :171-193) */
	uint8_t  aliased;       /* winnow only: observed channels are aliased into 26..50 (aliased_channel :449-452) */
	uint8_t  pad[2];
	uint8_t  afh_map[10];   /* bit c % 8 of byte c / 8 = channel c in use (btbb_piconet_set_afh_map) */
	uint8_t  pad2[2];
} btbb_b200_hop_cfg;

/* gen_hops (:311-362): entries [first, first + n) of the 2^27-entry channel sequence (index = CLK1-27),
 * written to device memory.  Any range on its own; 2^27 entries are the reference's whole table. */
int btbb_b200_hop_sequence_dev(btbb_b200_ctx *ctx, const btbb_b200_hop_cfg *cfg, int64_t first, int64_t n,
			       uint8_t *d_out, void *cuda_stream);
/*
 * Hop reversal: btbb_init_hop_reversal's candidate list (init_candidates :455-472: every CLK1-27 value
 * whose low six bits are known_clk6) filtered by n_obs observed hops as channel_winnow does
 * (:575-611): candidate c survives observation j iff sequence[(c + indices[j]) mod 2^27] (aliased if
 * cfg->aliased) == channels[j].  No 128 MiB table is built; the hop selection is evaluated per
 * candidate.  candidates[] receives the survivors of ALL observations in ascending order (the order of
 * the reference's list), *n_candidates their number (BTBB_B200_EOVERFLOW if more than
 * max_candidates); survivors_after[j] (n_obs entries, may be NULL) = candidates left after
 * observations 0..j -- pn->num_candidates after the reference's j-th channel_winnow call, which lets
 * a caller reproduce btbb_winnow's stop at <= 1 (:622-623).  Host arrays in, host arrays out.
 */
int btbb_b200_hop_winnow(btbb_b200_ctx *ctx, const btbb_b200_hop_cfg *cfg, uint32_t known_clk6, int n_obs,
			 const int32_t *indices, const uint8_t *channels, uint32_t *candidates, int64_t max_candidates,
			 int64_t *n_candidates, int32_t *survivors_after);

/*
 * BR/EDR capture records (pcap, DLT 255) from batch results: the bytes btbb_pcap_create_file /
 * btbb_pcap_append_packet (pcap.c:74-100, 176-209; record layout pcap-common.h:84-97) write,
 * serialised from btbb_b200_hit + btbb_b200_decoded instead of from a btbb_packet.  meta carries
 * what the reference takes from its arguments and from btbb_packet_set_data / _set_transport /
 * _set_modulation.  Host formatting only.  The reference logs pkt->payload even where the payload
 * decode failed (rv < 2): records produced with BTBB_B200_MODE_FLAG_RAW_PAYLOAD carry those bytes and
 * give byte-identical files for EVERY packet; without the flag such packets get zero payload bytes.
 */
typedef struct btbb_b200_pcap_meta {
	uint64_t ns;          /* timestamp, nanoseconds */
	int8_t   sigdbm;      /* signal power */
	int8_t   noisedbm;    /* noise power */
	uint8_t  channel;     /* rf channel (btbb_packet_get_channel) */
	uint8_t  transport;   /* BTBB_TRANSPORT_* */
	uint8_t  modulation;  /* BTBB_MOD_* */
	uint8_t  pad[3];
} btbb_b200_pcap_meta;

/* 24-byte file header; returns the bytes written or -1 */
int64_t btbb_b200_pcap_file_header(uint8_t *out, int64_t cap);
/* records for n packets; returns the bytes they take (written only when they fit in cap; out may be NULL) */
int64_t btbb_b200_pcap_bredr_records(const btbb_b200_hit *hits, const btbb_b200_decoded *dec,
				     const btbb_b200_pcap_meta *meta, int64_t n,
				     uint32_t reflap, uint8_t refuap, uint8_t *out, int64_t cap);

/* the same packets as pcapng enhanced packet blocks (btbb_pcapng_append_packet, pcapng-bt.c:176-264);
 * pad bytes are zero (upstream leaves uninitialised stack bytes there).  Section header and interface
 * description blocks are the capture program's to write. */
int64_t btbb_b200_pcapng_bredr_blocks(const btbb_b200_hit *hits, const btbb_b200_decoded *dec,
				      const btbb_b200_pcap_meta *meta, int64_t n,
				      uint32_t reflap, uint8_t refuap, uint8_t *out, int64_t cap);

/* ---- host small-call helpers (no device needed) ---- */
/* The short-search path of the classic btbb_find_ac (searches of at most 8192 positions are answered on
 * the host, see btbb_b200_classic_config) as a call of its own: FIRST hit of the search the reference
 * runs after btbb_init(table_errors) -- known LAP or BTBB_B200_LAP_ANY -- *found = 0 if there is none.
 * table_errors 0..4. */
int btbb_b200_find_first_smallcall(const char *stream, int search_length, uint32_t lap, int table_errors,
				   int max_ac_errors, btbb_b200_hit *hit, int *found);

/* The two formatters above on the device (pcap_dev.cu): hits, dec, meta and out are DEVICE pointers, so
 * the batch chain's records are serialised where they lie and only the file bytes (38 + payload bytes per
 * packet instead of a 16-byte hit + 372-byte record) cross PCIe.  format 0 = pcap records, 1 = pcapng
 * enhanced packet blocks; byte-identical to the host formatters.  *bytes (host) receives the size of the
 * n records; they are written only when that fits in cap (d_out may be NULL to ask for the size).
 * Synchronises the stream. */
int btbb_b200_capture_records_dev(btbb_b200_ctx *ctx, int format, const btbb_b200_hit *d_hits,
				  const btbb_b200_decoded *d_dec, const btbb_b200_pcap_meta *d_meta, int64_t n,
				  uint32_t reflap, uint8_t refuap, uint8_t *d_out, int64_t cap, int64_t *bytes,
				  void *cuda_stream);

/* ---- synthetic capture generator (SURVEY.md 8d "Synthetic input"; test/bench data only) ---- */
typedef struct btbb_b200_synth_cfg {
	uint64_t seed;          /* 0xB200B7BB by default */
	int64_t  n_symbols;     /* total symbols to generate (the buffer must hold this many bytes) */
	int64_t  first_symbol;  /* global index of buffer[0] (lets each rank generate its own shard) */
	int32_t  stride;        /* one planted packet per `stride` symbols; 0 = noise only */
	int32_t  n_laps;        /* planted LAPs are drawn from a table of this many (1..64) */
	uint32_t ber_q32;       /* bit-flip probability * 2^32 applied to every symbol */
	uint32_t packet_mix;    /* bit i set => packet kind i may be planted (see BTBB_B200_KIND_*) */
	uint32_t fixed_lap;     /* if n_laps == 1: the LAP to plant */
	uint32_t reserved;      /* bit 0: piconet-coherent capture (one UAP per LAP, CLK1-6 = slot + a per-LAP offset) */
} btbb_b200_synth_cfg;

#define BTBB_B200_KIND_ID    0   /* 68-symbol ID packet: access code only */
#define BTBB_B200_KIND_DM1   1
#define BTBB_B200_KIND_DH1   2
#define BTBB_B200_KIND_DM3   3
#define BTBB_B200_KIND_FHS   4
#define BTBB_B200_KIND_HV1   5
#define BTBB_B200_KIND_DM5   6
#define BTBB_B200_KIND_DH3   7
#define BTBB_B200_KIND_COUNT 8

/* Ground truth for planted packet `slot` (same on host and device). */
typedef struct btbb_b200_planted {
	int64_t  offset;     /* symbol index of the sync word */
	uint32_t lap;
	uint8_t  uap;
	uint8_t  kind;
	uint8_t  clk6;       /* whitening clock CLK1-6 */
	uint8_t  lt_addr;
	int32_t  n_symbols;  /* packet length in symbols from the sync word (68 for ID) */
	int32_t  body_bytes; /* payload body length */
} btbb_b200_planted;

int btbb_b200_synth_host(const btbb_b200_synth_cfg *cfg, uint8_t *buffer);
int btbb_b200_synth_dev(const btbb_b200_synth_cfg *cfg, uint8_t *d_buffer, void *cuda_stream);
int btbb_b200_synth_planted(const btbb_b200_synth_cfg *cfg, int64_t slot, btbb_b200_planted *out);

#ifdef __cplusplus
}
#endif
#endif /* BTBB_B200_H */
