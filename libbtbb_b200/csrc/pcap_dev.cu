/*
 * pcap_dev.cu -- the capture-record formatters of pcap_out.cpp on the device.
 *
 * The batch chain leaves hit records and 372-byte decode records in HBM; a capture file needs
 * 38 + payload_length bytes per packet (pcap, pcap.c:176-209) or a 4-byte padded enhanced packet
 * block (pcapng-bt.c:176-264).  Serialising on the device means one contiguous device-to-host copy
 * of the file bytes instead of the records (372 B each, mostly zero) plus a host pass over them.
 *
 *   capture_size_kernel   record sizes + exclusive prefix inside 1024-record blocks
 *   capture_base_kernel   exclusive prefix over the block totals (one CTA), grand total
 *   capture_write_kernel  one warp per record: header bytes from shared memory, payload bytes
 *                         straight from the decode record, consecutive lanes -> consecutive bytes
 *
 * Byte-for-byte the output of btbb_b200_pcap_bredr_records / btbb_b200_pcapng_bredr_blocks.
 */
#include <cuda_runtime.h>
#include "capi_internal.h"

namespace {

constexpr int BLK = 1024;
constexpr uint32_t MAX_PAYLOAD = 400;      /* BREDR_MAX_PAYLOAD */
constexpr uint32_t BB_HEADER = 22;         /* pcap_bluetooth_bredr_bb_header without the payload */

__device__ __forceinline__ uint32_t caplen_of(const btbb_b200_decoded *dec, int64_t i)
{
	int32_t len = dec[i].payload_length;
	if (len < 0) len = 0;
	if (len > (int32_t)MAX_PAYLOAD) len = MAX_PAYLOAD;
	return (uint32_t)len;
}
__device__ __forceinline__ uint32_t record_bytes(int format, uint32_t caplen)
{
	return format ? 4u * ((36u + BB_HEADER + caplen + 3u) / 4u) : 16u + BB_HEADER + caplen;
}

__global__ void __launch_bounds__(BLK) capture_size_kernel(const btbb_b200_decoded *dec, int64_t n, int format,
							     uint32_t *off, unsigned long long *bsum)
{
	__shared__ uint32_t wsum[32];
	const int64_t i = (int64_t)blockIdx.x * BLK + threadIdx.x;
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t sz = i < n ? record_bytes(format, caplen_of(dec, i)) : 0u;
	uint32_t v = sz;
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
		if (lane >= d) v += t;
	}
	if (lane == 31) wsum[wid] = v;
	__syncthreads();
	if (wid == 0) {
		uint32_t w = wsum[lane];
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t t = __shfl_up_sync(0xffffffffu, w, d);
			if (lane >= d) w += t;
		}
		wsum[lane] = w;
	}
	__syncthreads();
	const uint32_t incl = v + (wid ? wsum[wid - 1] : 0u);
	if (i < n) off[i] = incl - sz;
	if (threadIdx.x == BLK - 1) bsum[blockIdx.x] = incl;
}

/* one CTA: bsum[b] -> exclusive prefix in place, total[0] = the grand total */
__global__ void __launch_bounds__(BLK) capture_base_kernel(unsigned long long *bsum, int64_t nb, unsigned long long *total)
{
	__shared__ unsigned long long wsum[32];
	__shared__ unsigned long long carry;
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	if (threadIdx.x == 0) carry = 0;
	__syncthreads();
	for (int64_t first = 0; first < nb; first += BLK) {
		const int64_t b = first + threadIdx.x;
		const unsigned long long x = b < nb ? bsum[b] : 0ull;
		unsigned long long v = x;
		for (int d = 1; d < 32; d <<= 1) {
			const unsigned long long t = __shfl_up_sync(0xffffffffu, v, d);
			if (lane >= d) v += t;
		}
		if (lane == 31) wsum[wid] = v;
		__syncthreads();
		if (wid == 0) {
			unsigned long long w = wsum[lane];
			for (int d = 1; d < 32; d <<= 1) {
				const unsigned long long t = __shfl_up_sync(0xffffffffu, w, d);
				if (lane >= d) w += t;
			}
			wsum[lane] = w;
		}
		__syncthreads();
		const unsigned long long incl = carry + v + (wid ? wsum[wid - 1] : 0ull);
		if (b < nb) bsum[b] = incl - x;
		__syncthreads();
		if (threadIdx.x == BLK - 1) carry = incl;
		__syncthreads();
	}
	if (threadIdx.x == 0) total[0] = carry;
}

__device__ __forceinline__ void le16(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
__device__ __forceinline__ void le32(uint8_t *p, uint32_t v) { le16(p, v); le16(p + 2, v >> 16); }

constexpr int WR_WARPS = 8;

__global__ void __launch_bounds__(WR_WARPS * 32) capture_write_kernel(const btbb_b200_hit *hits, const btbb_b200_decoded *dec,
								     const btbb_b200_pcap_meta *meta, int64_t n, int format,
								     uint32_t reflap, uint32_t refuap, const uint32_t *off,
								     const unsigned long long *bbase, uint8_t *out)
{
	__shared__ uint8_t hdr[WR_WARPS][64];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const int64_t i = (int64_t)blockIdx.x * WR_WARPS + wid;
	if (i >= n) return;
	const uint32_t caplen = caplen_of(dec, i);
	const uint32_t plen = BB_HEADER + caplen, total = record_bytes(format, caplen), pre = format ? 28u : 16u;
	uint8_t *h = hdr[wid];
	if (lane == 0) {
		const btbb_b200_pcap_meta m = meta[i];
		/* flags: pcap-common.h:63-76 */
		uint32_t flags = 0x0001u | 0x0002u;                          /* dewhitened, signal power valid */
		if (m.noisedbm < m.sigdbm) flags |= 0x0004u;
		if (reflap != BTBB_B200_LAP_ANY) flags |= 0x0010u;
		if (refuap != 0xffu) flags |= 0x0080u;                       /* UAP_ANY, btbb.h:96 */
		if (caplen) flags |= 0x0020u;
		if (format) {                                                /* enhanced packet block, pcapng.h */
			le32(h, 6); le32(h + 4, total); le32(h + 8, 0);
			le32(h + 12, (uint32_t)(m.ns >> 32)); le32(h + 16, (uint32_t)m.ns);
			le32(h + 20, plen); le32(h + 24, plen);
		} else {                                                     /* pcaprec_hdr_t, nanosecond file (pcap.c:49-68) */
			le32(h, (uint32_t)(m.ns / 1000000000ull)); le32(h + 4, (uint32_t)(m.ns % 1000000000ull));
			le32(h + 8, plen); le32(h + 12, plen);
		}
		uint8_t *b = h + pre;                                        /* pcap_bluetooth_bredr_bb_header, pcap-common.h:84-97 */
		b[0] = m.channel; b[1] = (uint8_t)m.sigdbm; b[2] = (uint8_t)m.noisedbm; b[3] = hits[i].ac_errors;
		b[4] = (uint8_t)((m.transport << 4) | m.modulation);
		b[5] = 0; b[6] = 0; b[7] = 0;                                /* corrected header / payload bits: "TODO" upstream */
		le32(b + 8, hits[i].lap);
		le32(b + 12, (reflap & 0xffffffu) | (refuap << 24));
		le32(b + 16, dec[i].header_packed);
		le16(b + 20, flags);
	}
	__syncwarp();
	uint8_t *dst = out + bbase[i / BLK] + off[i];
	const uint8_t *pay = dec[i].payload;
	const uint32_t hb = pre + BB_HEADER;
	for (uint32_t j = lane; j < total; j += 32) {
		uint8_t v = 0;
		if (j < hb) v = h[j];
		else if (j < hb + caplen) v = j - hb < 344u ? pay[j - hb] : 0;   /* the record carries 344 bytes, the rest is zero */
		else if (format && j >= total - 4u) v = (uint8_t)(total >> (8u * (j - (total - 4u))));
		dst[j] = v;
	}
}

}  // namespace

/* format 0: btbb_b200_pcap_bredr_records, 1: btbb_b200_pcapng_bredr_blocks -- all pointers but
 * `bytes` are device pointers.  *bytes receives the size of the n records; they are written only when
 * that fits in cap (d_out may be NULL to ask for the size).  Synchronises the stream. */
extern "C" int btbb_b200_capture_records_dev(btbb_b200_ctx *ctx, int format, const btbb_b200_hit *d_hits,
					     const btbb_b200_decoded *d_dec, const btbb_b200_pcap_meta *d_meta, int64_t n,
					     uint32_t reflap, uint8_t refuap, uint8_t *d_out, int64_t cap, int64_t *bytes,
					     void *cuda_stream)
{
	if (!ctx || !bytes || n < 0 || (format != 0 && format != 1) || (n > 0 && (!d_hits || !d_dec || !d_meta)) || cap < 0)
		return btbb_b200_set_error(BTBB_B200_EINVAL, "capture_records_dev: bad arguments");
	*bytes = 0;
	if (n == 0) return BTBB_B200_OK;
	BT_CUDA_TRY(cudaSetDevice(ctx->device));
	cudaStream_t st = (cudaStream_t)cuda_stream;
	std::lock_guard<std::mutex> guard(*ctx->host_lock);          /* the scratch is the context's */
	const int64_t nb = (n + BLK - 1) / BLK;
	const size_t off_bytes = ((size_t)n * 4 + 15) & ~(size_t)15, need = off_bytes + (size_t)(nb + 1) * 8;
	if (need > ctx->scratch_cap[3]) {
		if (ctx->d_scratch[3]) cudaFree(ctx->d_scratch[3]);
		ctx->d_scratch[3] = NULL; ctx->scratch_cap[3] = 0;
		BT_CUDA_TRY(cudaMalloc(&ctx->d_scratch[3], need + need / 4));
		ctx->scratch_cap[3] = need + need / 4;
	}
	uint32_t *off = static_cast<uint32_t *>(ctx->d_scratch[3]);
	unsigned long long *bsum = reinterpret_cast<unsigned long long *>(static_cast<char *>(ctx->d_scratch[3]) + off_bytes);
	capture_size_kernel<<<(unsigned)nb, BLK, 0, st>>>(d_dec, n, format, off, bsum);
	capture_base_kernel<<<1, BLK, 0, st>>>(bsum, nb, bsum + nb);
	BT_CUDA_TRY(cudaGetLastError());
	unsigned long long total = 0;
	BT_CUDA_TRY(cudaMemcpyAsync(&total, bsum + nb, 8, cudaMemcpyDeviceToHost, st));
	BT_CUDA_TRY(cudaStreamSynchronize(st));
	*bytes = (int64_t)total;
	if (!d_out || (int64_t)total > cap) return BTBB_B200_OK;
	capture_write_kernel<<<(unsigned)((n + WR_WARPS - 1) / WR_WARPS), WR_WARPS * 32, 0, st>>>(
		d_hits, d_dec, d_meta, n, format, reflap, refuap, off, bsum, d_out);
	BT_CUDA_TRY(cudaGetLastError());
	BT_CUDA_TRY(cudaStreamSynchronize(st));
	return BTBB_B200_OK;
}
