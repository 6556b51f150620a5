/*
 * host_pack.cpp -- the host half of the packed-symbol transfer format (capi.cu,
 * scan_host_packed): 32 one-byte symbols -> one word, symbol i -> bit i, the reference's own
 * window bit order (air_to_host64, bluetooth_packet.c:235-242).  Plain C++ so that the host
 * compiler sees the intrinsics headers directly; AVX2 when the CPU has it, SSE2 otherwise
 * (always present on x86-64), scalar elsewhere.
 */
#include <stdint.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace {

#if !defined(__x86_64__)
void pack_scalar(const char *p, int64_t nwords, uint32_t *out)
{
	for (int64_t w = 0; w < nwords; w++) {
		uint32_t v = 0;
		for (int j = 0; j < 32; j++) v |= (uint32_t)(p[32 * w + j] & 1) << j;
		out[w] = v;
	}
}
#endif

#if defined(__x86_64__)
void pack_sse2(const char *p, int64_t nwords, uint32_t *out)
{
	for (int64_t w = 0; w < nwords; w++) {
		const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(p + 32 * w));
		const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(p + 32 * w + 16));
		out[w] = (uint32_t)_mm_movemask_epi8(_mm_slli_epi16(a, 7)) |
			 ((uint32_t)_mm_movemask_epi8(_mm_slli_epi16(b, 7)) << 16);
	}
}

__attribute__((target("avx2"))) void pack_avx2(const char *p, int64_t nwords, uint32_t *out)
{
	int64_t w = 0;
	for (; w + 4 <= nwords; w += 4) {
		_mm_prefetch(p + 32 * w + 2048, _MM_HINT_T0);
		const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(p + 32 * w));
		const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(p + 32 * w + 32));
		const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(p + 32 * w + 64));
		const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(p + 32 * w + 96));
		out[w] = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(a, 7));
		out[w + 1] = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(b, 7));
		out[w + 2] = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(c, 7));
		out[w + 3] = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(d, 7));
	}
	for (; w < nwords; w++) {
		const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(p + 32 * w));
		out[w] = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(a, 7));
	}
}

/* The same words, S sub-ranges of the block advanced in lock step: one core streaming a single
 * address range is limited by how many cache-line fills it keeps in flight (the hardware prefetcher
 * follows one stream per 4 KiB page); several concurrent streams raise that, +25..35 % per core where
 * measured.  q words per sub-range, q even. */
template <int S>
__attribute__((target("avx2"))) void pack_avx2_streams(const char *p, int64_t q, uint32_t *out)
{
	for (int64_t w = 0; w + 2 <= q; w += 2)
		for (int s = 0; s < S; s++) {
			const char *src = p + 32 * (s * q + w);
			const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src));
			const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + 32));
			out[s * q + w] = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(a, 7));
			out[s * q + w + 1] = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(b, 7));
		}
}
#endif

}  // namespace

/* symbols [first, first + 32 * nwords) of stream -> out[0 .. nwords); nothing at or past
 * `limit` is read (the last word of a stream is zero-filled) */
extern "C" void bt_pack_range_streams(const char *stream, int64_t first, int64_t nwords, int64_t limit, uint32_t *out, int streams)
{
	int64_t full = (limit - first) / 32;
	if (full > nwords) full = nwords;
	if (full < 0) full = 0;
#if defined(__x86_64__)
	static const int have_avx2 = __builtin_cpu_supports("avx2");
	int64_t done = 0;
	if (have_avx2 && streams > 1 && full >= 4096) {
		const int S = streams >= 8 ? 8 : streams >= 4 ? 4 : 2;
		const int64_t q = (full / S) & ~(int64_t)1;
		if (S == 8) pack_avx2_streams<8>(stream + first, q, out);
		else if (S == 4) pack_avx2_streams<4>(stream + first, q, out);
		else pack_avx2_streams<2>(stream + first, q, out);
		done = q * S;
	}
	if (have_avx2) pack_avx2(stream + first + 32 * done, full - done, out + done);
	else pack_sse2(stream + first, full, out);
#else
	pack_scalar(stream + first, full, out);
#endif
	for (int64_t w = full; w < nwords; w++) {
		uint32_t v = 0;
		for (int j = 0; j < 32 && first + 32 * w + j < limit; j++) v |= (uint32_t)(stream[first + 32 * w + j] & 1) << j;
		out[w] = v;
	}
}

extern "C" void bt_pack_range(const char *stream, int64_t first, int64_t nwords, int64_t limit, uint32_t *out)
{
	bt_pack_range_streams(stream, first, nwords, limit, out, 1);
}
