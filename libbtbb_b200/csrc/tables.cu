/*
 * tables.cu -- host-side construction of the device tables used by the access-code scan:
 *   * byte LUTs of the (64,30) code's syndrome (what sw_check_tables.h holds in the
 *     reference, here derived from g(x) and cut to the low 32 bits for the fast path),
 *   * the syndrome -> error-pattern map of gen_syndrome_map()/cycle()
 *     (bluetooth_packet.c:161-185) as an open-addressing table in global memory,
 *   * a two-hash Bloom bitmap over the map's keys that the kernel keeps in shared memory.
 */
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "bt_math.h"
#include "capi_internal.h"
#include "scan_hash.h"

static uint64_t g_bit_syn[64];

static void enumerate(std::vector<bt_err_slot> &out, uint64_t err, uint64_t syn, int start, int depth)
{
	for (int i = start; i < 58; i++) {      /* errors are confined to bits 0..57 (:167) */
		uint64_t e = err | (1ULL << i), s = syn ^ g_bit_syn[i];
		if (depth > 1)
			enumerate(out, e, s, i + 1, depth - 1);
		else
			out.push_back(bt_err_slot{s, e});
	}
}

int bt_tables_build(btbb_b200_ctx *ctx, int k)
{
	for (int i = 0; i < 64; i++)
		g_bit_syn[i] = bt_syndrome_slow(1ULL << i);

	/* --- byte LUTs + class constants --- */
	bt_scan_tables t;
	memset(&t, 0, sizeof(t));
	for (int b = 0; b < 256; b++) {
		uint64_t sa = 0, sb = 0, sc = 0;
		for (int j = 0; j < 8; j++)
			if ((b >> j) & 1) { sa ^= g_bit_syn[32 + j]; sb ^= g_bit_syn[40 + j]; sc ^= g_bit_syn[48 + j]; }
		t.t_a[b] = (uint32_t)sa; t.t_b[b] = (uint32_t)sb; t.t_c[b] = (uint32_t)sc;
	}
	t.t_56 = (uint32_t)g_bit_syn[56];
	t.c_class[0] = (uint32_t)bt_syndrome_slow(BT_PN ^ ((uint64_t)BT_BARKER_A << 57));
	t.c_class[1] = (uint32_t)bt_syndrome_slow(BT_PN ^ ((uint64_t)BT_BARKER_B << 57));
	BT_CUDA_TRY(cudaMalloc(&ctx->d_tables, sizeof(t)));
	BT_CUDA_TRY(cudaMemcpy(ctx->d_tables, &t, sizeof(t), cudaMemcpyHostToDevice));

	/* --- error patterns --- */
	std::vector<bt_err_slot> ents;
	for (int i = 1; i <= k; i++)
		enumerate(ents, 0, 0, 0, i);
	ctx->err_entries = (long)ents.size();
	ctx->table_k = k;

	/* --- Bloom bitmap over low-32 syndromes (zero syndrome included) --- */
	int blog = 13;
	while (blog < 20 && (1L << blog) < 64L * (long)(ents.size() + 1)) blog++;
	ctx->bloom_log2 = blog;
	std::vector<uint32_t> bloom((size_t)1 << (blog - 5), 0u);
	auto bloom_add = [&](uint32_t s32) {
		uint32_t h1 = bt_bloom_h1(s32, blog), h2 = bt_bloom_h2(s32, blog);
		bloom[h1 >> 5] |= 1u << (h1 & 31);
		bloom[h2 >> 5] |= 1u << (h2 & 31);
	};
	bloom_add(0);
	for (auto &e : ents) bloom_add((uint32_t)e.syn);
	BT_CUDA_TRY(cudaMalloc(&ctx->d_bloom, bloom.size() * sizeof(uint32_t)));
	BT_CUDA_TRY(cudaMemcpy(ctx->d_bloom, bloom.data(), bloom.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));

	/* --- constants of the bulk kernels' exact test: the 34-bit syndrome of PN ^ (legal tail << 57) per
	 * Barker class, and parity masks over codeword bits 32..56 for syndrome bits 0 / 32 / 33 --- */
	ctx->cc[0] = bt_syndrome_slow(BT_PN ^ ((uint64_t)BT_BARKER_A << 57));
	ctx->cc[1] = bt_syndrome_slow(BT_PN ^ ((uint64_t)BT_BARKER_B << 57));
	ctx->m32 = ctx->m33 = ctx->m0 = 0;
	ctx->d_lut7 = NULL; ctx->d_map7 = NULL; ctx->d_map7b = NULL; ctx->d_map7g = NULL; ctx->map7g_log2 = 0;
	for (int j = 0; j < 25; j++) {
		if (g_bit_syn[32 + j] & 1) ctx->m0 |= 1u << j;
		if ((g_bit_syn[32 + j] >> 32) & 1) ctx->m32 |= 1u << j;
		if ((g_bit_syn[32 + j] >> 33) & 1) ctx->m33 |= 1u << j;
	}

	/* v7 (scan_v7.cuh) works on syndrome bits 1..32: field tables A / B / C over
	 * codeword bits 34..40 / 41..48 / 49..56, a first-level map addressed by byte (byte = the
	 * top 16 / 15 value bits, bit = bits 0..2) and a second-level map -- for k <= 2 in shared
	 * memory over value bits 3..19 (word = bits 8..19, bit = bits 3..7), for k = 3 (68 558
	 * reachable values) 2^27 bits in global memory (word = bits 10..31, bit = bits 5..9) */
	if (k <= 3) {
		const size_t m2_words = (size_t)1 << (17 - 5);
		std::vector<uint32_t> l7;
		const int fw7[3] = {7, 8, 8};
		for (int f = 0, pos = 0; f < 3; pos += fw7[f], f++)
			for (uint32_t v = 0; v < (1u << fw7[f]); v++) {
				uint64_t sy = 0;
				for (int j = 0; j < fw7[f]; j++) if ((v >> j) & 1) sy ^= g_bit_syn[34 + pos + j];
				l7.push_back((uint32_t)(sy >> 1));
			}
		BT_CUDA_TRY(cudaMalloc(&ctx->d_lut7, l7.size() * sizeof(uint32_t)));
		BT_CUDA_TRY(cudaMemcpy(ctx->d_lut7, l7.data(), l7.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
		/* two sizes of the first-level map: 2^16 bytes (layout<0>) and 2^15 bytes (layout<1>) */
		for (int variant = 0; variant < 2; variant++) {
			const int lg = variant ? 15 : 16;
			const size_t m1_bytes = (size_t)1 << lg;
			std::vector<uint32_t> map7(m1_bytes / 4 + m2_words, 0u);
			uint8_t *mb = reinterpret_cast<uint8_t *>(map7.data());
			uint32_t *m2 = map7.data() + m1_bytes / 4;
			auto map7_add = [&](uint64_t s34) {
				for (int c = 0; c < 2; c++) {
					uint32_t v = (uint32_t)((s34 ^ ctx->cc[c]) >> 1);
					mb[v >> (32 - lg)] |= (uint8_t)(1u << (v & 7));
					m2[(v >> 8) & (m2_words - 1)] |= 1u << ((v >> 3) & 31);
				}
			};
			map7_add(0);
			for (auto &e : ents) map7_add(e.syn);
			uint32_t **dst = variant ? &ctx->d_map7b : &ctx->d_map7;
			BT_CUDA_TRY(cudaMalloc(dst, map7.size() * sizeof(uint32_t)));
			BT_CUDA_TRY(cudaMemcpy(*dst, map7.data(), map7.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
		}
		if (k == 3) {
			ctx->map7g_log2 = 27;
			std::vector<uint32_t> g((size_t)1 << (27 - 5), 0u);
			auto g_add = [&](uint64_t s34) {
				for (int c = 0; c < 2; c++) {
					uint32_t v = (uint32_t)((s34 ^ ctx->cc[c]) >> 1);
					g[v >> 10] |= 1u << ((v >> 5) & 31);
				}
			};
			g_add(0);
			for (auto &e : ents) g_add(e.syn);
			BT_CUDA_TRY(cudaMalloc(&ctx->d_map7g, g.size() * sizeof(uint32_t)));
			BT_CUDA_TRY(cudaMemcpy(ctx->d_map7g, g.data(), g.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
		}
	}

	/* 4 / 5 errors (982 130 / 11 060 036 reachable values): the v7 kernel's FIRST-level map in
	 * global memory, byte = the top (log2 - 3) value bits, bit = value bits 0..2, sized for
	 * ~1 % (k = 4: 2^27 bits, 16 MiB) and ~2 % (k = 5: 2^29 bits, 64 MiB) false positives */
	if (k >= 4) {
		const int lg = k == 4 ? 27 : 29;
		ctx->map7g_log2 = lg;
		std::vector<uint8_t> g((size_t)1 << (lg - 3), 0);
		auto g_add = [&](uint64_t s34) {
			for (int c = 0; c < 2; c++) {
				uint32_t v = (uint32_t)((s34 ^ ctx->cc[c]) >> 1);
				g[v >> (32 - (lg - 3))] |= (uint8_t)(1u << (v & 7));
			}
		};
		g_add(0);
		for (auto &e : ents) g_add(e.syn);
		BT_CUDA_TRY(cudaMalloc(&ctx->d_map7g, g.size()));
		BT_CUDA_TRY(cudaMemcpy(ctx->d_map7g, g.data(), g.size(), cudaMemcpyHostToDevice));
		std::vector<uint32_t> l7;
		const int fw7[3] = {7, 8, 8};
		for (int f = 0, pos = 0; f < 3; pos += fw7[f], f++)
			for (uint32_t v = 0; v < (1u << fw7[f]); v++) {
				uint64_t sy = 0;
				for (int j = 0; j < fw7[f]; j++) if ((v >> j) & 1) sy ^= g_bit_syn[34 + pos + j];
				l7.push_back((uint32_t)(sy >> 1));
			}
		BT_CUDA_TRY(cudaMalloc(&ctx->d_lut7, l7.size() * sizeof(uint32_t)));
		BT_CUDA_TRY(cudaMemcpy(ctx->d_lut7, l7.data(), l7.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
	}

	/* --- open-addressing map --- */
	ctx->d_err = NULL; ctx->h_err = NULL; ctx->err_log2 = 0;
	if (!ents.empty()) {
		int lg = 4;
		while ((1UL << lg) < 2 * ents.size()) lg++;
		std::vector<bt_err_slot> tab((size_t)1 << lg, bt_err_slot{0, 0});
		uint64_t mask = ((uint64_t)1 << lg) - 1;
		for (auto &e : ents) {
			uint64_t h = bt_err_hash(e.syn, lg);
			while (tab[h].syn) h = (h + 1) & mask;
			tab[h] = e;
		}
		BT_CUDA_TRY(cudaMalloc(&ctx->d_err, tab.size() * sizeof(bt_err_slot)));
		BT_CUDA_TRY(cudaMemcpy(ctx->d_err, tab.data(), tab.size() * sizeof(bt_err_slot), cudaMemcpyHostToDevice));
		ctx->err_log2 = lg;
		/* host copy for the small-call path of the classic btbb_find_ac (find_ac_host.cpp) */
		ctx->h_err = (bt_err_slot *)malloc(tab.size() * sizeof(bt_err_slot));
		if (!ctx->h_err) return btbb_b200_set_error(BTBB_B200_ENOMEM, "tables: out of host memory");
		memcpy(ctx->h_err, tab.data(), tab.size() * sizeof(bt_err_slot));
	}
	return BTBB_B200_OK;
}

void bt_tables_free(btbb_b200_ctx *ctx)
{
	if (ctx->d_tables) cudaFree(ctx->d_tables);
	if (ctx->d_bloom) cudaFree(ctx->d_bloom);
	if (ctx->d_err) cudaFree(ctx->d_err);
	free(ctx->h_err); ctx->h_err = NULL;
	if (ctx->d_lut7) cudaFree(ctx->d_lut7);
	if (ctx->d_map7) cudaFree(ctx->d_map7);
	if (ctx->d_map7b) cudaFree(ctx->d_map7b);
	if (ctx->d_map7g) cudaFree(ctx->d_map7g);
	ctx->d_map7g = NULL;
	ctx->d_lut7 = NULL; ctx->d_map7 = NULL; ctx->d_map7b = NULL;
	ctx->d_tables = NULL; ctx->d_bloom = NULL; ctx->d_err = NULL;
}
