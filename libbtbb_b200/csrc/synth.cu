/*
 * synth.cu -- K6: synthetic capture generator, host and device entry points
 * (btbb_b200_synth_host / _dev / _planted, include/btbb_b200.h).  Test/bench data only;
 * the detection path never depends on it.
 */
#include <cuda_runtime.h>
#include <string.h>
#include "synth_common.h"
#include "capi_internal.h"

static int synth_cfg_ok(const btbb_b200_synth_cfg *c)
{
	if (!c || c->n_symbols < 0 || c->first_symbol < 0) return 0;
	if (c->stride != 0 && c->stride < 128) return 0;
	return 1;
}

extern "C" int btbb_b200_synth_planted(const btbb_b200_synth_cfg *cfg, int64_t slot, btbb_b200_planted *out)
{
	if (!synth_cfg_ok(cfg) || !out || cfg->stride == 0 || slot < 0)
		return btbb_b200_set_error(BTBB_B200_EINVAL, "synth_planted: bad arguments");
	memset(out, 0, sizeof(*out));
	synth_params(cfg, slot, out);
	return BTBB_B200_OK;
}

extern "C" int btbb_b200_synth_host(const btbb_b200_synth_cfg *cfg, uint8_t *buf)
{
	if (!synth_cfg_ok(cfg) || (!buf && cfg->n_symbols))
		return btbb_b200_set_error(BTBB_B200_EINVAL, "synth_host: bad arguments");
	const int64_t first = cfg->first_symbol, n = cfg->n_symbols;
	for (int64_t i = 0; i < n; i++) {
		int64_t g = first + i;
		buf[i] = (uint8_t)(synth_noise_symbol(cfg->seed, g) ^ synth_flip(cfg->seed, cfg->ber_q32, g));
	}
	if (cfg->stride && n) {
		uint32_t bits[SYNTH_WORDS];
		for (int64_t slot = first / cfg->stride; slot * (int64_t)cfg->stride < first + n; slot++) {
			btbb_b200_planted p;
			synth_params(cfg, slot, &p);
			int len = synth_encode(cfg, &p, bits);
			for (int i = 0; i < len; i++) {
				int64_t g = p.offset + i;
				if (g < first || g >= first + n) continue;
				buf[g - first] = (uint8_t)(((bits[i >> 5] >> (i & 31)) & 1u) ^
							   synth_flip(cfg->seed, cfg->ber_q32, g));
			}
		}
	}
	return BTBB_B200_OK;
}

/* one thread = 16 consecutive symbols of noise (one 16-byte store) */
__global__ void synth_noise_kernel(btbb_b200_synth_cfg cfg, uint8_t *buf)
{
	const int64_t n = cfg.n_symbols;
	const int64_t nchunk = (n + 15) >> 4;
	for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < nchunk;
	     c += (int64_t)gridDim.x * blockDim.x) {
		int64_t i0 = c << 4, g0 = cfg.first_symbol + i0;
		uint64_t w0 = synth_noise_word(cfg.seed, (uint64_t)g0 >> 6);
		uint64_t w1 = synth_noise_word(cfg.seed, ((uint64_t)g0 >> 6) + 1);
		int sh = (int)(g0 & 63);
		uint32_t v = (uint32_t)(sh ? (w0 >> sh) | (w1 << (64 - sh)) : w0) & 0xffffu;
		uint32_t out[4];
		#pragma unroll
		for (int q = 0; q < 4; q++) {
			uint32_t word = 0;
			#pragma unroll
			for (int b = 0; b < 4; b++) {
				int j = q * 4 + b;
				uint32_t s = ((v >> j) & 1u) ^ synth_flip(cfg.seed, cfg.ber_q32, g0 + j);
				word |= s << (8 * b);
			}
			out[q] = word;
		}
		if (i0 + 16 <= n && ((reinterpret_cast<uintptr_t>(buf + i0) & 15) == 0))
			*reinterpret_cast<uint4 *>(buf + i0) = make_uint4(out[0], out[1], out[2], out[3]);
		else
			for (int j = 0; j < 16 && i0 + j < n; j++)
				buf[i0 + j] = (uint8_t)((out[j >> 2] >> (8 * (j & 3))) & 0xff);
	}
}

/* one thread = one planted packet */
__global__ void synth_plant_kernel(btbb_b200_synth_cfg cfg, uint8_t *buf, int64_t slot0, int64_t nslots)
{
	int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
	if (t >= nslots) return;
	btbb_b200_planted p;
	uint32_t bits[SYNTH_WORDS];
	synth_params(&cfg, slot0 + t, &p);
	int len = synth_encode(&cfg, &p, bits);
	for (int i = 0; i < len; i++) {
		int64_t g = p.offset + i;
		if (g < cfg.first_symbol || g >= cfg.first_symbol + cfg.n_symbols) continue;
		buf[g - cfg.first_symbol] = (uint8_t)(((bits[i >> 5] >> (i & 31)) & 1u) ^
						     synth_flip(cfg.seed, cfg.ber_q32, g));
	}
}

extern "C" int btbb_b200_synth_dev(const btbb_b200_synth_cfg *cfg, uint8_t *d_buf, void *cuda_stream)
{
	if (!synth_cfg_ok(cfg) || (!d_buf && cfg->n_symbols))
		return btbb_b200_set_error(BTBB_B200_EINVAL, "synth_dev: bad arguments");
	cudaStream_t st = (cudaStream_t)cuda_stream;
	if (cfg->n_symbols == 0) return BTBB_B200_OK;
	int64_t nchunk = (cfg->n_symbols + 15) >> 4;
	int blocks = (int)((nchunk + 255) / 256 < 148 * 16 ? (nchunk + 255) / 256 : 148 * 16);
	synth_noise_kernel<<<blocks, 256, 0, st>>>(*cfg, d_buf);
	if (cfg->stride) {
		int64_t slot0 = cfg->first_symbol / cfg->stride;
		int64_t slot1 = (cfg->first_symbol + cfg->n_symbols - 1) / cfg->stride;
		int64_t nslots = slot1 - slot0 + 1;
		synth_plant_kernel<<<(unsigned)((nslots + 127) / 128), 128, 0, st>>>(*cfg, d_buf, slot0, nslots);
	}
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess)
		return btbb_b200_set_error(BTBB_B200_ECUDA, cudaGetErrorString(e));
	return BTBB_B200_OK;
}
