/*
 * oracle/synth_gen.c -- TEST INFRASTRUCTURE ONLY: the synthetic capture generator
 * (libbtbb_b200/csrc/synth_common.h, the one definition of the bench / test input) compiled on its
 * own, multi-threaded, so that the CPU reference arm of bench.py can build its input without
 * mapping the product library into its process.  Same bytes as btbb_b200_synth_host / _dev.
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include "../libbtbb_b200/csrc/synth_common.h"

typedef struct { btbb_b200_synth_cfg cfg; uint8_t *buf; } job_t;

static void fill(const btbb_b200_synth_cfg *cfg, uint8_t *buf)
{
	const int64_t first = cfg->first_symbol, n = cfg->n_symbols;
	int64_t i, slot;
	for (i = 0; i < n; ) {
		/* one noise word covers 64 symbols */
		const int64_t g = first + i;
		const uint64_t w = synth_noise_word(cfg->seed, (uint64_t)g >> 6);
		int b = (int)(g & 63);
		for (; b < 64 && i < n; b++, i++)
			buf[i] = (uint8_t)(((w >> b) & 1u) ^ synth_flip(cfg->seed, cfg->ber_q32, first + i));
	}
	if (cfg->stride && n) {
		uint32_t bits[SYNTH_WORDS];
		slot = first / cfg->stride - 1;
		if (slot < 0) slot = 0;
		for (; slot * (int64_t)cfg->stride < first + n; slot++) {
			btbb_b200_planted p;
			int len, k;
			synth_params(cfg, slot, &p);
			len = synth_encode(cfg, &p, bits);
			for (k = 0; k < len; k++) {
				const int64_t g = p.offset + k;
				if (g < first || g >= first + n) continue;
				buf[g - first] = (uint8_t)(((bits[k >> 5] >> (k & 31)) & 1u) ^ synth_flip(cfg->seed, cfg->ber_q32, g));
			}
		}
	}
}

static void *worker(void *a) { job_t *j = (job_t *)a; fill(&j->cfg, j->buf); return NULL; }

int orc_synth(const btbb_b200_synth_cfg *cfg, uint8_t *buf, int threads)
{
	job_t *jobs;
	pthread_t *tid;
	int t;
	int64_t per;
	if (!cfg || (!buf && cfg->n_symbols) || cfg->n_symbols < 0 || cfg->first_symbol < 0 || (cfg->stride && cfg->stride < 128)) return -1;
	if (threads < 1) threads = 1;
	if (threads > 256) threads = 256;
	per = (cfg->n_symbols + threads - 1) / threads;
	per = (per + 63) & ~(int64_t)63;
	jobs = (job_t *)calloc((size_t)threads, sizeof(*jobs));
	tid = (pthread_t *)calloc((size_t)threads, sizeof(*tid));
	if (!jobs || !tid) { free(jobs); free(tid); return -1; }
	for (t = 0; t < threads; t++) {
		int64_t b = per * t, e = b + per < cfg->n_symbols ? b + per : cfg->n_symbols;
		jobs[t].cfg = *cfg;
		jobs[t].cfg.first_symbol = cfg->first_symbol + b;
		jobs[t].cfg.n_symbols = e > b ? e - b : 0;
		jobs[t].buf = buf + b;
		if (pthread_create(&tid[t], NULL, worker, &jobs[t])) { worker(&jobs[t]); tid[t] = 0; }
	}
	for (t = 0; t < threads; t++) if (tid[t]) pthread_join(tid[t], NULL);
	free(jobs); free(tid);
	return 0;
}
