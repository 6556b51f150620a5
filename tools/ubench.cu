// Micro-benchmarks that decided the scan kernel's design (developer tool).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu && tools/ubench
// Reports cycles per warp-instruction per SM for: random LDS, SHFL, both interleaved, and a
// pure 256-bit streaming read (GB/s).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int MODE>   // 0 = LDS random, 1 = SHFL, 2 = both, 3 = LDS conflict-free
__global__ void __launch_bounds__(1024, 1) k_lds_shfl(uint32_t *out, int iters)
{
	__shared__ uint32_t tab[8192];
	for (int i = threadIdx.x; i < 8192; i += blockDim.x) tab[i] = i * 2654435761u;
	__syncthreads();
	uint32_t x = threadIdx.x * 747796405u + 1, acc = 0, t = threadIdx.x;
	for (int i = 0; i < iters; i++) {
		#pragma unroll
		for (int u = 0; u < 8; u++) {
			x = x * 1664525u + 1013904223u;
			if (MODE == 0 || MODE == 2) acc ^= tab[(x >> 8) & 8191];
			if (MODE == 3) acc ^= tab[((x >> 8) & 8160) | (threadIdx.x & 31)];
			if (MODE == 1 || MODE == 2) t ^= __shfl_sync(0xffffffffu, t + u, x >> 27);
		}
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc ^ t;
}

__global__ void __launch_bounds__(1024, 1) k_stream(const uint8_t *p, int64_t n, uint32_t *out)
{
	uint32_t acc = 0;
	int64_t stride = (int64_t)gridDim.x * blockDim.x * 32;
	for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 32; i + 32 <= n; i += stride) {
		uint32_t r[8];
		asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
			     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p + i));
		acc ^= r[0] ^ r[1] ^ r[2] ^ r[3] ^ r[4] ^ r[5] ^ r[6] ^ r[7];
	}
	if (acc == 0x12345) out[0] = acc;
}

template <typename F> static float timeit(F f)
{
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
	f(); cudaDeviceSynchronize();
	cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
	float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main()
{
	cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
	int sms = pr.multiProcessorCount, khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
	uint32_t *out; CK(cudaMalloc(&out, sms * 1024 * 4));
	const int iters = 2000;
	const char *names[4] = {"LDS random (32 lanes, 8192-word table)", "SHFL idx", "LDS random + SHFL interleaved", "LDS conflict-free"};
	float ms[4];
	ms[0] = timeit([&] { k_lds_shfl<0><<<sms, 1024>>>(out, iters); });
	ms[1] = timeit([&] { k_lds_shfl<1><<<sms, 1024>>>(out, iters); });
	ms[2] = timeit([&] { k_lds_shfl<2><<<sms, 1024>>>(out, iters); });
	ms[3] = timeit([&] { k_lds_shfl<3><<<sms, 1024>>>(out, iters); });
	for (int m = 0; m < 4; m++) {
		double winst = 32.0 * iters * 8 * (m == 2 ? 2 : 1);   // warp-instructions of the measured kind per SM
		printf("%-45s %.3f ms  -> %.2f clk per warp-instr per SM (at %d MHz nominal)\n", names[m], ms[m],
		       ms[m] * 1e-3 * khz * 1e3 / winst, khz / 1000);
	}
	int64_t n = (int64_t)4 << 30; uint8_t *buf; CK(cudaMalloc(&buf, n)); CK(cudaMemset(buf, 1, n));
	float s = timeit([&] { k_stream<<<sms, 1024>>>(buf, n, out); });
	printf("streaming 256-bit read of %lld bytes: %.3f ms -> %.0f GB/s\n", (long long)n, s, n / (s * 1e-3) / 1e9);
	float s2 = timeit([&] { k_stream<<<sms * 2, 512>>>(buf, n, out); });
	printf("streaming read, 2 x 512 threads per SM: %.3f ms -> %.0f GB/s\n", s2, n / (s2 * 1e-3) / 1e9);
	return 0;
}
