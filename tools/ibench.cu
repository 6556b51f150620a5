// Instruction-issue micro-benchmarks behind scan_v7.cuh's pipe choices (developer tool).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ibench tools/ibench.cu && tools/ibench
// 32 warps per SM, 8 independent dependency chains per thread; reports SMSP cycles per
// warp-instruction for each opcode alone and for ALU/FMA pairs issued alternately.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 8
#define UNROLL 4
#define OP1(name, body) \
__global__ void __launch_bounds__(1024, 1) k_##name(uint32_t *out, int iters, uint32_t s0, uint32_t s1) { \
	uint32_t x[CHAINS], y = s1; \
	for (int j = 0; j < CHAINS; j++) x[j] = threadIdx.x * 2654435761u + j + s0; \
	for (int i = 0; i < iters; i++) { \
		_Pragma("unroll") for (int u = 0; u < UNROLL; u++) { \
			_Pragma("unroll") for (int j = 0; j < CHAINS; j++) { body } } } \
	uint32_t a = 0; for (int j = 0; j < CHAINS; j++) a ^= x[j]; \
	out[blockIdx.x * blockDim.x + threadIdx.x] = a ^ y; }

OP1(imad,   asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(x[j]) : "r"(y));)
OP1(imadhi, asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[j]) : "r"(y));)
OP1(madhi,  asm volatile("mad.hi.u32 %0, %0, %1, %0;" : "+r"(x[j]) : "r"(y));)
OP1(lop3,   asm volatile("lop3.b32 %0, %0, %1, %0, 0x96;" : "+r"(x[j]) : "r"(y));)
OP1(shf,    asm volatile("shf.r.wrap.b32 %0, %0, %1, %1;" : "+r"(x[j]) : "r"(y));)
OP1(prmt,   asm volatile("prmt.b32 %0, %0, %1, 0x5514;" : "+r"(x[j]) : "r"(y));)
OP1(flo,    asm volatile("bfind.u32 %0, %0;" : "+r"(x[j]));)
OP1(brev,   asm volatile("brev.b32 %0, %0;" : "+r"(x[j]));)
OP1(popc,   asm volatile("popc.b32 %0, %0;" : "+r"(x[j]));)
OP1(bmsk,   asm volatile("bmsk.clamp.b32 %0, %0, 1;" : "+r"(x[j]));)
OP1(iadd,   asm volatile("add.u32 %0, %0, %1;" : "+r"(x[j]) : "r"(y));)
OP1(dp4a,   asm volatile("dp4a.u32.u32 %0, %0, %1, %0;" : "+r"(x[j]) : "r"(y));)
OP1(lop_imad,   asm volatile("lop3.b32 %0, %0, %1, %0, 0x96;\n\tmad.lo.u32 %0, %0, %1, %0;" : "+r"(x[j]) : "r"(y));)
OP1(lop_imadhi, asm volatile("lop3.b32 %0, %0, %1, %0, 0x96;\n\tmul.hi.u32 %0, %0, %1;" : "+r"(x[j]) : "r"(y));)
OP1(shf_imad,   asm volatile("shf.r.wrap.b32 %0, %0, %1, %1;\n\tmad.lo.u32 %0, %0, %1, %0;" : "+r"(x[j]) : "r"(y));)
OP1(lop_lop_imad, asm volatile("lop3.b32 %0, %0, %1, %0, 0x96;\n\tlop3.b32 %0, %0, %1, %0, 0x69;\n\tmad.lo.u32 %0, %0, %1, %0;" : "+r"(x[j]) : "r"(y));)
OP1(lop_imad_imad, asm volatile("lop3.b32 %0, %0, %1, %0, 0x96;\n\tmad.lo.u32 %0, %0, %1, %0;\n\tmad.lo.u32 %0, %0, %1, %1;" : "+r"(x[j]) : "r"(y));)

template <int MODE>   // 0 = lane-private LDS.32, 1 = random LDS.U8 over 64 KiB, 2 = random LDS.32 over 64 KiB
__global__ void __launch_bounds__(1024, 1) k_lds(uint32_t *out, int iters)
{
	extern __shared__ uint32_t tab[];
	for (int i = threadIdx.x; i < 32768; i += blockDim.x) tab[i] = i * 2654435761u;
	__syncthreads();
	const uint32_t base = (uint32_t)__cvta_generic_to_shared(tab);
	uint32_t x = threadIdx.x * 747796405u + 1, acc = 0;
	for (int i = 0; i < iters; i++) {
		#pragma unroll
		for (int u = 0; u < 8; u++) {
			x = x * 1664525u + 1013904223u;
			uint32_t v;
			if (MODE == 0) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(base + (((x >> 8) & 0xff80u) | ((threadIdx.x & 31) * 4))));
			else if (MODE == 1) asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(base + (x >> 16)));
			else asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(base + ((x >> 16) & 0xfffcu)));
			acc ^= v;
		}
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
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
	cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
	int sms = pr.multiProcessorCount, khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
	uint32_t *out; cudaMalloc(&out, sms * 1024 * 4);
	const int iters = 4000;
	printf("%s, %d SMs, %d MHz nominal; SMSP cycles per warp-instruction (8 warps per SMSP)\n", pr.name, sms, khz / 1000);
#define RUN(name, n) do { float ms = timeit([&] { k_##name<<<sms, 1024>>>(out, iters, 1u, 0x9e3779b1u); }); \
	double wi = 8.0 * iters * UNROLL * CHAINS * (n); \
	printf("%-16s %8.3f ms  %.3f clk/instr\n", #name, ms, ms * 1e-3 * khz * 1e3 / wi); } while (0)
	RUN(imad, 1); RUN(imadhi, 1); RUN(madhi, 1); RUN(lop3, 1); RUN(shf, 1); RUN(prmt, 1); RUN(flo, 1); RUN(brev, 1);
	RUN(popc, 1); RUN(bmsk, 1); RUN(iadd, 1); RUN(dp4a, 1);
	RUN(lop_imad, 2); RUN(lop_imadhi, 2); RUN(shf_imad, 2); RUN(lop_lop_imad, 3); RUN(lop_imad_imad, 3);
	const char *ln[3] = {"LDS.32 lane-private", "LDS.U8 random 64K", "LDS.32 random 64K"};
	cudaFuncSetAttribute(k_lds<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
	cudaFuncSetAttribute(k_lds<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
	cudaFuncSetAttribute(k_lds<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
	float m0 = timeit([&] { k_lds<0><<<sms, 1024, 131072>>>(out, 2000); });
	float m1 = timeit([&] { k_lds<1><<<sms, 1024, 131072>>>(out, 2000); });
	float m2 = timeit([&] { k_lds<2><<<sms, 1024, 131072>>>(out, 2000); });
	float mm[3] = {m0, m1, m2};
	for (int i = 0; i < 3; i++)
		printf("%-22s %8.3f ms  %.2f SM clk per warp-load\n", ln[i], mm[i], mm[i] * 1e-3 * khz * 1e3 / (32.0 * 2000 * 8));
	return 0;
}
