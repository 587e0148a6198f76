// Micro-benchmark (measurement only): throughput of Blackwell's packed fp32 instructions
// (fma.rn.f32x2 / mul.rn.f32x2 / add.rn.f32x2 -> FFMA2 / FMUL2 / FADD2) against scalar FFMA, alone and
// mixed with integer ALU work, to decide whether the blend kernels should process two pixels per lane.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o bloomscene_b200/_build/probe_ffma2 tools/probe_ffma2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
	unsigned long long d;
	asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
	return d;
}
__device__ __forceinline__ unsigned long long fmul2(unsigned long long a, unsigned long long b)
{
	unsigned long long d;
	asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
	return d;
}
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b)
{
	unsigned long long d;
	asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
	return d;
}
__device__ __forceinline__ unsigned long long pack(float lo, float hi)
{
	return ((unsigned long long)__float_as_uint(hi) << 32) | __float_as_uint(lo);
}

// mode 0: 8 independent scalar FFMA chains; 1: 8 FFMA2 chains (16 FMAs/iter-slot); 2: FMUL2; 3: FADD2;
// 4: 8 scalar FFMA + 8 LOP3/IADD per slot; 5: 4 FFMA2 (same FMAs as 8 scalar) + 8 int ops;
// 6: 8 scalar FFMA + 8 FMNMX (alu pipe); 7: 4 FFMA2 + 8 FMNMX
template <int MODE>
__global__ void __launch_bounds__(256) probe(float* sink, int iters, float a, float b)
{
	float x[8];
	unsigned long long y[8];
	uint32_t z[8];
	float m[8];
#pragma unroll
	for (int i = 0; i < 8; i++) {
		x[i] = threadIdx.x + i;
		y[i] = pack(threadIdx.x + i, threadIdx.x + 2 * i);
		z[i] = threadIdx.x * 7 + i;
		m[i] = threadIdx.x * 0.5f + i;
	}
	const unsigned long long a2 = pack(a, a), b2 = pack(b, b);
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int u = 0; u < 8; u++) {
			if (MODE == 0 || MODE == 4 || MODE == 6) {
#pragma unroll
				for (int i = 0; i < 8; i++)
					x[i] = __fmaf_rn(x[i], a, b);
			}
			if (MODE == 1) {
#pragma unroll
				for (int i = 0; i < 8; i++)
					y[i] = ffma2(y[i], a2, b2);
			}
			if (MODE == 2) {
#pragma unroll
				for (int i = 0; i < 8; i++)
					y[i] = fmul2(y[i], a2);
			}
			if (MODE == 3) {
#pragma unroll
				for (int i = 0; i < 8; i++)
					y[i] = fadd2(y[i], b2);
			}
			if (MODE == 5 || MODE == 7) {
#pragma unroll
				for (int i = 0; i < 4; i++)
					y[i] = ffma2(y[i], a2, b2);
			}
			if (MODE == 4 || MODE == 5) {
#pragma unroll
				for (int i = 0; i < 8; i++)
					z[i] = (z[i] ^ (uint32_t)it) + 0x9e3779b9u;
			}
			if (MODE == 6 || MODE == 7) {
#pragma unroll
				for (int i = 0; i < 8; i++)
					m[i] = fminf(m[i], a) + 0.0f * 0 + 0;   // FMNMX
			}
		}
	}
	float r = 0.f;
#pragma unroll
	for (int i = 0; i < 8; i++)
		r += x[i] + __uint_as_float((uint32_t)y[i]) + __uint_as_float((uint32_t)(y[i] >> 32)) + (float)z[i] + m[i];
	if (r == 123.456f)
		sink[0] = r;
}

template <int MODE>
static void run(const char* name, double fma_per_slot, float* sink, int sms)
{
	const int blocks = sms * 8, iters = 2048;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	float best = 1e30f;
	for (int rep = 0; rep < 5; rep++) {
		cudaEventRecord(e0);
		probe<MODE><<<blocks, 256>>>(sink, iters, 0.999f, 0.001f);
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		float ms = 0.f;
		cudaEventElapsedTime(&ms, e0, e1);
		if (rep > 0 && ms < best)
			best = ms;
	}
	const double slots = 8.0 * iters * 256.0 * blocks;   // (unrolled u) x iters x threads
	const double tflops = 2.0 * fma_per_slot * slots / (best * 1e-3) / 1e12;
	printf("%-44s %8.3f ms  %7.2f TFLOP/s (fp32 flops)\n", name, best, tflops);
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
}

int main()
{
	float* sink;
	cudaMalloc(&sink, 256);
	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	printf("SMs %d\n", sms);
	run<0>("scalar FFMA x8", 8, sink, sms);
	run<1>("FFMA2 x8 (16 FMA)", 16, sink, sms);
	run<2>("FMUL2 x8 (counted as 16 'FMA')", 16, sink, sms);
	run<3>("FADD2 x8 (counted as 16 'FMA')", 16, sink, sms);
	run<4>("scalar FFMA x8 + 8 int (LOP3+IADD)", 8, sink, sms);
	run<5>("FFMA2 x4 (8 FMA) + 8 int", 8, sink, sms);
	run<6>("scalar FFMA x8 + 8 FMNMX", 8, sink, sms);
	run<7>("FFMA2 x4 (8 FMA) + 8 FMNMX", 8, sink, sms);
	cudaError_t e = cudaDeviceSynchronize();
	printf("status %s\n", cudaGetErrorString(e));
	return e != cudaSuccess;
}
