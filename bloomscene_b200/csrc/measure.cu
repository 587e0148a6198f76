// Measurement-only kernels (bench / profiling; never on the render path):
//   * count_pairs_kernel: (pixel, instance) pair counts of a finished forward under the REFERENCE's
//     per-pixel semantics (forward.cu:409-452) — the algorithmic work unit of SURVEY.md §8(d);
//   * fp32_probe_kernel: dependent-FFMA throughput, the denominator of the blend kernels' roofline.
#include "common.cuh"
#include "kernels.h"

namespace brs {

namespace {

constexpr int BATCH = 256;

__global__ void __launch_bounds__(256) count_pairs_kernel(BlendFwdArgs a, unsigned long long* out)
{
	__shared__ float4 s_geo[BATCH];
	__shared__ float4 s_con[BATCH];
	const uint32_t tid = threadIdx.x, lane = tid & 31;
	const uint32_t px = blockIdx.x * TILE_X + (tid & 15), py = blockIdx.y * TILE_Y + (tid >> 4);
	const bool inside = px < (uint32_t)a.W && py < (uint32_t)a.H;
	const float pixfx = (float)px, pixfy = (float)py;
	const uint2 range = __ldg(a.ranges + blockIdx.y * a.grid_x + blockIdx.x);
	const int n = (int)(range.y - range.x);
	bool done = !inside;
	float T = 1.0f;
	unsigned long long evaluated = 0, contributing = 0;
	for (int base = 0; base < n; base += BATCH) {
		if (__syncthreads_and(done))
			break;
		const int cnt = min(BATCH, n - base);
		if ((int)tid < cnt) {
			const uint32_t id = __ldg(a.point_list + range.x + base + tid);
			s_geo[tid] = __ldg(a.records + 3 * (size_t)id);
			s_con[tid] = __ldg(a.records + 3 * (size_t)id + 1);
		}
		__syncthreads();
		for (int j = 0; !done && j < cnt; j++) {
			evaluated++;
			const float4 g = s_geo[j];
			const float4 con = s_con[j];
			const float dx = g.x - pixfx, dy = g.y - pixfy;
			const float t1 = __fmul_rn(dy, __fmul_rn(dy, con.z));
			const float s = __fmaf_rn(dx, __fmul_rn(dx, con.x), t1);
			const float t3 = __fmul_rn(dy, __fmul_rn(dx, con.y));
			const float power = __fmaf_rn(s, -0.5f, -t3);
			if (power > 0.0f)
				continue;
			const float alpha = fminf(0.99f, __fmul_rn(con.w, expf(power)));
			if (alpha < 1.0f / 255.0f)
				continue;
			const float test_T = __fmul_rn(T, 1.0f - alpha);
			if (test_T < 0.0001f) {
				done = true;
				continue;
			}
			T = test_T;
			contributing++;
		}
	}
	unsigned long long nb = inside ? (unsigned long long)__ldg(a.n_contrib + (size_t)a.W * py + px) : 0ull;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		evaluated += __shfl_xor_sync(0xffffffffu, evaluated, o);
		contributing += __shfl_xor_sync(0xffffffffu, contributing, o);
		nb += __shfl_xor_sync(0xffffffffu, nb, o);
	}
	if (lane == 0) {
		atomicAdd(out + 0, evaluated);
		atomicAdd(out + 1, contributing);
		atomicAdd(out + 2, nb);
	}
}

__global__ void __launch_bounds__(256) fp32_probe_kernel(float* sink, int iters, float a, float b)
{
	float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
	for (int i = 0; i < iters; i++) {
#pragma unroll
		for (int u = 0; u < 8; u++) {
			x0 = __fmaf_rn(x0, a, b); x1 = __fmaf_rn(x1, a, b); x2 = __fmaf_rn(x2, a, b); x3 = __fmaf_rn(x3, a, b);
			x4 = __fmaf_rn(x4, a, b); x5 = __fmaf_rn(x5, a, b); x6 = __fmaf_rn(x6, a, b); x7 = __fmaf_rn(x7, a, b);
		}
	}
	const float r = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
	if (r == 123.456f)
		sink[0] = r;
}

} // namespace

cudaError_t launch_count_pairs(const BlendFwdArgs& a, unsigned long long* out, cudaStream_t stream)
{
	cudaError_t e = cudaMemsetAsync(out, 0, 3 * sizeof(unsigned long long), stream);
	if (e != cudaSuccess || a.W <= 0 || a.H <= 0)
		return e;
	count_pairs_kernel<<<dim3(a.grid_x, a.grid_y, 1), 256, 0, stream>>>(a, out);
	return cudaGetLastError();
}

double probe_fp32_tflops(cudaStream_t stream)
{
	float* sink = nullptr;
	if (cudaMalloc(&sink, 256) != cudaSuccess)
		return 0.0;
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	const int blocks = sms * 8, iters = 4096;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	double best = 0.0;
	for (int rep = 0; rep < 4; rep++) {
		cudaEventRecord(e0, stream);
		fp32_probe_kernel<<<blocks, 256, 0, stream>>>(sink, iters, 0.999f, 0.001f);
		cudaEventRecord(e1, stream);
		cudaEventSynchronize(e1);
		float ms = 0.f;
		cudaEventElapsedTime(&ms, e0, e1);
		const double flops = 2.0 * 64.0 * iters * 256.0 * blocks;
		if (ms > 0.f && rep > 0)
			best = fmax(best, flops / (ms * 1e-3) / 1e12);
	}
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	cudaFree(sink);
	return best;
}

} // namespace brs
