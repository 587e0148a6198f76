// Fused epilogue of BloomScene's neural-Gaussian generation (the step right before the
// rasterizer) <- reference gaussian_renderer/__init__.py:168-203: after the opacity / colour / covariance MLPs the
// reference builds a [N*K, 25] concatenation (repeat of [grid_scaling | anchor], colour, scale_rot, offsets),
// boolean-indexes it with mask = neural_opacity > 0, splits it again and post-processes the pieces:
//   scaling = grid_scaling[:, 3:] * sigmoid(scale_rot[:, :3])        rot = normalize(scale_rot[:, 3:7])
//   xyz     = anchor + offsets * grid_scaling[:, :3]                 opacity = neural_opacity[mask]
// i.e. ~12 torch kernels and two [N*K, 25] temporaries.  Here ONE kernel does mask -> stable compaction -> the
// arithmetic and writes xyz / colour / opacity / scaling / rot straight in the layout the rasterizer reads, and ONE
// kernel does the whole backward.  One thread per anchor (its K offsets are consecutive rows, so the compaction is
// an exclusive scan of per-anchor counts: decoupled look-back over blocks, as in the binning).
#include "common.cuh"
#include "kernels.h"

namespace brs {

namespace {

constexpr int NG_THREADS = 128;
constexpr uint32_t NG_LOCAL = 1u << 30, NG_INCL = 2u << 30, NG_FLAGS = 3u << 30;

__device__ __forceinline__ uint32_t ng_ld(const uint32_t* p)
{
	uint32_t v;
	asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void ng_st(uint32_t* p, uint32_t v)
{
	asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(NG_THREADS) neural_forward_kernel(NeuralFwdArgs a)
{
	__shared__ uint32_t s_warp[NG_THREADS / 32];
	__shared__ uint32_t s_bcast[2];
	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid == 0)
		s_bcast[0] = atomicAdd(a.ticket, 1u);
	__syncthreads();
	const uint32_t blk = s_bcast[0];
	const int n = (int)(blk * NG_THREADS + tid);
	const int K = a.K;
	// which of this anchor's K offsets survive (neural_opacity > 0), as a bit mask (K <= 32)
	uint32_t keep = 0;
	if (n < a.N) {
		for (int k = 0; k < K; k++)
			if (__ldg(a.neural_opacity + (size_t)n * K + k) > 0.0f)
				keep |= 1u << k;
	}
	const uint32_t cnt = __popc(keep);
	// block-exclusive scan of the counts
	uint32_t incl = cnt;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
		if (lane >= (uint32_t)o)
			incl += v;
	}
	if (lane == 31)
		s_warp[warp] = incl;
	__syncthreads();
	uint32_t warp_off = 0, block_total = 0;
#pragma unroll
	for (int w = 0; w < NG_THREADS / 32; w++) {
		if ((uint32_t)w < warp)
			warp_off += s_warp[w];
		block_total += s_warp[w];
	}
	// chained scan over blocks: decoupled look-back by the first warp, 32 predecessors per poll
	if (warp == 0) {
		if (lane == 0)
			ng_st(a.status + blk, (blk == 0 ? NG_INCL : NG_LOCAL) | block_total);
		uint32_t excl = 0;
		int p = (int)blk - 1;
		while (p >= 0) {
			const int q = p - (int)lane;
			const uint32_t v = (q >= 0) ? ng_ld(a.status + q) : NG_INCL;
			const uint32_t f = v & NG_FLAGS;
			const uint32_t not_ready = __ballot_sync(0xffffffffu, f == 0);
			const uint32_t inclusive = __ballot_sync(0xffffffffu, f == NG_INCL);
			const int first_nr = not_ready ? __ffs(not_ready) - 1 : 32;
			const int first_in = inclusive ? __ffs(inclusive) - 1 : 32;
			const int take = (first_in < first_nr) ? first_in + 1 : first_nr;
			uint32_t c = ((int)lane < take) ? (v & ~NG_FLAGS) : 0u;
#pragma unroll
			for (int o = 16; o > 0; o >>= 1)
				c += __shfl_xor_sync(0xffffffffu, c, o);
			excl += c;
			if (first_in < first_nr)
				break;
			p -= first_nr;
		}
		if (lane == 0) {
			if (blk > 0)
				ng_st(a.status + blk, NG_INCL | (excl + block_total));
			s_bcast[1] = excl;
			if (blk == gridDim.x - 1)
				*a.count = excl + block_total; // M: rows that survive (tickets are handed out in order: this is the last block)
		}
	}
	__syncthreads();
	if (n >= a.N)
		return;
	uint32_t row = s_bcast[1] + warp_off + incl - cnt; // first output row of this anchor
	const float* gs = a.grid_scaling + 6 * (size_t)n;
	const float g0 = __ldg(gs), g1 = __ldg(gs + 1), g2 = __ldg(gs + 2), g3 = __ldg(gs + 3), g4 = __ldg(gs + 4), g5 = __ldg(gs + 5);
	const float ax = __ldg(a.anchor + 3 * (size_t)n), ay = __ldg(a.anchor + 3 * (size_t)n + 1), az = __ldg(a.anchor + 3 * (size_t)n + 2);
	for (int k = 0; k < K; k++) {
		const size_t j = (size_t)n * K + k;
		if (!((keep >> k) & 1u)) {
			a.index[j] = -1;
			continue;
		}
		a.index[j] = (int)row;
		const float* sr = a.scale_rot + 7 * j;
		const float s0 = __ldg(sr), s1 = __ldg(sr + 1), s2 = __ldg(sr + 2);
		const float q0 = __ldg(sr + 3), q1 = __ldg(sr + 4), q2 = __ldg(sr + 5), q3 = __ldg(sr + 6);
		const float* of = a.offsets + 3 * j;
		float* o;
		o = a.xyz + 3 * (size_t)row;
		o[0] = ax + __ldg(of) * g0; o[1] = ay + __ldg(of + 1) * g1; o[2] = az + __ldg(of + 2) * g2;
		o = a.out_color + 3 * (size_t)row;
		o[0] = __ldg(a.color + 3 * j); o[1] = __ldg(a.color + 3 * j + 1); o[2] = __ldg(a.color + 3 * j + 2);
		a.out_opacity[row] = __ldg(a.neural_opacity + j);
		o = a.scaling + 3 * (size_t)row;
		o[0] = g3 * (1.0f / (1.0f + expf(-s0))); o[1] = g4 * (1.0f / (1.0f + expf(-s1))); o[2] = g5 * (1.0f / (1.0f + expf(-s2)));
		// torch.nn.functional.normalize: v / max(||v||, 1e-12)
		const float inv = 1.0f / fmaxf(sqrtf(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3), 1e-12f);
		o = a.rot + 4 * (size_t)row;
		o[0] = q0 * inv; o[1] = q1 * inv; o[2] = q2 * inv; o[3] = q3 * inv;
		row++;
	}
}

// One thread per anchor: gradients of its K rows (zeros for the masked-out ones) and, summed over the rows that
// survived, of the anchor position and its grid scaling - no atomics.
__global__ void __launch_bounds__(NG_THREADS) neural_backward_kernel(NeuralBwdArgs a)
{
	const int n = blockIdx.x * NG_THREADS + threadIdx.x;
	if (n >= a.N)
		return;
	const int K = a.K;
	const float* gs = a.grid_scaling + 6 * (size_t)n;
	const float g0 = __ldg(gs), g1 = __ldg(gs + 1), g2 = __ldg(gs + 2), g3 = __ldg(gs + 3), g4 = __ldg(gs + 4), g5 = __ldg(gs + 5);
	float dA0 = 0.f, dA1 = 0.f, dA2 = 0.f, dG0 = 0.f, dG1 = 0.f, dG2 = 0.f, dG3 = 0.f, dG4 = 0.f, dG5 = 0.f;
	for (int k = 0; k < K; k++) {
		const size_t j = (size_t)n * K + k;
		const int row = __ldg(a.index + j);
		float d_op = 0.f, d_c0 = 0.f, d_c1 = 0.f, d_c2 = 0.f, d_o0 = 0.f, d_o1 = 0.f, d_o2 = 0.f;
		float d_sr[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
		if (row >= 0) {
			const float* gx = a.d_xyz + 3 * (size_t)row;
			const float x0 = __ldg(gx), x1 = __ldg(gx + 1), x2 = __ldg(gx + 2);
			const float* of = a.offsets + 3 * j;
			dA0 += x0; dA1 += x1; dA2 += x2;
			d_o0 = x0 * g0; d_o1 = x1 * g1; d_o2 = x2 * g2;
			dG0 += x0 * __ldg(of); dG1 += x1 * __ldg(of + 1); dG2 += x2 * __ldg(of + 2);
			d_c0 = __ldg(a.d_color + 3 * (size_t)row); d_c1 = __ldg(a.d_color + 3 * (size_t)row + 1); d_c2 = __ldg(a.d_color + 3 * (size_t)row + 2);
			d_op = __ldg(a.d_opacity + row);
			const float* sr = a.scale_rot + 7 * j;
			const float* ds = a.d_scaling + 3 * (size_t)row;
			const float gsc[3] = {g3, g4, g5};
			float* dGs[3] = {&dG3, &dG4, &dG5};
#pragma unroll
			for (int c = 0; c < 3; c++) {
				const float sg = 1.0f / (1.0f + expf(-__ldg(sr + c)));
				const float up = __ldg(ds + c);
				*dGs[c] += up * sg;
				d_sr[c] = up * gsc[c] * sg * (1.0f - sg);
			}
			const float q0 = __ldg(sr + 3), q1 = __ldg(sr + 4), q2 = __ldg(sr + 5), q3 = __ldg(sr + 6);
			const float* dr = a.d_rot + 4 * (size_t)row;
			const float r0 = __ldg(dr), r1 = __ldg(dr + 1), r2 = __ldg(dr + 2), r3 = __ldg(dr + 3);
			const float len = sqrtf(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
			if (len >= 1e-12f) { // y = q / |q|: dq = (r - y (y . r)) / |q|
				const float inv = 1.0f / len;
				const float y0 = q0 * inv, y1 = q1 * inv, y2 = q2 * inv, y3 = q3 * inv;
				const float p = y0 * r0 + y1 * r1 + y2 * r2 + y3 * r3;
				d_sr[3] = (r0 - y0 * p) * inv; d_sr[4] = (r1 - y1 * p) * inv; d_sr[5] = (r2 - y2 * p) * inv; d_sr[6] = (r3 - y3 * p) * inv;
			} else { // clamped denominator: y = q / 1e-12
				d_sr[3] = r0 * 1e12f; d_sr[4] = r1 * 1e12f; d_sr[5] = r2 * 1e12f; d_sr[6] = r3 * 1e12f;
			}
		}
		a.d_neural_opacity[j] = d_op;
		float* o = a.d_color_in + 3 * j;
		o[0] = d_c0; o[1] = d_c1; o[2] = d_c2;
		o = a.d_offsets + 3 * j;
		o[0] = d_o0; o[1] = d_o1; o[2] = d_o2;
		o = a.d_scale_rot + 7 * j;
#pragma unroll
		for (int c = 0; c < 7; c++)
			o[c] = d_sr[c];
	}
	float* o = a.d_anchor + 3 * (size_t)n;
	o[0] = dA0; o[1] = dA1; o[2] = dA2;
	o = a.d_grid_scaling + 6 * (size_t)n;
	o[0] = dG0; o[1] = dG1; o[2] = dG2; o[3] = dG3; o[4] = dG4; o[5] = dG5;
}

} // namespace

size_t neural_scratch_bytes(int N) { return align_up(256 + sizeof(uint32_t) * (size_t)((N + NG_THREADS - 1) / NG_THREADS + 1), 256); }

cudaError_t launch_neural_forward(const NeuralFwdArgs& args, void* scratch, cudaStream_t stream)
{
	NeuralFwdArgs a = args;
	cudaError_t e = cudaMemsetAsync(scratch, 0, neural_scratch_bytes(a.N), stream);
	if (e != cudaSuccess)
		return e;
	a.ticket = static_cast<uint32_t*>(scratch);
	a.status = a.ticket + 64;
	if (a.N <= 0) {
		return cudaMemsetAsync(a.count, 0, sizeof(uint32_t), stream);
	}
	neural_forward_kernel<<<(a.N + NG_THREADS - 1) / NG_THREADS, NG_THREADS, 0, stream>>>(a);
	count_launch();
	return cudaGetLastError();
}

cudaError_t launch_neural_backward(const NeuralBwdArgs& a, cudaStream_t stream)
{
	if (a.N <= 0)
		return cudaSuccess;
	neural_backward_kernel<<<(a.N + NG_THREADS - 1) / NG_THREADS, NG_THREADS, 0, stream>>>(a);
	count_launch();
	return cudaGetLastError();
}

} // namespace brs
