// Backward alpha-blend for sm_100a  <- reference renderCUDA (cuda_rasterizer/backward.cu:399-586).
//
// Same tile / warp / pixel mapping and shared-memory staging as the forward kernel, replayed
// back-to-front.  Differences from the reference that change cost, not results:
//   * the tile starts at max(n_contrib) over its pixels and each warp skips records behind the
//     max(n_contrib) of ITS 32 pixels (the reference walks them with a per-thread `continue`);
//   * per-warp conservative culling with the records' alpha>=1/255 boxes (see blend_fwd.cu);
//   * the nine per-Gaussian partial gradients are reduced across the warp with a transposing
//     shuffle tree (14 shuffles for 9 values) and leave as ONE predicated RED.ADD.F32 instruction
//     whose 9 active lanes hit 9 consecutive floats of a 48-byte per-Gaussian accumulator —
//     instead of 9 separate 32-lane atomics per contributing pair (backward.cu:537,574-583);
//   * constant factors (0.5*W, 0.5*H, -0.5) are applied once per Gaussian in the preprocess
//     backward instead of once per pair.
// Parity quirks kept: the alpha derivative ignores the min(0.99,.) clamp, T is recovered by
// division from T_final (backward.cu:513-517), and depth carries no gradient (all uses commented
// out in the reference, backward.cu:443-554).  The alpha tests use the forward's pinned arithmetic
// so exactly the same pairs contribute.
#include "common.cuh"
#include "kernels.h"

namespace brs {

namespace {

constexpr int BLEND_THREADS = TILE_X * TILE_Y;
constexpr int BATCH = 256;

// Sum nine per-lane values over the warp.  On return: lanes with (lane & 3) == 0 hold the total of
// slot (lane >> 2) in `r8`, and every lane holds the total of slot 8 in `r9`.
__device__ __forceinline__ void warp_reduce9(const float v[9], uint32_t lane, float& r8, float& r9)
{
	const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
	float r[4];
#pragma unroll
	for (int i = 0; i < 4; i++) {
		const float send = h16 ? v[i] : v[i + 4];
		const float keep = h16 ? v[i + 4] : v[i];
		r[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
	}
	float q[2];
#pragma unroll
	for (int i = 0; i < 2; i++) {
		const float send = h8 ? r[i] : r[i + 2];
		const float keep = h8 ? r[i + 2] : r[i];
		q[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
	}
	{
		const float send = h4 ? q[0] : q[1];
		const float keep = h4 ? q[1] : q[0];
		r8 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
	}
	r8 += __shfl_xor_sync(0xffffffffu, r8, 2);
	r8 += __shfl_xor_sync(0xffffffffu, r8, 1);
	float w = v[8];
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
		w += __shfl_xor_sync(0xffffffffu, w, o);
	r9 = w;
}

__global__ void __launch_bounds__(BLEND_THREADS) blend_backward_kernel(BlendBwdArgs a)
{
	__shared__ float4 s_geo[BATCH];
	__shared__ float4 s_con[BATCH];
	__shared__ float4 s_col[BATCH];
	__shared__ uint32_t s_id[BATCH];
	__shared__ uint32_t s_max[BLEND_THREADS / 32];

	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t tile_x = blockIdx.x, tile_y = blockIdx.y;
	const uint32_t bx = tile_x * TILE_X + (warp & 1) * 8;
	const uint32_t by = tile_y * TILE_Y + (warp >> 1) * 4;
	const uint32_t px = bx + (lane & 7), py = by + (lane >> 3);
	const bool inside = px < (uint32_t)a.W && py < (uint32_t)a.H;
	const uint32_t pix_id = (uint32_t)a.W * py + px;
	const float pixfx = (float)px, pixfy = (float)py;
	const float wx0 = (float)bx, wx1 = (float)(bx + 7), wy0 = (float)by, wy1 = (float)(by + 3);

	const uint2 range = __ldg(a.ranges + tile_y * a.grid_x + tile_x);

	const float T_final = inside ? __ldg(a.final_T + pix_id) : 0.0f;
	float T = T_final;
	const int last_contributor = inside ? (int)__ldg(a.n_contrib + pix_id) : 0;

	// records at list positions >= max(n_contrib) are skipped by every pixel of the warp / tile
	int warp_max = last_contributor;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
		warp_max = max(warp_max, __shfl_xor_sync(0xffffffffu, warp_max, o));
	if (lane == 0)
		s_max[warp] = (uint32_t)warp_max;
	__syncthreads();
	int tile_max = 0;
#pragma unroll
	for (int w = 0; w < BLEND_THREADS / 32; w++)
		tile_max = max(tile_max, (int)s_max[w]);
	const int n = min((int)(range.y - range.x), tile_max);

	float accum_rec0 = 0.f, accum_rec1 = 0.f, accum_rec2 = 0.f;
	float dpx0 = 0.f, dpx1 = 0.f, dpx2 = 0.f;
	if (inside) {
		const size_t plane = (size_t)a.W * a.H;
		dpx0 = __ldg(a.dL_dpixels + pix_id);
		dpx1 = __ldg(a.dL_dpixels + plane + pix_id);
		dpx2 = __ldg(a.dL_dpixels + 2 * plane + pix_id);
	}
	float last_alpha = 0.f, last_c0 = 0.f, last_c1 = 0.f, last_c2 = 0.f;
	float bg_dot_dpixel = 0.f;
	bg_dot_dpixel += __ldg(a.bg + 0) * dpx0;
	bg_dot_dpixel += __ldg(a.bg + 1) * dpx1;
	bg_dot_dpixel += __ldg(a.bg + 2) * dpx2;

	// batches walk the list backwards: shared slot s holds list position n-1-(base+s)
	for (int base = 0; base < n; base += BATCH) {
		__syncthreads();
		const int cnt = min(BATCH, n - base);
		if ((int)tid < cnt) {
			const uint32_t id = __ldg(a.point_list + range.x + (n - 1 - (base + (int)tid)));
			const float4* rec = a.records + 3 * (size_t)id;
			s_id[tid] = id;
			s_geo[tid] = __ldg(rec);
			s_con[tid] = __ldg(rec + 1);
			s_col[tid] = __ldg(rec + 2);
		}
		__syncthreads();

		for (int c0 = 0; c0 < cnt; c0 += 32) {
			// first list position of this chunk is the largest; skip whole chunk if behind the warp
			const int e = c0 + (int)lane;
			const int pos_e = n - 1 - (base + e);
			bool hit = false;
			if (e < cnt && pos_e < warp_max) {
				const float4 g = s_geo[e];
				hit = (g.x + g.z >= wx0) && (g.x - g.z <= wx1) && (g.y + g.w >= wy0) && (g.y - g.w <= wy1);
			}
			uint32_t mask = __ballot_sync(0xffffffffu, hit);
			while (mask) {
				const int j = __ffs(mask) - 1;
				mask &= mask - 1;
				const int idx = c0 + j;
				const int pos = n - 1 - (base + idx);
				const float4 g = s_geo[idx];
				const float4 con = s_con[idx];
				const float dx = g.x - pixfx;
				const float dy = g.y - pixfy;
				const float t1 = __fmul_rn(dy, __fmul_rn(dy, con.z));
				const float s = __fmaf_rn(dx, __fmul_rn(dx, con.x), t1);
				const float t3 = __fmul_rn(dy, __fmul_rn(dx, con.y));
				const float power = __fmaf_rn(s, -0.5f, -t3);
				const float G = expf(power);
				const float alpha = fminf(0.99f, __fmul_rn(con.w, G));
				const bool ok = (pos < last_contributor) && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
				if (!__any_sync(0xffffffffu, ok))
					continue;

				const float4 col = s_col[idx];
				float v[9];
#pragma unroll
				for (int i = 0; i < 9; i++)
					v[i] = 0.f;
				if (ok) {
					// 1 - alpha >= 0.01 (alpha is clamped to 0.99): one MUFU.RCP (<= 1 ulp) replaces the
					// reference's two IEEE divisions; gradients are tolerance-checked (rel-L2 <= 1e-4)
					const float rcp = __fdividef(1.0f, 1.f - alpha);
					T = T * rcp;
					const float dchannel_dcolor = alpha * T;

					float dL_dalpha = 0.0f;
					accum_rec0 = last_alpha * last_c0 + (1.f - last_alpha) * accum_rec0;
					last_c0 = col.x;
					dL_dalpha += (col.x - accum_rec0) * dpx0;
					accum_rec1 = last_alpha * last_c1 + (1.f - last_alpha) * accum_rec1;
					last_c1 = col.y;
					dL_dalpha += (col.y - accum_rec1) * dpx1;
					accum_rec2 = last_alpha * last_c2 + (1.f - last_alpha) * accum_rec2;
					last_c2 = col.z;
					dL_dalpha += (col.z - accum_rec2) * dpx2;
					dL_dalpha *= T;
					last_alpha = alpha;
					dL_dalpha += (-T_final * rcp) * bg_dot_dpixel;

					const float dL_dG = con.w * dL_dalpha;
					const float gdx = G * dx;
					const float gdy = G * dy;
					const float dG_ddelx = -gdx * con.x - gdy * con.y;
					const float dG_ddely = -gdy * con.z - gdx * con.y;

					v[0] = dL_dG * dG_ddelx; // x 0.5*W later
					v[1] = dL_dG * dG_ddely; // x 0.5*H later
					v[2] = gdx * dx * dL_dG; // x -0.5 later
					v[3] = gdx * dy * dL_dG;
					v[4] = gdy * dy * dL_dG;
					v[5] = G * dL_dalpha;
					v[6] = dchannel_dcolor * dpx0;
					v[7] = dchannel_dcolor * dpx1;
					v[8] = dchannel_dcolor * dpx2;
				}
				float r8, r9;
				warp_reduce9(v, lane, r8, r9);
				float* dst = a.accum + (size_t)s_id[idx] * ACCUM_STRIDE;
				if ((lane & 3) == 0)
					atomicAdd(dst + (lane >> 2), r8);
				else if (lane == 1)
					atomicAdd(dst + 8, r9);
			}
		}
	}
}

} // namespace

cudaError_t launch_blend_backward(const BlendBwdArgs& a, cudaStream_t stream)
{
	if (a.W <= 0 || a.H <= 0)
		return cudaSuccess;
	dim3 grid(a.grid_x, a.grid_y, 1);
	blend_backward_kernel<<<grid, BLEND_THREADS, 0, stream>>>(a);
	count_launch();
	return cudaGetLastError();
}

} // namespace brs
