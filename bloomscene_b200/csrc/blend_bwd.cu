// Backward alpha-blend for sm_100a  <- reference renderCUDA (cuda_rasterizer/backward.cu:399-586).
//
// Same tile / warp / pixel mapping and shared-memory staging as the forward kernel, replayed
// back-to-front.  Differences from the reference that change cost, not results:
//   * the tile starts at max(n_contrib) over its pixels and each warp skips records behind the
//     max(n_contrib) of ITS 32 pixels (the reference walks them with a per-thread `continue`);
//   * per-warp conservative culling with the records' alpha>=1/255 boxes (see blend_fwd.cu);
//   * the per-Gaussian gradient sums are NOT reduced pair by pair.  All nine are fixed linear
//     functions of two per-pixel scalars,
//         w = G * dL_dalpha            u = alpha * T
//         dL_dopacity = S(w)           dL_dcolor_c = S(u * dL_dpixel_c)
//         dL_dmean2D, dL_dconic  <-  S(w dx), S(w dy), S(w dx dx), S(w dx dy), S(w dy dy)
//     so the pixel loop only stores (w, u) of a contributing record into a warp-private staging
//     slot (2 STS).  Every 8 staged records the warp "flushes": lane (r, q) takes record r and
//     pixel row q of the warp's 8x4 block, forms the row's moment / colour sums from 4 LDS.128,
//     the 4 rows are combined with an 8-shuffle transposing tree, and the record leaves as one
//     red.global.add.v2.f32 per lane (+1 scalar) on a 48-byte accumulator row.  That is ~20
//     instructions per contributing (warp, record) instead of ~60 for a 9-value warp reduction,
//     against the reference's 9 x 32-lane atomics per pair (backward.cu:537,574-583);
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
constexpr int BLEND_WARPS = BLEND_THREADS / 32;
constexpr int BATCH = 256;
constexpr int GROUP = 8;         // contributing records staged per warp between flushes (8 records x 4 rows = 32 lanes)
constexpr int STAGE_STRIDE = 68; // floats per staged record: w[32] | u[32] | batch slot + 3 pad -> the flush's LDS.128 are conflict-free

__device__ __forceinline__ float rcp_approx(float x)
{
	float r;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r;
}

__device__ __forceinline__ void red_add_v2(float* addr, float a, float b)
{
	asm volatile("red.relaxed.gpu.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

// Flush `cnt` (<= GROUP) staged records of this warp.  Lane (r = lane >> 2, q = lane & 3) owns
// record r and the q-th row (8 pixels) of the warp's pixel block.
template <bool DEPTH>
__device__ __forceinline__ void flush_group(const float* __restrict__ stage, int cnt,
                                            const StagedRecord* __restrict__ rec, const uint32_t* __restrict__ s_id,
                                            const float* __restrict__ s_dpx, float bxf, float byf, uint32_t lane,
                                            float* __restrict__ accum)
{
	__syncwarp();
	const uint32_t r = lane >> 2, q = lane & 3;
	const bool live = (int)r < cnt;
	float v[10]; // v[9]: dL_dz, only with DEPTH
#pragma unroll
	for (int i = 0; i < 10; i++)
		v[i] = 0.f;
	uint32_t id = 0;
	if (live) {
		const uint32_t idx = (uint32_t)__float_as_int(stage[r * STAGE_STRIDE + 64]);
		const float2 g = *reinterpret_cast<const float2*>(&rec[idx].geo);
		const float4 con = rec[idx].con;
		id = s_id[idx];
		const float4* st = reinterpret_cast<const float4*>(stage + r * STAGE_STRIDE + q * 8);
		const float4 wa = st[0], wb = st[1], ua = st[8], ub = st[9];
		const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
		const float u[8] = {ua.x, ua.y, ua.z, ua.w, ub.x, ub.y, ub.z, ub.w};
		const float dy = g.y - (byf + (float)q);
		const float dx0 = g.x - bxf;
		float S0 = 0.f, Sx = 0.f, Sxx = 0.f;
#pragma unroll
		for (int i = 0; i < 8; i++) {
			const float dx = dx0 - (float)i;
			const float t = w[i] * dx;
			S0 += w[i];
			Sx += t;
			Sxx = fmaf(t, dx, Sxx);
		}
		const float4* dp = reinterpret_cast<const float4*>(s_dpx + q * 8);
		constexpr int NCH = DEPTH ? 4 : 3; // with DEPTH the 4th row of s_dpx holds gD = dL_dD
		float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
		for (int ch = 0; ch < NCH; ch++) {
			const float4 da = dp[ch * 8], db = dp[ch * 8 + 1];
			float s = u[0] * da.x;
			s = fmaf(u[1], da.y, s);
			s = fmaf(u[2], da.z, s);
			s = fmaf(u[3], da.w, s);
			s = fmaf(u[4], db.x, s);
			s = fmaf(u[5], db.y, s);
			s = fmaf(u[6], db.z, s);
			s = fmaf(u[7], db.w, s);
			c[ch] = s;
		}
		const float Sy = dy * S0, Sxy = dy * Sx, Syy = dy * Sy;
		const float o = con.w;
		// dL_dG * G = o * w;  dG_ddelx = -G (dx a + dy b),  dG_ddely = -G (dy c + dx b)
		v[0] = -o * (con.x * Sx + con.y * Sy); // x 0.5*W later
		v[1] = -o * (con.z * Sy + con.y * Sx); // x 0.5*H later
		v[2] = o * Sxx;                         // x -0.5 later
		v[3] = o * Sxy;
		v[4] = o * Syy;
		v[5] = S0;
		v[6] = c[0];
		v[7] = c[1];
		v[8] = c[2];
		v[9] = c[3];
	}
	// combine the 4 rows (lanes q = 0..3 of a record): 9 -> 5 -> 3 values per lane, 8 shuffles.
	// Afterwards lane q holds slots 2q, 2q + 1 in (a, b) and every lane holds slot 8 in c8.
	const bool h2 = lane & 2, h1 = lane & 1;
	float r4[4];
#pragma unroll
	for (int i = 0; i < 4; i++) {
		const float send = h2 ? v[i] : v[i + 4];
		const float keep = h2 ? v[i + 4] : v[i];
		r4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
	}
	float c8 = v[8] + __shfl_xor_sync(0xffffffffu, v[8], 2);
	float a, b;
	{
		const float send = h1 ? r4[0] : r4[2];
		const float keep = h1 ? r4[2] : r4[0];
		a = keep + __shfl_xor_sync(0xffffffffu, send, 1);
	}
	{
		const float send = h1 ? r4[1] : r4[3];
		const float keep = h1 ? r4[3] : r4[1];
		b = keep + __shfl_xor_sync(0xffffffffu, send, 1);
	}
	c8 += __shfl_xor_sync(0xffffffffu, c8, 1);
	float c9 = 0.f;
	if (DEPTH) {
		c9 = v[9] + __shfl_xor_sync(0xffffffffu, v[9], 2);
		c9 += __shfl_xor_sync(0xffffffffu, c9, 1);
	}
	if (live) {
		float* dst = accum + (size_t)id * ACCUM_STRIDE;
		red_add_v2(dst + 2 * q, a, b);
		if (q == 0) {
			if (DEPTH)
				red_add_v2(dst + 8, c8, c9);
			else
				atomicAdd(dst + 8, c8);
		}
	}
	__syncwarp();
}

// DEPTH (extension, off by default = the reference): the depth image carries gradient too.  The forward's
// depth is D / acc (D = sum T alpha z, acc = 1e-6 + sum T alpha, gate acc > 0.5: forward.cu:464-468), so
// D and acc join the recurrence as two more blended channels with per-Gaussian values z and 1, pixel
// gradients gD = g / acc and gA = -g depth / acc and no background term — the shape of the reference's
// commented-out depth lines (backward.cu:539-542) plus the normalisation they lack.  In the scalar
// recurrence this only extends cd = colour . dL_dpixel by z gD + gA; z collects dL_dz = S(u gD) in
// accumulator slot 9.  acc is recovered as 1e-6 + (1 - T_final) (sum T alpha telescopes to 1 - T_final).
template <bool DEPTH>
__global__ void __launch_bounds__(BLEND_THREADS) blend_backward_kernel(BlendBwdArgs a)
{
	__shared__ StagedRecord s_rec[2][BATCH];
	__shared__ uint32_t s_ids[2][BATCH];
	__shared__ __align__(16) float s_stage[BLEND_WARPS][GROUP * STAGE_STRIDE];
	__shared__ __align__(16) float s_dpx[BLEND_WARPS][(DEPTH ? 4 : 3) * 32];
	__shared__ uint32_t s_max[BLEND_WARPS];

	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t tile_x = blockIdx.x, tile_y = blockIdx.y;
	const uint32_t bx = tile_x * TILE_X + (warp & 1) * 8;
	const uint32_t by = tile_y * TILE_Y + (warp >> 1) * 4;
	const uint32_t px = bx + (lane & 7), py = by + (lane >> 3);
	const bool inside = px < (uint32_t)a.W && py < (uint32_t)a.H;
	const uint32_t pix_id = (uint32_t)a.W * py + px;
	const float pixfx = (float)px, pixfy = (float)py;
	const float wx0 = (float)bx, wx1 = (float)(bx + 7), wy0 = (float)by, wy1 = (float)(by + 3);

	const ExpConsts ek = {a.exp_c_scale, a.exp_c_252};
	const uint2 range = __ldg(a.ranges + tile_y * a.grid_x + tile_x);

	const float T_final = inside ? __ldg(a.final_T + pix_id) : 0.0f;
	float T = T_final;
	const int last_contributor = inside ? (int)__ldg(a.n_contrib + pix_id) : 0;

	float dpx0 = 0.f, dpx1 = 0.f, dpx2 = 0.f;
	if (inside) {
		const size_t plane = (size_t)a.W * a.H;
		dpx0 = __ldg(a.dL_dpixels + pix_id);
		dpx1 = __ldg(a.dL_dpixels + plane + pix_id);
		dpx2 = __ldg(a.dL_dpixels + 2 * plane + pix_id);
	}
	s_dpx[warp][lane] = dpx0;
	s_dpx[warp][32 + lane] = dpx1;
	s_dpx[warp][64 + lane] = dpx2;
	float gD = 0.f, gA = 0.f;
	if (DEPTH) {
		if (inside) {
			const float depth = __ldg(a.out_depth + pix_id);
			if (depth > 0.f) { // the forward's acc > 0.5 gate (view-space z > 0.2, so D / acc > 0 exactly when it passed)
				const float acc = 0.000001f + (1.0f - T_final);
				const float g_over_acc = __fdividef(__ldg(a.dL_ddepth + pix_id), acc);
				gD = g_over_acc;
				gA = -g_over_acc * depth;
			}
		}
		s_dpx[warp][96 + lane] = gD;
	}

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
	for (int w = 0; w < BLEND_WARPS; w++)
		tile_max = max(tile_max, (int)s_max[w]);
	const int n = min((int)(range.y - range.x), tile_max);

	float beta = 0.f, last_alpha = 0.f, last_cd = 0.f;
	float bg_dot_dpixel = 0.f;
	bg_dot_dpixel += __ldg(a.bg + 0) * dpx0;
	bg_dot_dpixel += __ldg(a.bg + 1) * dpx1;
	bg_dot_dpixel += __ldg(a.bg + 2) * dpx2;
	const float neg_Tf_bg = -T_final * bg_dot_dpixel;

	float* const stage = s_stage[warp];
	const float* const dpx_rows = s_dpx[warp];
	int staged = 0;
	float* st = stage + lane; // this lane's column of the next free staging slot

	// batches walk the list backwards: slot s of the batch at `base` holds list position n-1-(base+s).
	// Double-buffered cp.async staging as in the forward: batch b+1 lands while batch b is processed.
	const uint32_t* list_end = a.point_list + range.x + (n - 1);
	if ((int)tid < n) {
		const uint32_t id = __ldg(list_end - tid);
		s_ids[0][tid] = id;
		stage_record_async(&s_rec[0][tid], a.records, id);
	}
	uint32_t id_next = (BATCH + (int)tid < n) ? __ldg(list_end - (BATCH + tid)) : 0u;

	for (int base = 0, buf = 0; base < n; base += BATCH, buf ^= 1) {
		cp_async_wait_all();
		__syncthreads(); // buffer `buf` complete and visible; buffer `buf ^ 1` no longer read
		if (base + BATCH + (int)tid < n) {
			s_ids[buf ^ 1][tid] = id_next;
			stage_record_async(&s_rec[buf ^ 1][tid], a.records, id_next);
		}
		id_next = (base + 2 * BATCH + (int)tid < n) ? __ldg(list_end - (base + 2 * BATCH + tid)) : 0u;
		const StagedRecord* rec = s_rec[buf];
		const uint32_t* s_id = s_ids[buf];
		const int cnt = min(BATCH, n - base);

		// slot idx is list position n-1-base-idx; it is in front of this pixel's last contributor
		// iff idx > first_live (and of this warp's iff idx > warp_first_live)
		const int first_live = n - 1 - base - last_contributor;
		const int warp_first_live = n - 1 - base - warp_max;

		for (int c0 = 0; c0 < cnt; c0 += 32) {
			const int e = c0 + (int)lane;
			bool hit = false;
			if (e < cnt && e > warp_first_live) {
				hit = block_may_contribute(rec[e].geo, rec[e].con, wx0, wx1, wy0, wy1);
			}
			uint32_t mask = __ballot_sync(0xffffffffu, hit);
			while (mask) {
				const int j = __ffs(mask) - 1;
				mask &= mask - 1;
				const int idx = c0 + j;
				const StagedRecord* r = rec + idx;
				const float2 g = *reinterpret_cast<const float2*>(&r->geo);
				const float4 con = r->con;
				const float dx = g.x - pixfx;
				const float dy = g.y - pixfy;
				const float t1 = __fmul_rn(dy, __fmul_rn(dy, con.z));
				const float s = __fmaf_rn(dx, __fmul_rn(dx, con.x), t1);
				const float t3 = __fmul_rn(dy, __fmul_rn(dx, con.y));
				const float power = __fmaf_rn(s, -0.5f, -t3);
				const float G = expf_exact(power, ek);
				const float alpha = fminf(0.99f, __fmul_rn(con.w, G));
				const bool ok = (idx > first_live) && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
				if (!__any_sync(0xffffffffu, ok))
					continue;

				const float4 col = r->col;
				float w_ = 0.f, u_ = 0.f;
				if (ok) {
					// 1 - alpha >= 0.01 (alpha is clamped to 0.99): one MUFU.RCP (<= 1 ulp) replaces the
					// reference's two IEEE divisions; gradients are tolerance-checked (rel-L2 <= 1e-4)
					const float rcp = rcp_approx(1.f - alpha);
					T = T * rcp;
					u_ = alpha * T; // dchannel_dcolor

					// reference backward.cu:519-533 keeps, per channel, the colour accumulated BEHIND this record
					// (accum_rec) and sums (c - accum_rec) * dL_dpixel over the channels.  Only that sum is
					// needed, so the recurrence runs on its projection: beta = accum_rec . dL_dpixel and
					// cd = colour . dL_dpixel are scalars.
					beta = fmaf(last_alpha, last_cd - beta, beta);
					float cd = fmaf(col.z, dpx2, fmaf(col.y, dpx1, col.x * dpx0));
					if (DEPTH)
						cd = fmaf(col.w, gD, cd) + gA;
					float dL_dalpha = (cd - beta) * T;
					last_cd = cd;
					last_alpha = alpha;
					dL_dalpha += neg_Tf_bg * rcp;
					w_ = G * dL_dalpha;
				}
				st[0] = w_;
				st[32] = u_;
				if (lane == 0)
					st[64] = __int_as_float(idx); // the record's batch slot rides in the row's padding
				st += STAGE_STRIDE;
				if (++staged == GROUP) {
					flush_group<DEPTH>(stage, GROUP, rec, s_id, dpx_rows, wx0, wy0, lane, a.accum);
					staged = 0;
					st = stage + lane;
				}
			}
		}
		// staged slots refer to this batch's shared records: flush before they are overwritten
		if (staged) {
			flush_group<DEPTH>(stage, staged, rec, s_id, dpx_rows, wx0, wy0, lane, a.accum);
			staged = 0;
			st = stage + lane;
		}
	}
}

} // namespace

cudaError_t launch_blend_backward(const BlendBwdArgs& args, cudaStream_t stream)
{
	BlendBwdArgs a = args;
	const ExpConsts ek = exp_consts();
	a.exp_c_scale = ek.c_scale;
	a.exp_c_252 = ek.c_252;
	if (a.W <= 0 || a.H <= 0)
		return cudaSuccess;
	dim3 grid(a.grid_x, a.grid_y, 1);
	if (a.dL_ddepth != nullptr && a.out_depth != nullptr)
		blend_backward_kernel<true><<<grid, BLEND_THREADS, 0, stream>>>(a);
	else
		blend_backward_kernel<false><<<grid, BLEND_THREADS, 0, stream>>>(a);
	count_launch();
	return cudaGetLastError();
}

} // namespace brs
