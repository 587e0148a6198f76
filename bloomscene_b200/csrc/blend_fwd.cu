// Forward alpha-blend for sm_100a  <- reference renderCUDA (cuda_rasterizer/forward.cu:341-471).
//
// One CTA per 16x16 tile (tile ids are parity outputs), 8 warps, each warp owns a compact 8x4-pixel
// block of the tile.  A batch of up to 256 sorted Gaussian records is staged into shared memory by
// the CTA with three 16-byte cp.async gathers per record into a DOUBLE buffer: batch b+1 lands
// while batch b is blended, the ids of batch b+2 are already in a register, and one barrier per
// batch both publishes the new buffer and retires the old one.  Then every warp
//   1. culls the batch against ITS pixel block: lane k tests whether record k's alpha >= 1/255 ellipse
//      (threshold tau from preprocess) meets the block, exactly -> ballot -> bit mask of candidates;
//   2. walks only the set bits, evaluating all 32 pixels for that record with warp-uniform smem
//      broadcasts (LDS.128), and leaves as soon as all of its 32 pixels have saturated.
// Culling is conservative: a skipped (pixel, record) pair is one the reference would `continue`
// past (forward.cu:420-429), so T, n_contrib, colour and depth are unchanged.  The position of the
// last contributor is tracked from the list index, not by counting, so skipping does not shift it.
//
// Arithmetic is pinned to the reference's sm_100a SASS (nvcc default -fmad=true) with explicit
// intrinsics, because alpha decides threshold tests (1/255, T<1e-4) that flip whole contributions:
//   power = fma(fma(dx, dx*a, dy*(dy*c)), -0.5, -(dy*(dx*b)));  alpha = min(0.99, o*expf_exact(power, ek));
//   C = fma(T, alpha*col, C);  D = fma(T, alpha*depth, D);  acc = fma(T, alpha, acc).
#include "common.cuh"
#include "kernels.h"

namespace brs {

namespace {

constexpr int BLEND_THREADS = TILE_X * TILE_Y; // 256
constexpr int BATCH = 256;

__global__ void __launch_bounds__(BLEND_THREADS) blend_forward_kernel(BlendFwdArgs a)
{
	__shared__ StagedRecord s_rec[2][BATCH]; // geo {x, y, tau, -} | con {a, b, c, opacity} | col {r, g, b, depth}

	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t tile_x = blockIdx.x, tile_y = blockIdx.y;
	const uint32_t bx = tile_x * TILE_X + (warp & 1) * 8; // this warp's 8x4 pixel block
	const uint32_t by = tile_y * TILE_Y + (warp >> 1) * 4;
	const uint32_t px = bx + (lane & 7), py = by + (lane >> 3);
	const bool inside = px < (uint32_t)a.W && py < (uint32_t)a.H;
	const uint32_t pix_id = (uint32_t)a.W * py + px;
	const float pixfx = (float)px, pixfy = (float)py;
	const float wx0 = (float)bx, wx1 = (float)(bx + 7), wy0 = (float)by, wy1 = (float)(by + 3);

	const ExpConsts ek = {a.exp_c_scale, a.exp_c_252};
	const uint2 range = __ldg(a.ranges + tile_y * a.grid_x + tile_x);
	const int n = (int)(range.y - range.x);
	const uint32_t* list = a.point_list + range.x;

	bool done = !inside;
	float T = 1.0f;
	uint32_t last_contributor = 0;
	float C0 = 0.f, C1 = 0.f, C2 = 0.f, D = 0.f, acc = 0.000001f;

	// prologue: batch 0 in flight, ids of batch 1 in a register
	if ((int)tid < n)
		stage_record_async(&s_rec[0][tid], a.records, __ldg(list + tid));
	uint32_t id_next = (BATCH + (int)tid < n) ? __ldg(list + BATCH + tid) : 0u;

	for (int base = 0, buf = 0; base < n; base += BATCH, buf ^= 1) {
		cp_async_wait_all();
		bool warp_done = __all_sync(0xffffffffu, done);
		(void)warp_done;
		// one barrier per batch: buffer `buf` is complete and visible, buffer `buf ^ 1` is no longer read
		if (__syncthreads_and(warp_done))
			break;
		if (base + BATCH + (int)tid < n)
			stage_record_async(&s_rec[buf ^ 1][tid], a.records, id_next);
		id_next = (base + 2 * BATCH + (int)tid < n) ? __ldg(list + base + 2 * BATCH + tid) : 0u;

		const StagedRecord* rec = s_rec[buf];
		const int cnt = min(BATCH, n - base);
		for (int c0 = 0; c0 < cnt; c0 += 32) {
			// a warp leaves as soon as all of its 32 pixels have saturated (checked once per chunk: after the
			// exact culling ~97 % of the evaluated records contribute, so a per-record vote does not pay)
			if (__all_sync(0xffffffffu, done)) {
				warp_done = true;
				break;
			}
			const int e = c0 + (int)lane;
			bool hit = false;
			if (e < cnt)
				hit = block_may_contribute(rec[e].geo, rec[e].con, wx0, wx1, wy0, wy1);
			uint32_t mask = __ballot_sync(0xffffffffu, hit);
			while (mask) {
				const int j = __ffs(mask) - 1;
				mask &= mask - 1;
				const StagedRecord* r = rec + (c0 + j);
				const float2 gxy = *reinterpret_cast<const float2*>(&r->geo);
				const float4 con = r->con;
				const float4 col = r->col;
				const float dx = gxy.x - pixfx;
				const float dy = gxy.y - pixfy;
				const float t1 = __fmul_rn(dy, __fmul_rn(dy, con.z));
				const float s = __fmaf_rn(dx, __fmul_rn(dx, con.x), t1);
				const float t3 = __fmul_rn(dy, __fmul_rn(dx, con.y));
				const float power = __fmaf_rn(s, -0.5f, -t3);
				const float alpha = fminf(0.99f, __fmul_rn(con.w, expf_exact(power, ek)));
				bool ok = !done && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
				const float test_T = __fmul_rn(T, 1.0f - alpha);
				if (ok && test_T < 0.0001f) {
					done = true;
					ok = false;
				}
				if (ok) {
					C0 = __fmaf_rn(T, __fmul_rn(alpha, col.x), C0);
					C1 = __fmaf_rn(T, __fmul_rn(alpha, col.y), C1);
					C2 = __fmaf_rn(T, __fmul_rn(alpha, col.z), C2);
					D = __fmaf_rn(T, __fmul_rn(alpha, col.w), D);
					acc = __fmaf_rn(T, alpha, acc);
					T = test_T;
					last_contributor = (uint32_t)(base + c0 + j + 1);
				}
			}
		}
	}
	cp_async_wait_all();

	if (inside) {
		const size_t plane = (size_t)a.W * a.H;
		a.final_T[pix_id] = T;
		a.n_contrib[pix_id] = last_contributor;
		a.out_color[pix_id] = __fmaf_rn(__ldg(a.bg + 0), T, C0);
		a.out_color[plane + pix_id] = __fmaf_rn(__ldg(a.bg + 1), T, C1);
		a.out_color[2 * plane + pix_id] = __fmaf_rn(__ldg(a.bg + 2), T, C2);
		a.out_depth[pix_id] = (acc > 0.5f) ? __fdiv_rn(D, acc) : 0.0f;
	}
}

} // namespace

cudaError_t launch_blend_forward(const BlendFwdArgs& args, cudaStream_t stream)
{
	BlendFwdArgs a = args;
	const ExpConsts ek = exp_consts();
	a.exp_c_scale = ek.c_scale;
	a.exp_c_252 = ek.c_252;
	if (a.W <= 0 || a.H <= 0)
		return cudaSuccess;
	dim3 grid(a.grid_x, a.grid_y, 1);
	blend_forward_kernel<<<grid, BLEND_THREADS, 0, stream>>>(a);
	count_launch();
	return cudaGetLastError();
}

} // namespace brs
