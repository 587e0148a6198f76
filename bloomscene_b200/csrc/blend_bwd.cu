// Backward alpha-blend for sm_100a  <- reference renderCUDA (cuda_rasterizer/backward.cu:399-586).
//
// Same tile / warp / pixel mapping, packed fp32 arithmetic and shared-memory staging as the forward
// kernel (blend_fwd.cu): 4 warps per 16x16 tile, an 8x8-pixel block per warp, TWO PIXELS PER LANE —
// (x, y) and (x, y + 4) — replayed back-to-front.  Differences from the reference that change cost,
// not results:
//   * the tile starts at max(n_contrib) over its pixels and each warp skips records behind the
//     max(n_contrib) of ITS 64 pixels (the reference walks them with a per-thread `continue`);
//   * per-warp exact culling with the records' alpha>=1/255 ellipses (see blend_fwd.cu);
//   * the per-pixel chain is branch-free: a pixel that does not take a record runs it with alpha := 0,
//     which leaves T (T * rcp(1) == T), the behind-colour recurrence (fma(0, x, beta) == beta) and the
//     staged weights (0) exactly as skipping would;
//   * the per-Gaussian gradient sums are NOT reduced pair by pair.  All nine are fixed linear
//     functions of two per-pixel scalars,
//         w = G * dL_dalpha            u = alpha * T
//         dL_dopacity = S(w)           dL_dcolor_c = S(u * dL_dpixel_c)
//         dL_dmean2D, dL_dconic  <-  S(w dx), S(w dy), S(w dx dx), S(w dx dy), S(w dy dy)
//     so the pixel loop only stores the lane's (w, u) pairs of a contributing record into a
//     warp-private staging slot (2 STS.64).  Every 8 staged records the warp "flushes": lane (r, q)
//     takes record r and pixel rows q and q + 4 of the warp's block — which sit side by side in the
//     staged float2s, so the row sums of both run packed — the 4 lanes of a record are combined with
//     an 8-shuffle transposing tree, and the record leaves as one red.global.add.v2.f32 per lane
//     (+1) on a 48-byte accumulator row, against the reference's 9 x 32-lane atomics per pair
//     (backward.cu:537,574-583);
//   * constant factors (0.5*W, 0.5*H, -0.5) are applied once per Gaussian in the preprocess
//     backward instead of once per pair.
// Parity quirks kept: the alpha derivative ignores the min(0.99,.) clamp, T is recovered by
// division from T_final (backward.cu:513-517), and depth carries no gradient (all uses commented
// out in the reference, backward.cu:443-554) unless the caller opts in (template DEPTH).  The alpha
// tests use the forward's pinned arithmetic so exactly the same pairs contribute.
#include "common.cuh"
#include "kernels.h"

namespace brs {

namespace {

constexpr int BLEND_THREADS = 128; // 4 warps x (8x8 pixels), two pixels per lane
constexpr int BLEND_WARPS = BLEND_THREADS / 32;
constexpr int BATCH = 128;
constexpr int GROUP = 8; // contributing records staged per warp between flushes (8 records x 4 lanes = 32 lanes)
// Staged record (136 floats): W2 rows | batch slot | U2 rows, a row = 8 float2, a float2 = the values of the
// pixels (x, row) and (x, row + 4).  W2 rows sit at float offsets {0, 16, 36, 52}, the slot index in the gap at 32,
// U2 rows at 68 + {0, 16, 36, 52}; records are 136 apart (= 8 mod 32): the pixel loop's STS.64 (half-warp = two
// rows) and the flush's LDS.128 (lane (r, q): record r, row q) are both bank-conflict-free.  s_dpx keeps planes
// of 4 rows x (8 float2 + 4 pad).
__device__ __forceinline__ constexpr int stage_row(int row) { return row * 16 + (row >> 1) * 4; }
constexpr int STAGE_U = 68;
constexpr int STAGE_SLOT = 32;
constexpr int STAGE_STRIDE = 136;
constexpr int ROW_PITCH = 20;
constexpr int PLANE = 4 * ROW_PITCH; // 80 floats

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
// record r and the pixel rows q (lo halves) and q + 4 (hi halves) of the warp's 8x8 block.
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
		const float* srow = stage + r * STAGE_STRIDE;
		const uint32_t idx = (uint32_t)__float_as_int(srow[STAGE_SLOT]);
		const float2 g = *reinterpret_cast<const float2*>(&rec[idx].geo);
		const float4 con = rec[idx].con;
		id = s_id[idx];
		// phase 1: the moments of w along x (both rows packed); w's registers are dead before u is fetched
		const float dyA = g.y - (byf + (float)q), dyB = g.y - (byf + (float)(q + 4));
		const float dx0 = g.x - bxf;
		f32x2 S0, Sx, Sxx;
		{
			const ulonglong2* wp = reinterpret_cast<const ulonglong2*>(srow + stage_row(q));
			{ // k = 0 initialises the sums (no additions of a zero accumulator)
				const ulonglong2 a = wp[0];
				const f32x2 dxa = bc2(dx0), dxb = bc2(dx0 - 1.0f);
				const f32x2 ta = mul2(a.x, dxa), tb = mul2(a.y, dxb);
				S0 = add2(a.x, a.y);
				Sx = add2(ta, tb);
				Sxx = fma2(tb, dxb, mul2(ta, dxa));
			}
#pragma unroll
			for (int k = 1; k < 4; k++) {
				const ulonglong2 a = wp[k];
				const f32x2 dxa = bc2(dx0 - (float)(2 * k)), dxb = bc2(dx0 - (float)(2 * k + 1));
				const f32x2 ta = mul2(a.x, dxa), tb = mul2(a.y, dxb);
				S0 = add2(S0, add2(a.x, a.y));
				Sx = add2(Sx, add2(ta, tb));
				Sxx = fma2(ta, dxa, Sxx);
				Sxx = fma2(tb, dxb, Sxx);
			}
		}
		// phase 2: the colour sums of u
		constexpr int NCH = DEPTH ? 4 : 3; // with DEPTH the 4th plane of s_dpx holds gD = dL_dD
		float c[4] = {0.f, 0.f, 0.f, 0.f};
		{
			const ulonglong2* up = reinterpret_cast<const ulonglong2*>(srow + STAGE_U + stage_row(q));
			const ulonglong2 u0 = up[0], u1 = up[1], u2 = up[2], u3 = up[3];
#pragma unroll
			for (int ch = 0; ch < NCH; ch++) {
				const ulonglong2* dp = reinterpret_cast<const ulonglong2*>(s_dpx + ch * PLANE + q * ROW_PITCH);
				const ulonglong2 d0 = dp[0], d1 = dp[1], d2 = dp[2], d3 = dp[3];
				f32x2 s = mul2(u0.x, d0.x);
				s = fma2(u0.y, d0.y, s);
				s = fma2(u1.x, d1.x, s);
				s = fma2(u1.y, d1.y, s);
				s = fma2(u2.x, d2.x, s);
				s = fma2(u2.y, d2.y, s);
				s = fma2(u3.x, d3.x, s);
				s = fma2(u3.y, d3.y, s);
				c[ch] = lo2(s) + hi2(s);
			}
		}
		const float S0a = lo2(S0), S0b = hi2(S0), Sxa = lo2(Sx), Sxb = hi2(Sx);
		const float S0t = S0a + S0b, Sxt = Sxa + Sxb, Sxxt = lo2(Sxx) + hi2(Sxx);
		const float Sya = dyA * S0a, Syb = dyB * S0b;
		const float Sy = Sya + Syb;
		const float Sxy = fmaf(dyA, Sxa, dyB * Sxb);
		const float Syy = fmaf(dyA, Sya, dyB * Syb);
		const float o = con.w;
		// dL_dG * G = o * w;  dG_ddelx = -G (dx a + dy b),  dG_ddely = -G (dy c + dx b)
		v[0] = -o * (con.x * Sxt + con.y * Sy); // x 0.5*W later
		v[1] = -o * (con.z * Sy + con.y * Sxt); // x 0.5*H later
		v[2] = o * Sxxt;                         // x -0.5 later
		v[3] = o * Sxy;
		v[4] = o * Syy;
		v[5] = S0t;
		v[6] = c[0];
		v[7] = c[1];
		v[8] = c[2];
		v[9] = c[3];
	}
	// combine the 4 lanes of a record (q = 0..3): 9 -> 5 -> 3 values per lane, 8 shuffles.
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
	__shared__ __align__(16) float s_dpx[BLEND_WARPS][(DEPTH ? 4 : 3) * PLANE];
	__shared__ uint32_t s_max[BLEND_WARPS];

	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t tile_x = blockIdx.x, tile_y = blockIdx.y;
	const uint32_t bx = tile_x * TILE_X + (warp & 1) * 8; // this warp's 8x8 pixel block
	const uint32_t by = tile_y * TILE_Y + (warp >> 1) * 8;
	const uint32_t px = bx + (lane & 7), py0 = by + (lane >> 3), py1 = py0 + 4;
	const bool inside0 = px < (uint32_t)a.W && py0 < (uint32_t)a.H;
	const bool inside1 = px < (uint32_t)a.W && py1 < (uint32_t)a.H;
	const uint32_t pix0 = (uint32_t)a.W * py0 + px, pix1 = (uint32_t)a.W * py1 + px;
	const float pixfx = (float)px;
	const f32x2 pixfy = pk2((float)py0, (float)py1);
	const float wx0 = (float)bx, wx1 = (float)(bx + 7), wy0 = (float)by, wy1 = (float)(by + 7);

	const ExpConsts ek = {a.exp_c_scale, a.exp_c_252};
	const uint2 range = __ldg(a.ranges + tile_y * a.grid_x + tile_x);

	const float Tf0 = inside0 ? __ldg(a.final_T + pix0) : 0.0f, Tf1 = inside1 ? __ldg(a.final_T + pix1) : 0.0f;
	F2 T = {Tf0, Tf1};
	const int last0 = inside0 ? (int)__ldg(a.n_contrib + pix0) : 0, last1 = inside1 ? (int)__ldg(a.n_contrib + pix1) : 0;

	// per-pixel upstream gradients; the flush reads them from shared memory as (row, row + 4) pairs
	const size_t plane = (size_t)a.W * a.H;
	float2 d0 = {0.f, 0.f}, d1 = d0, d2 = d0;
	if (inside0) {
		d0.x = __ldg(a.dL_dpixels + pix0);
		d1.x = __ldg(a.dL_dpixels + plane + pix0);
		d2.x = __ldg(a.dL_dpixels + 2 * plane + pix0);
	}
	if (inside1) {
		d0.y = __ldg(a.dL_dpixels + pix1);
		d1.y = __ldg(a.dL_dpixels + plane + pix1);
		d2.y = __ldg(a.dL_dpixels + 2 * plane + pix1);
	}
	const uint32_t pair_off = (lane >> 3) * ROW_PITCH + (lane & 7) * 2; // this lane's float2 inside a plane
	float* const dpx_rows = s_dpx[warp];
	*reinterpret_cast<float2*>(dpx_rows + pair_off) = d0;
	*reinterpret_cast<float2*>(dpx_rows + PLANE + pair_off) = d1;
	*reinterpret_cast<float2*>(dpx_rows + 2 * PLANE + pair_off) = d2;
	float2 gD = {0.f, 0.f}, gA = {0.f, 0.f};
	if (DEPTH) {
		if (inside0) {
			const float depth = __ldg(a.out_depth + pix0);
			if (depth > 0.f) { // the forward's acc > 0.5 gate (view-space z > 0.2, so D / acc > 0 exactly when it passed)
				const float g_over_acc = __fdividef(__ldg(a.dL_ddepth + pix0), 0.000001f + (1.0f - Tf0));
				gD.x = g_over_acc;
				gA.x = -g_over_acc * depth;
			}
		}
		if (inside1) {
			const float depth = __ldg(a.out_depth + pix1);
			if (depth > 0.f) {
				const float g_over_acc = __fdividef(__ldg(a.dL_ddepth + pix1), 0.000001f + (1.0f - Tf1));
				gD.y = g_over_acc;
				gA.y = -g_over_acc * depth;
			}
		}
		*reinterpret_cast<float2*>(dpx_rows + 3 * PLANE + pair_off) = gD;
	}
	const f32x2 dpx0 = pk2(d0.x, d0.y), dpx1 = pk2(d1.x, d1.y), dpx2 = pk2(d2.x, d2.y);
	const f32x2 gD2 = pk2(gD.x, gD.y), gA2 = pk2(gA.x, gA.y);

	// records at list positions >= max(n_contrib) are skipped by every pixel of the warp / tile
	int warp_max = max(last0, last1);
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

	F2 beta = {0.f, 0.f}, last_alpha = {0.f, 0.f}, last_cd = {0.f, 0.f};
	const float bg0 = __ldg(a.bg + 0), bg1 = __ldg(a.bg + 1), bg2 = __ldg(a.bg + 2);
	float bgd0 = 0.f, bgd1 = 0.f;
	bgd0 += bg0 * d0.x; bgd0 += bg1 * d1.x; bgd0 += bg2 * d2.x;
	bgd1 += bg0 * d0.y; bgd1 += bg1 * d1.y; bgd1 += bg2 * d2.y;
	const f32x2 neg_Tf_bg = pk2(-Tf0 * bgd0, -Tf1 * bgd1);

	float* const stage = s_stage[warp];
	int staged = 0;
	const uint32_t stage_off = stage_row(lane >> 3) + (lane & 7) * 2;
	float* st = stage + stage_off; // this lane's W2 float2 in the next free staging slot

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

		// slot idx is list position n-1-base-idx; it is in front of a pixel's last contributor
		// iff idx > first_live (and of this warp's iff idx > warp_first_live)
		const int first_live0 = n - 1 - base - last0, first_live1 = n - 1 - base - last1;
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
				const float dxa = __fmul_rn(dx, con.x);
				const float dxb = __fmul_rn(dx, con.y);
				const f32x2 ndy = sub2(pixfy, bc2(g.y)); // -(dy): see blend_fwd.cu
				const f32x2 t1 = mul2(ndy, mul2(ndy, bc2(con.z)));
				const f32x2 s = fma2(bc2(dx), bc2(dxa), t1);
				const f32x2 nt3 = mul2(ndy, bc2(dxb));
				const f32x2 power = fma2(s, bc2(-0.5f), nt3);
				const f32x2 G = expf_exact2(power, ek);
				const f32x2 oG = mul2(bc2(con.w), G);
				const float alpha0 = fminf(0.99f, lo2(oG)), alpha1 = fminf(0.99f, hi2(oG));
				const bool ok0 = (idx > first_live0) && !(lo2(power) > 0.0f) && !(alpha0 < 1.0f / 255.0f);
				const bool ok1 = (idx > first_live1) && !(hi2(power) > 0.0f) && !(alpha1 < 1.0f / 255.0f);
				if (!__any_sync(0xffffffffu, ok0 || ok1))
					continue;

				const float4 col = r->col;
				// the behind-colour recurrence only needs loop-carried values: formed right after the skip test and
				// before this record's alpha is selected, so that it updates beta in place and the previous record's
				// alpha is dead by the time the new one lands in its registers
				const f32x2 b_old = pk2(beta);
				const f32x2 bn = fma2(pk2(last_alpha), sub2(pk2(last_cd), b_old), b_old);
				// a pixel that skips the record runs the chain with alpha = 0 (see the header)
				const f32x2 al = pk2(ok0 ? alpha0 : 0.f, ok1 ? alpha1 : 0.f);
				// 1 - alpha >= 0.01 (alpha is clamped to 0.99): one MUFU.RCP (<= 1 ulp, exact for 1.0) replaces the
				// reference's two IEEE divisions; gradients are tolerance-checked (rel-L2 <= 1e-4)
				const f32x2 om = sub2(bc2(1.0f), al);
				const f32x2 rcp = pk2(rcp_approx(lo2(om)), rcp_approx(hi2(om)));
				const f32x2 Tn = mul2(pk2(T), rcp);
				const f32x2 u_ = mul2(al, Tn); // dchannel_dcolor

				// reference backward.cu:519-533 keeps, per channel, the colour accumulated BEHIND this record
				// (accum_rec) and sums (c - accum_rec) * dL_dpixel over the channels.  Only that sum is
				// needed, so the recurrence runs on its projection: beta = accum_rec . dL_dpixel and
				// cd = colour . dL_dpixel are scalars.
				f32x2 cd = fma2(bc2(col.z), dpx2, fma2(bc2(col.y), dpx1, mul2(bc2(col.x), dpx0)));
				if (DEPTH)
					cd = add2(fma2(bc2(col.w), gD2, cd), gA2);
				f32x2 dL_dalpha = mul2(sub2(cd, bn), Tn);
				dL_dalpha = fma2(neg_Tf_bg, rcp, dL_dalpha);
				const f32x2 wv = mul2(G, dL_dalpha);
				T = unpk2(Tn);
				beta = unpk2(bn);
				last_cd = unpk2(cd);
				last_alpha = unpk2(al);
				float2 w2, u2;
				w2.x = ok0 ? lo2(wv) : 0.f;
				w2.y = ok1 ? hi2(wv) : 0.f;
				u2.x = lo2(u_);
				u2.y = hi2(u_);
				*reinterpret_cast<float2*>(st) = w2;
				*reinterpret_cast<float2*>(st + STAGE_U) = u2;
				if (lane == 0)
					st[STAGE_SLOT] = __int_as_float(idx); // the record's batch slot (lane 0's stage_off is 0: st == slot base)
				st += STAGE_STRIDE;
				if (++staged == GROUP) {
					flush_group<DEPTH>(stage, GROUP, rec, s_id, dpx_rows, wx0, wy0, lane, a.accum);
					staged = 0;
					st = stage + stage_off;
				}
			}
		}
		// staged slots refer to this batch's shared records: flush before they are overwritten
		if (staged) {
			flush_group<DEPTH>(stage, staged, rec, s_id, dpx_rows, wx0, wy0, lane, a.accum);
			staged = 0;
			st = stage + stage_off;
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
