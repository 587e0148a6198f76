// Forward alpha-blend for sm_100a  <- reference renderCUDA (cuda_rasterizer/forward.cu:341-471).
//
// One CTA per 16x16 tile (tile ids are parity outputs), 4 warps, each warp owns a compact 8x8-pixel
// block of the tile and each LANE TWO PIXELS of it — (x, y) and (x, y + 4) — so that the arithmetic
// runs on Blackwell's packed fp32 instructions (FFMA2 / FMUL2 / FADD2: one issue slot for both
// pixels; the kernel is issue-slot bound, see profiles/) and every warp-uniform cost of a record
// (shared-memory broadcasts, loop control) is paid once per 64 pixels.
// A batch of up to 256 sorted Gaussian records is staged into shared memory by the CTA with three
// 16-byte cp.async gathers per record into a DOUBLE buffer: batch b+1 lands while batch b is
// blended, the ids of batch b+2 are already in registers, and one barrier per batch both publishes
// the new buffer and retires the old one.  Then every warp
//   1. culls the batch against ITS pixel block: lane k tests whether record k's alpha >= 1/255 ellipse
//      (threshold tau from preprocess) meets the block, exactly -> ballot -> bit mask of candidates;
//   2. walks only the set bits, evaluating all 64 pixels for that record with warp-uniform smem
//      broadcasts (LDS.128), and leaves as soon as all of its pixels have saturated.
// Culling is conservative: a skipped (pixel, record) pair is one the reference would `continue`
// past (forward.cu:420-429), so T, n_contrib, colour and depth are unchanged.  The position of the
// last contributor is tracked from the list index, not by counting, so skipping does not shift it.
//
// Arithmetic is pinned to the reference's sm_100a SASS (nvcc default -fmad=true) with explicit
// operations, because alpha decides threshold tests (1/255, T<1e-4) that flip whole contributions:
//   power = fma(fma(dx, dx*a, dy*(dy*c)), -0.5, -(dy*(dx*b)));  alpha = min(0.99, o*expf_exact(power));
//   C = fma(T, alpha*col, C);  D = fma(T, alpha*depth, D);  acc = fma(T, alpha, acc).
// The packed form works on ndy = py - gy, the exact negation of dy: (-dy)*((-dy)*c) = dy*(dy*c) and
// (-dy)*(dx*b) = -(dy*(dx*b)) bit for bit, which supplies the negated operand without an instruction.
// A pixel that does not take a record keeps its state through alpha := 0 (fma(T, 0*col, C) == C for
// finite colours), so both pixels of a lane share one straight-line update.
#include "common.cuh"
#include "kernels.h"

namespace brs {

namespace {

constexpr int BLEND_THREADS = 128; // 4 warps x (8x8 pixels), two pixels per lane
constexpr int BATCH = 256;
constexpr int PER_THREAD = BATCH / BLEND_THREADS;

__global__ void __launch_bounds__(BLEND_THREADS) blend_forward_kernel(BlendFwdArgs a)
{
	__shared__ StagedRecord s_rec[2][BATCH]; // geo {x, y, tau, -} | con {a, b, c, opacity} | col {r, g, b, depth}

	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t tile_x = blockIdx.x, tile_y = blockIdx.y;
	const uint32_t bx = tile_x * TILE_X + (warp & 1) * 8; // this warp's 8x8 pixel block
	const uint32_t by = tile_y * TILE_Y + (warp >> 1) * 8;
	const uint32_t px = bx + (lane & 7), py0 = by + (lane >> 3), py1 = py0 + 4;
	const bool inside0 = px < (uint32_t)a.W && py0 < (uint32_t)a.H;
	const bool inside1 = px < (uint32_t)a.W && py1 < (uint32_t)a.H;
	const float pixfx = (float)px;
	const f32x2 pixfy = pk2((float)py0, (float)py1);
	const float wx0 = (float)bx, wx1 = (float)(bx + 7), wy0 = (float)by, wy1 = (float)(by + 7);

	const ExpConsts ek = {a.exp_c_scale, a.exp_c_252};
	const uint32_t view = blockIdx.z; // a stack of views: tile rows [view * grid_y, (view + 1) * grid_y) of `ranges`
	const uint2 range = __ldg(a.ranges + (view * a.grid_y + tile_y) * a.grid_x + tile_x);
	const int n = (int)(range.y - range.x);
	const uint32_t* list = a.point_list + range.x;

	bool done0 = !inside0, done1 = !inside1;
	F2 T = {1.0f, 1.0f}, T2 = {1.0f, 1.0f};
	F2 live = {done0 ? 0.f : 1.f, done1 ? 0.f : 1.f};
	uint32_t last0 = 0, last1 = 0;
	F2 C0 = {0.f, 0.f}, C1 = C0, C2 = C0, D = C0, acc = {0.000001f, 0.000001f};

	// prologue: batch 0 in flight, ids of batch 1 in registers
	uint32_t id_next[PER_THREAD];
#pragma unroll
	for (int k = 0; k < PER_THREAD; k++) {
		const int s = k * BLEND_THREADS + (int)tid;
		if (s < n)
			stage_record_async(&s_rec[0][s], a.records, __ldg(list + s));
		id_next[k] = (BATCH + s < n) ? __ldg(list + BATCH + s) : 0u;
	}

	for (int base = 0, buf = 0; base < n; base += BATCH, buf ^= 1) {
		cp_async_wait_all();
		bool warp_done = __all_sync(0xffffffffu, done0 && done1);
		// one barrier per batch: buffer `buf` is complete and visible, buffer `buf ^ 1` is no longer read
		if (__syncthreads_and(warp_done))
			break;
#pragma unroll
		for (int k = 0; k < PER_THREAD; k++) {
			const int s = k * BLEND_THREADS + (int)tid;
			if (base + BATCH + s < n)
				stage_record_async(&s_rec[buf ^ 1][s], a.records, id_next[k]);
			id_next[k] = (base + 2 * BATCH + s < n) ? __ldg(list + base + 2 * BATCH + s) : 0u;
		}

		const StagedRecord* rec = s_rec[buf];
		const int cnt = min(BATCH, n - base);
		for (int c0 = 0; c0 < cnt; c0 += 32) {
			// a warp leaves as soon as all of its pixels have saturated (checked once per chunk: after the
			// exact culling ~97 % of the evaluated records contribute, so a per-record vote does not pay)
			if (__all_sync(0xffffffffu, done0 && done1))
				break;
			const int e = c0 + (int)lane;
			bool hit = false;
			if (e < cnt)
				hit = block_may_contribute(rec[e].geo, rec[e].con, wx0, wx1, wy0, wy1);
			uint32_t mask = __ballot_sync(0xffffffffu, hit);
			// one record for the warp's 64 pixels; the transmittance goes Tin -> Tout so that two calls per loop trip
			// ping-pong between two register pairs (a single pair costs a register-move pair per record)
			auto blend_one = [&](const F2& Tin, F2& Tout) {
				const int j = __ffs(mask) - 1;
				mask &= mask - 1;
				const StagedRecord* r = rec + (c0 + j);
				const float2 gxy = *reinterpret_cast<const float2*>(&r->geo);
				const float4 con = r->con;
				const float4 col = r->col;
				const float dx = gxy.x - pixfx;
				const float dxa = __fmul_rn(dx, con.x);
				const float dxb = __fmul_rn(dx, con.y);
				const f32x2 ndy = sub2(pixfy, bc2(gxy.y));
				const f32x2 t1 = mul2(ndy, mul2(ndy, bc2(con.z)));
				const f32x2 s = fma2(bc2(dx), bc2(dxa), t1);
				const f32x2 nt3 = mul2(ndy, bc2(dxb));
				const f32x2 power = fma2(s, bc2(-0.5f), nt3);
				// a saturated pixel multiplies its alpha by live = 0 (exact: o*G*1 == o*G), so the loop
				// carries no per-pixel `done` test.  A non-finite opacity turns that product into NaN and
				// fminf(0.99, NaN) = 0.99, but a finished pixel cannot be revived by it: it stopped with
				// T * (1 - alpha) < 1e-4 for some alpha <= 0.99, i.e. T < 1e-2, so T * (1 - 0.99) < 1e-4 and the
				// saturation branch below discards the record again (measured: masking by predicate instead
				// costs 1 % of the kernel; tests/test_gpu_parity.py::test_non_finite_inputs_... covers the case)
				const f32x2 oG = mul2(mul2(bc2(con.w), expf_exact2(power, ek)), pk2(live));
				const float alpha0 = fminf(0.99f, lo2(oG)), alpha1 = fminf(0.99f, hi2(oG));
				const bool ok0 = !(lo2(power) > 0.0f) && !(alpha0 < 1.0f / 255.0f);
				const bool ok1 = !(hi2(power) > 0.0f) && !(alpha1 < 1.0f / 255.0f);
				// a pixel that skips the record blends alpha = 0: T * (1 - 0) == T and fma(T, 0 * col, C) == C
				F2 al = {ok0 ? alpha0 : 0.f, ok1 ? alpha1 : 0.f};
				const f32x2 Tp = pk2(Tin);
				const F2 test_T = unpk2(mul2(Tp, sub2(bc2(1.0f), pk2(al))));
				const uint32_t pos = (uint32_t)(base + c0 + j + 1);
				bool take0 = ok0, take1 = ok1, keep0 = false, keep1 = false;
				if (__any_sync(0xffffffffu, fminf(test_T.lo, test_T.hi) < 0.0001f)) { // warp-uniform: no reconvergence point
					// rare (once per pixel): the record that would saturate the pixel is not blended
					// (reference forward.cu:431-436) and the pixel stops
					const bool s0 = test_T.lo < 0.0001f, s1 = test_T.hi < 0.0001f;
					done0 |= s0;
					done1 |= s1;
					live.lo = s0 ? 0.f : live.lo;
					live.hi = s1 ? 0.f : live.hi;
					al.lo = s0 ? 0.f : al.lo;
					al.hi = s1 ? 0.f : al.hi;
					take0 = take0 && !s0;
					take1 = take1 && !s1;
					keep0 = s0;
					keep1 = s1;
				}
				// state is updated in place after the rare branch (no old / new pair live across it)
				last0 = take0 ? pos : last0;
				last1 = take1 ? pos : last1;
				Tout.lo = keep0 ? Tin.lo : test_T.lo;
				Tout.hi = keep1 ? Tin.hi : test_T.hi;
				const f32x2 alp = pk2(al);
				C0 = unpk2(fma2(mul2(alp, bc2(col.x)), Tp, pk2(C0)));
				C1 = unpk2(fma2(mul2(alp, bc2(col.y)), Tp, pk2(C1)));
				C2 = unpk2(fma2(mul2(alp, bc2(col.z)), Tp, pk2(C2)));
				D = unpk2(fma2(mul2(alp, bc2(col.w)), Tp, pk2(D)));
				acc = unpk2(fma2(alp, Tp, pk2(acc)));
			};
			if (mask) do {
				blend_one(T, T2);
				if (!mask) {
					T = T2;
					break;
				}
				blend_one(T2, T);
			} while (mask);
		}
	}
	cp_async_wait_all();

	const size_t plane = (size_t)a.W * a.H;
	const float bg0 = __ldg(a.bg + 0), bg1 = __ldg(a.bg + 1), bg2 = __ldg(a.bg + 2);
	float* out_color = a.out_color + view * (NUM_CHANNELS * plane);
	const size_t pix_base = view * plane;
	if (inside0) {
		const uint32_t pix_id = (uint32_t)a.W * py0 + px;
		if (a.final_T != nullptr) {
			a.final_T[pix_base + pix_id] = T.lo;
			a.n_contrib[pix_base + pix_id] = last0;
		}
		out_color[pix_id] = __fmaf_rn(bg0, T.lo, C0.lo);
		out_color[plane + pix_id] = __fmaf_rn(bg1, T.lo, C1.lo);
		out_color[2 * plane + pix_id] = __fmaf_rn(bg2, T.lo, C2.lo);
		a.out_depth[pix_base + pix_id] = (acc.lo > 0.5f) ? __fdiv_rn(D.lo, acc.lo) : 0.0f;
	}
	if (inside1) {
		const uint32_t pix_id = (uint32_t)a.W * py1 + px;
		if (a.final_T != nullptr) {
			a.final_T[pix_base + pix_id] = T.hi;
			a.n_contrib[pix_base + pix_id] = last1;
		}
		out_color[pix_id] = __fmaf_rn(bg0, T.hi, C0.hi);
		out_color[plane + pix_id] = __fmaf_rn(bg1, T.hi, C1.hi);
		out_color[2 * plane + pix_id] = __fmaf_rn(bg2, T.hi, C2.hi);
		a.out_depth[pix_base + pix_id] = (acc.hi > 0.5f) ? __fdiv_rn(D.hi, acc.hi) : 0.0f;
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
	dim3 grid(a.grid_x, a.grid_y, a.views > 1 ? a.views : 1);
	blend_forward_kernel<<<grid, BLEND_THREADS, 0, stream>>>(a);
	count_launch();
	return cudaGetLastError();
}

} // namespace brs
