// Backward of the per-Gaussian preprocessing for sm_100a: ONE kernel for the reference's two
//   computeCov2DCUDA (cuda_rasterizer/backward.cu:144-274)  and
//   preprocessCUDA   (backward.cu:346-396) with computeColorFromSH (20-139) / computeCov3D (278-341).
//
// Fusing them removes the dL_dcov3D / dL_dmeans round trip through HBM, and this kernel also
//   * consumes the blend-backward accumulator (48 B per Gaussian) and applies the per-Gaussian
//     constants (0.5*W, 0.5*H, -0.5) that the reference multiplies into every pair;
//   * writes EVERY element of every gradient tensor (zeros for Gaussians with radii <= 0), so the
//     caller needs no zero-fill: the reference zero-fills (108 + 12 M) bytes per Gaussian first
//     (rasterize_points.cu:154-162);
//   * or, in ACCUMULATE mode (brs_grads.accumulate, used by the view-sharded step), adds the parameter
//     gradients of the visible Gaussians straight into the caller's gradient bucket with float
//     reductions (red.global.add, 16-byte vectors for the SH rows), which replaces one fresh
//     (44 + 12 M)-byte-per-Gaussian tensor set plus a read-modify-write pass over the bucket per view
//     by a single reduction of the visible rows — and is safe when views on different streams add
//     into the same bucket concurrently;
//   * recomputes cov3D and the SH colour sign (`clamped`) with the forward's device functions
//     instead of reading stored copies;
//   * stages SH rows (in) and dL_dsh rows (out) through shared memory so that global traffic is
//     coalesced 128-bit even though a Gaussian's row is 192 bytes.
#include "common.cuh"
#include "gaussian_math.cuh"
#include "kernels.h"

namespace brs {

namespace {

constexpr int PB_THREADS = 128;

// 16-byte vector reduction (sm_90+): four float adds in one L2 atomic, no return value.
__device__ __forceinline__ void red_add_v4(float* addr, float x, float y, float z, float w)
{
	asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(x), "f"(y), "f"(z), "f"(w)
	             : "memory");
}
constexpr int FACT_PITCH = 19; // 16 basis factors + dL_dRGB, odd pitch -> conflict-free

// reference auxiliary.h:107-117
__device__ __forceinline__ float3 dnormvdv(float3 v, float3 dv)
{
	float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
	float invsum32 = 1.0f / sqrt(sum2 * sum2 * sum2);

	float3 r;
	r.x = ((+sum2 - v.x * v.x) * dv.x - v.y * v.x * dv.y - v.z * v.x * dv.z) * invsum32;
	r.y = (-v.x * v.y * dv.x + (sum2 - v.y * v.y) * dv.y - v.z * v.y * dv.z) * invsum32;
	r.z = (-v.x * v.z * dv.x - v.y * v.z * dv.y + (sum2 - v.z * v.z) * dv.z) * invsum32;
	return r;
}

// reference backward.cu:20-139.  `sh` = this Gaussian's coefficients (constant indices only),
// fact[0..15] receive dRGB/dsh_k, returns dL/dmean contribution through the view direction.
template <class SH>
__device__ __forceinline__ float3 sh_backward(int deg, const v3 pos, const v3 campos, const SH sh,
                                              const v3 dL_dRGB, float* fact)
{
	v3 dir_orig = pos - campos;
	v3 dir = dir_orig / length3(dir_orig);

#define SHC(k) make_v3(sh[3 * (k)], sh[3 * (k) + 1], sh[3 * (k) + 2])
	v3 dRGBdx = make_v3(0, 0, 0);
	v3 dRGBdy = make_v3(0, 0, 0);
	v3 dRGBdz = make_v3(0, 0, 0);
	float x = dir.x;
	float y = dir.y;
	float z = dir.z;

#pragma unroll
	for (int k = 0; k < 16; k++)
		fact[k] = 0.f;
	fact[0] = SH_C0;
	if (deg > 0) {
		fact[1] = -SH_C1 * y;
		fact[2] = SH_C1 * z;
		fact[3] = -SH_C1 * x;

		dRGBdx = -SH_C1 * SHC(3);
		dRGBdy = -SH_C1 * SHC(1);
		dRGBdz = SH_C1 * SHC(2);

		if (deg > 1) {
			float xx = x * x, yy = y * y, zz = z * z;
			float xy = x * y, yz = y * z, xz = x * z;

			fact[4] = SH_C2[0] * xy;
			fact[5] = SH_C2[1] * yz;
			fact[6] = SH_C2[2] * (2.f * zz - xx - yy);
			fact[7] = SH_C2[3] * xz;
			fact[8] = SH_C2[4] * (xx - yy);

			dRGBdx += SH_C2[0] * y * SHC(4) + SH_C2[2] * 2.f * -x * SHC(6) + SH_C2[3] * z * SHC(7) +
			          SH_C2[4] * 2.f * x * SHC(8);
			dRGBdy += SH_C2[0] * x * SHC(4) + SH_C2[1] * z * SHC(5) + SH_C2[2] * 2.f * -y * SHC(6) +
			          SH_C2[4] * 2.f * -y * SHC(8);
			dRGBdz += SH_C2[1] * y * SHC(5) + SH_C2[2] * 2.f * 2.f * z * SHC(6) + SH_C2[3] * x * SHC(7);

			if (deg > 2) {
				fact[9] = SH_C3[0] * y * (3.f * xx - yy);
				fact[10] = SH_C3[1] * xy * z;
				fact[11] = SH_C3[2] * y * (4.f * zz - xx - yy);
				fact[12] = SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
				fact[13] = SH_C3[4] * x * (4.f * zz - xx - yy);
				fact[14] = SH_C3[5] * z * (xx - yy);
				fact[15] = SH_C3[6] * x * (xx - 3.f * yy);

				dRGBdx += (SH_C3[0] * SHC(9) * 3.f * 2.f * xy + SH_C3[1] * SHC(10) * yz +
				           SH_C3[2] * SHC(11) * -2.f * xy + SH_C3[3] * SHC(12) * -3.f * 2.f * xz +
				           SH_C3[4] * SHC(13) * (-3.f * xx + 4.f * zz - yy) + SH_C3[5] * SHC(14) * 2.f * xz +
				           SH_C3[6] * SHC(15) * 3.f * (xx - yy));

				dRGBdy += (SH_C3[0] * SHC(9) * 3.f * (xx - yy) + SH_C3[1] * SHC(10) * xz +
				           SH_C3[2] * SHC(11) * (-3.f * yy + 4.f * zz - xx) + SH_C3[3] * SHC(12) * -3.f * 2.f * yz +
				           SH_C3[4] * SHC(13) * -2.f * xy + SH_C3[5] * SHC(14) * -2.f * yz +
				           SH_C3[6] * SHC(15) * -3.f * 2.f * xy);

				dRGBdz += (SH_C3[1] * SHC(10) * xy + SH_C3[2] * SHC(11) * 4.f * 2.f * yz +
				           SH_C3[3] * SHC(12) * 3.f * (2.f * zz - xx - yy) + SH_C3[4] * SHC(13) * 4.f * 2.f * xz +
				           SH_C3[5] * SHC(14) * (xx - yy));
			}
		}
	}
#undef SHC

	v3 dL_ddir = make_v3(dot3(dRGBdx, dL_dRGB), dot3(dRGBdy, dL_dRGB), dot3(dRGBdz, dL_dRGB));
	return dnormvdv(float3{dir_orig.x, dir_orig.y, dir_orig.z}, float3{dL_ddir.x, dL_ddir.y, dL_ddir.z});
}

// reference backward.cu:278-341 (no quaternion-normalisation Jacobian: the caller normalises)
__device__ __forceinline__ void cov3d_backward(const v3 scale, float mod, const float4 rot, const float* dL_dcov3D,
                                               float3& dL_dscale, float4& dL_drot)
{
	float r = rot.x;
	float x = rot.y;
	float y = rot.z;
	float z = rot.w;

	mat3 R = make_mat3(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
	                   2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
	                   2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));

	mat3 S = make_mat3(1.0f, 0.0f, 0.0f, 0.0f, 1.0f, 0.0f, 0.0f, 0.0f, 1.0f);
	v3 s = mod * scale;
	S.c[0].x = s.x;
	S.c[1].y = s.y;
	S.c[2].z = s.z;

	mat3 M = mul(S, R);

	mat3 dL_dSigma = make_mat3(dL_dcov3D[0], 0.5f * dL_dcov3D[1], 0.5f * dL_dcov3D[2], 0.5f * dL_dcov3D[1],
	                           dL_dcov3D[3], 0.5f * dL_dcov3D[4], 0.5f * dL_dcov3D[2], 0.5f * dL_dcov3D[4],
	                           dL_dcov3D[5]);

	mat3 dL_dM = mul(scale_cols(2.0f, M), dL_dSigma);

	mat3 Rt = transpose(R);
	mat3 dL_dMt = transpose(dL_dM);

	dL_dscale.x = dot3(Rt.c[0], dL_dMt.c[0]);
	dL_dscale.y = dot3(Rt.c[1], dL_dMt.c[1]);
	dL_dscale.z = dot3(Rt.c[2], dL_dMt.c[2]);

	dL_dMt.c[0] *= s.x;
	dL_dMt.c[1] *= s.y;
	dL_dMt.c[2] *= s.z;

#define MT(c_, r_) dL_dMt.at(c_, r_)
	dL_drot.x = 2 * z * (MT(0, 1) - MT(1, 0)) + 2 * y * (MT(2, 0) - MT(0, 2)) + 2 * x * (MT(1, 2) - MT(2, 1));
	dL_drot.y = 2 * y * (MT(1, 0) + MT(0, 1)) + 2 * z * (MT(2, 0) + MT(0, 2)) + 2 * r * (MT(1, 2) - MT(2, 1)) -
	            4 * x * (MT(2, 2) + MT(1, 1));
	dL_drot.z = 2 * x * (MT(1, 0) + MT(0, 1)) + 2 * r * (MT(2, 0) - MT(0, 2)) + 2 * z * (MT(1, 2) + MT(2, 1)) -
	            4 * y * (MT(2, 2) + MT(0, 0));
	dL_drot.w = 2 * r * (MT(0, 1) - MT(1, 0)) + 2 * x * (MT(2, 0) + MT(0, 2)) + 2 * y * (MT(1, 2) + MT(2, 1)) -
	            4 * z * (MT(1, 1) + MT(0, 0));
#undef MT
}

template <bool VEC>
__global__ void __launch_bounds__(PB_THREADS, 6) preprocess_backward_kernel(PreprocessBwdArgs a)
{
	extern __shared__ float4 s_dyn[]; // SH rows in, then basis factors out
	__shared__ float s_cam[36];

	{
		const int t = threadIdx.x;
		if (t < 16)
			s_cam[t] = __ldg(a.viewmatrix + t);
		else if (t < 32)
			s_cam[t] = __ldg(a.projmatrix + t - 16);
		else if (t < 35 && a.campos != nullptr) // only the SH chain reads it; colors_precomp callers may pass none
			s_cam[t] = __ldg(a.campos + t - 32);
	}

	const int block_first = blockIdx.x * PB_THREADS;
	const int idx = block_first + threadIdx.x;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const bool in_range = idx < a.P;
	const bool visible = in_range && (__ldg(a.radii + idx) > 0);
	const uint32_t vis_mask = __ballot_sync(0xffffffffu, visible);
	__syncthreads(); // camera block
	const float* view = s_cam;
	const float* proj = s_cam + 16;

	const int row_f = 3 * a.M;
	const int warp_first = block_first + 32 * warp;
	// ---- every warp stages the SH rows of ITS visible Gaussians (no block barrier between the phases) ----
	if (a.shs != nullptr) {
		if (VEC) {
			const float4* src = reinterpret_cast<const float4*>(a.shs) + (size_t)warp_first * 12;
			float4* dst = s_dyn + 32 * warp * 13;
#pragma unroll
			for (int k = 0; k < 12; k++) {
				const int f = lane + 32 * k;
				const int row = f / 12, col = f - row * 12;
				if ((vis_mask >> row) & 1u)
					dst[row * 13 + col] = ldg_stream_f4(src + f);
			}
		} else {
			const int pitch = row_f | 1;
			float* s_sh = reinterpret_cast<float*>(s_dyn) + 32 * warp * pitch;
			const float* src = a.shs + (size_t)warp_first * row_f;
			const int total = 32 * row_f;
			for (int f = lane; f < total; f += 32) {
				const int row = f / row_f, col = f - row * row_f;
				if ((vis_mask >> row) & 1u)
					s_sh[row * pitch + col] = __ldg(src + f);
			}
		}
		__syncwarp();
	}

	// basis factors + dL_dRGB of this Gaussian go straight to their own shared-memory row (not through
	// registers); rows of Gaussians that do not run the SH chain must read as zero
	float* const s_fact = reinterpret_cast<float*>(s_dyn) + a.fact_offset;
	float* const fact = s_fact + threadIdx.x * FACT_PITCH;
	const bool sh_rows = a.dL_dsh != nullptr && a.M > 0;
	if (sh_rows && !(visible && a.shs != nullptr)) {
#pragma unroll
		for (int k = 0; k < FACT_PITCH; k++)
			fact[k] = 0.f;
	}
	v3 dL_dRGB = make_v3(0.f, 0.f, 0.f);

	float3 o_mean2D = {0.f, 0.f, 0.f}, o_color = {0.f, 0.f, 0.f}, o_mean3D = {0.f, 0.f, 0.f};
	float3 o_scale = {0.f, 0.f, 0.f};
	float4 o_rot = {0.f, 0.f, 0.f, 0.f};
	float o_opacity = 0.f;
	float o_cov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};

	if (visible) {
		const float4* ap = reinterpret_cast<const float4*>(a.accum + (size_t)idx * ACCUM_STRIDE);
		const float4 a0 = __ldg(ap), a1 = __ldg(ap + 1), a2 = __ldg(ap + 2);
		// per-Gaussian constants the reference applies per pair (backward.cu:473-474,574-580)
		const float ddelx_dx = 0.5 * a.W;
		const float ddely_dy = 0.5 * a.H;
		o_mean2D = {a0.x * ddelx_dx, a0.y * ddely_dy, 0.f};
		const float3 dL_dconic = {-0.5f * a0.z, -0.5f * a0.w, -0.5f * a1.x};
		o_opacity = a1.y;
		o_color = {a1.z, a1.w, a2.x};

		const float* mp = a.means3D + 3 * (size_t)idx;
		const float3 mean = {__ldg(mp), __ldg(mp + 1), __ldg(mp + 2)};

		// ---- cov3D (recomputed exactly as in the forward) ----
		float cov3D[6];
		v3 scale = make_v3(0.f, 0.f, 0.f);
		float4 rot = make_float4(0.f, 0.f, 0.f, 0.f);
		if (a.cov3D_precomp != nullptr) {
#pragma unroll
			for (int i = 0; i < 6; i++)
				cov3D[i] = __ldg(a.cov3D_precomp + 6 * (size_t)idx + i);
		} else {
			const float* sp = a.scales + 3 * (size_t)idx;
			scale = make_v3(__ldg(sp), __ldg(sp + 1), __ldg(sp + 2));
			const float* rp = a.rotations + 4 * (size_t)idx;
			rot = make_float4(__ldg(rp), __ldg(rp + 1), __ldg(rp + 2), __ldg(rp + 3));
			compute_cov3d(scale, a.scale_modifier, rot, cov3D);
		}

		// ---- computeCov2DCUDA (backward.cu:164-273) ----
		float3 t = transform_point_4x3(mean, view);
		const float limx = 1.3f * a.tan_fovx;
		const float limy = 1.3f * a.tan_fovy;
		const float txtz = t.x / t.z;
		const float tytz = t.y / t.z;
		t.x = min(limx, max(-limx, txtz)) * t.z;
		t.y = min(limy, max(-limy, tytz)) * t.z;

		const float x_grad_mul = txtz < -limx || txtz > limx ? 0 : 1;
		const float y_grad_mul = tytz < -limy || tytz > limy ? 0 : 1;

		const float h_x = a.focal_x, h_y = a.focal_y;
		mat3 J = make_mat3(h_x / t.z, 0.0f, -(h_x * t.x) / (t.z * t.z), 0.0f, h_y / t.z, -(h_y * t.y) / (t.z * t.z), 0,
		                   0, 0);
		mat3 W = make_mat3(view[0], view[4], view[8], view[1], view[5], view[9], view[2], view[6], view[10]);
		mat3 Vrk = make_mat3(cov3D[0], cov3D[1], cov3D[2], cov3D[1], cov3D[3], cov3D[4], cov3D[2], cov3D[4], cov3D[5]);
		mat3 T = mul(W, J);
		mat3 cov2D = mul(mul(transpose(T), transpose(Vrk)), T);

		float ca = cov2D.c[0].x += 0.3f;
		float cb = cov2D.c[0].y;
		float cc = cov2D.c[1].y += 0.3f;

		float denom = ca * cc - cb * cb;
		float dL_da = 0, dL_db = 0, dL_dc = 0;
		float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);

#define TT(c_, r_) T.at(c_, r_)
#define VV(c_, r_) Vrk.at(c_, r_)
#define WW(c_, r_) W.at(c_, r_)
		if (denom2inv != 0) {
			dL_da = denom2inv * (-cc * cc * dL_dconic.x + 2 * cb * cc * dL_dconic.y + (denom - ca * cc) * dL_dconic.z);
			dL_dc = denom2inv * (-ca * ca * dL_dconic.z + 2 * ca * cb * dL_dconic.y + (denom - ca * cc) * dL_dconic.x);
			dL_db = denom2inv * 2 * (cb * cc * dL_dconic.x - (denom + 2 * cb * cb) * dL_dconic.y + ca * cb * dL_dconic.z);

			o_cov[0] = (TT(0, 0) * TT(0, 0) * dL_da + TT(0, 0) * TT(1, 0) * dL_db + TT(1, 0) * TT(1, 0) * dL_dc);
			o_cov[3] = (TT(0, 1) * TT(0, 1) * dL_da + TT(0, 1) * TT(1, 1) * dL_db + TT(1, 1) * TT(1, 1) * dL_dc);
			o_cov[5] = (TT(0, 2) * TT(0, 2) * dL_da + TT(0, 2) * TT(1, 2) * dL_db + TT(1, 2) * TT(1, 2) * dL_dc);

			o_cov[1] = 2 * TT(0, 0) * TT(0, 1) * dL_da + (TT(0, 0) * TT(1, 1) + TT(0, 1) * TT(1, 0)) * dL_db +
			           2 * TT(1, 0) * TT(1, 1) * dL_dc;
			o_cov[2] = 2 * TT(0, 0) * TT(0, 2) * dL_da + (TT(0, 0) * TT(1, 2) + TT(0, 2) * TT(1, 0)) * dL_db +
			           2 * TT(1, 0) * TT(1, 2) * dL_dc;
			o_cov[4] = 2 * TT(0, 2) * TT(0, 1) * dL_da + (TT(0, 1) * TT(1, 2) + TT(0, 2) * TT(1, 1)) * dL_db +
			           2 * TT(1, 1) * TT(1, 2) * dL_dc;
		}

		float dL_dT00 = 2 * (TT(0, 0) * VV(0, 0) + TT(0, 1) * VV(0, 1) + TT(0, 2) * VV(0, 2)) * dL_da +
		                (TT(1, 0) * VV(0, 0) + TT(1, 1) * VV(0, 1) + TT(1, 2) * VV(0, 2)) * dL_db;
		float dL_dT01 = 2 * (TT(0, 0) * VV(1, 0) + TT(0, 1) * VV(1, 1) + TT(0, 2) * VV(1, 2)) * dL_da +
		                (TT(1, 0) * VV(1, 0) + TT(1, 1) * VV(1, 1) + TT(1, 2) * VV(1, 2)) * dL_db;
		float dL_dT02 = 2 * (TT(0, 0) * VV(2, 0) + TT(0, 1) * VV(2, 1) + TT(0, 2) * VV(2, 2)) * dL_da +
		                (TT(1, 0) * VV(2, 0) + TT(1, 1) * VV(2, 1) + TT(1, 2) * VV(2, 2)) * dL_db;
		float dL_dT10 = 2 * (TT(1, 0) * VV(0, 0) + TT(1, 1) * VV(0, 1) + TT(1, 2) * VV(0, 2)) * dL_dc +
		                (TT(0, 0) * VV(0, 0) + TT(0, 1) * VV(0, 1) + TT(0, 2) * VV(0, 2)) * dL_db;
		float dL_dT11 = 2 * (TT(1, 0) * VV(1, 0) + TT(1, 1) * VV(1, 1) + TT(1, 2) * VV(1, 2)) * dL_dc +
		                (TT(0, 0) * VV(1, 0) + TT(0, 1) * VV(1, 1) + TT(0, 2) * VV(1, 2)) * dL_db;
		float dL_dT12 = 2 * (TT(1, 0) * VV(2, 0) + TT(1, 1) * VV(2, 1) + TT(1, 2) * VV(2, 2)) * dL_dc +
		                (TT(0, 0) * VV(2, 0) + TT(0, 1) * VV(2, 1) + TT(0, 2) * VV(2, 2)) * dL_db;

		float dL_dJ00 = WW(0, 0) * dL_dT00 + WW(0, 1) * dL_dT01 + WW(0, 2) * dL_dT02;
		float dL_dJ02 = WW(2, 0) * dL_dT00 + WW(2, 1) * dL_dT01 + WW(2, 2) * dL_dT02;
		float dL_dJ11 = WW(1, 0) * dL_dT10 + WW(1, 1) * dL_dT11 + WW(1, 2) * dL_dT12;
		float dL_dJ12 = WW(2, 0) * dL_dT10 + WW(2, 1) * dL_dT11 + WW(2, 2) * dL_dT12;
#undef TT
#undef VV
#undef WW

		float tz = 1.f / t.z;
		float tz2 = tz * tz;
		float tz3 = tz2 * tz;

		float dL_dtx = x_grad_mul * -h_x * tz2 * dL_dJ02;
		float dL_dty = y_grad_mul * -h_y * tz2 * dL_dJ12;
		float dL_dtz = -h_x * tz2 * dL_dJ00 - h_y * tz2 * dL_dJ11 + (2 * h_x * t.x) * tz3 * dL_dJ02 +
		               (2 * h_y * t.y) * tz3 * dL_dJ12;

		// mean gradient, part 1: through the covariance (reference overwrites dL_dmeans here)
		o_mean3D = transform_vec_4x3_transpose({dL_dtx, dL_dty, dL_dtz}, view);

		// ---- preprocessCUDA backward (backward.cu:370-395) ----
		float4 m_hom = transform_point_4x4(mean, proj);
		float m_w = 1.0f / (m_hom.w + 0.0000001f);
		float mul1 = (proj[0] * mean.x + proj[4] * mean.y + proj[8] * mean.z + proj[12]) * m_w * m_w;
		float mul2 = (proj[1] * mean.x + proj[5] * mean.y + proj[9] * mean.z + proj[13]) * m_w * m_w;
		float3 dm;
		dm.x = (proj[0] * m_w - proj[3] * mul1) * o_mean2D.x + (proj[1] * m_w - proj[3] * mul2) * o_mean2D.y;
		dm.y = (proj[4] * m_w - proj[7] * mul1) * o_mean2D.x + (proj[5] * m_w - proj[7] * mul2) * o_mean2D.y;
		dm.z = (proj[8] * m_w - proj[11] * mul1) * o_mean2D.x + (proj[9] * m_w - proj[11] * mul2) * o_mean2D.y;
		o_mean3D.x += dm.x;
		o_mean3D.y += dm.y;
		o_mean3D.z += dm.z;
		if (a.depth_gradient) {
			// opt-in depth gradient: z = p_view.z = view[2] x + view[6] y + view[10] z + view[14] (forward.cu:186)
			const float dz = a2.y;
			o_mean3D.x += view[2] * dz;
			o_mean3D.y += view[6] * dz;
			o_mean3D.z += view[10] * dz;
		}

		if (a.shs != nullptr) {
			const v3 pos = make_v3(mean.x, mean.y, mean.z);
			const v3 cam = make_v3(s_cam[32], s_cam[33], s_cam[34]);
			v3 rgb;
			float3 dmean_sh;
			if (VEC) {
				// coefficients are read from the staged row where they are used (conflict-free LDS.128)
				const ShRowView c{s_dyn + threadIdx.x * 13};
				rgb = eval_sh(a.D, pos, cam, c);
				dL_dRGB = make_v3(rgb.x < 0 ? 0.f : o_color.x, rgb.y < 0 ? 0.f : o_color.y, rgb.z < 0 ? 0.f : o_color.z);
				dmean_sh = sh_backward(a.D, pos, cam, c, dL_dRGB, fact);
			} else {
				const float* sh = reinterpret_cast<const float*>(s_dyn) + threadIdx.x * (row_f | 1);
				rgb = eval_sh(a.D, pos, cam, sh);
				dL_dRGB = make_v3(rgb.x < 0 ? 0.f : o_color.x, rgb.y < 0 ? 0.f : o_color.y, rgb.z < 0 ? 0.f : o_color.z);
				dmean_sh = sh_backward(a.D, pos, cam, sh, dL_dRGB, fact);
			}
			o_mean3D.x += dmean_sh.x;
			o_mean3D.y += dmean_sh.y;
			o_mean3D.z += dmean_sh.z;
			if (sh_rows) {
				fact[16] = dL_dRGB.x;
				fact[17] = dL_dRGB.y;
				fact[18] = dL_dRGB.z;
			}
		}

		if (a.scales != nullptr)
			cov3d_backward(scale, a.scale_modifier, rot, o_cov, o_scale, o_rot);
	}

	// ---- per-Gaussian outputs ----
	// plain mode: every row of every tensor is written (zeros for culled Gaussians);
	// accumulate mode: parameter gradients of VISIBLE Gaussians are added to what the buffers hold,
	// other rows are left alone; dL_dmeans2D is a per-view statistic and is always overwritten.
	if (in_range) {
		float* p;
		p = a.dL_dmeans2D + 3 * (size_t)idx;
		p[0] = o_mean2D.x; p[1] = o_mean2D.y; p[2] = 0.f;
		if (!a.accumulate) {
			p = a.dL_dcolors + 3 * (size_t)idx;
			p[0] = o_color.x; p[1] = o_color.y; p[2] = o_color.z;
			a.dL_dopacity[idx] = o_opacity;
			p = a.dL_dmeans3D + 3 * (size_t)idx;
			p[0] = o_mean3D.x; p[1] = o_mean3D.y; p[2] = o_mean3D.z;
			p = a.dL_dcov3D + 6 * (size_t)idx;
#pragma unroll
			for (int i = 0; i < 6; i++)
				p[i] = o_cov[i];
			p = a.dL_dscales + 3 * (size_t)idx;
			p[0] = o_scale.x; p[1] = o_scale.y; p[2] = o_scale.z;
			p = a.dL_drotations + 4 * (size_t)idx;
			p[0] = o_rot.x; p[1] = o_rot.y; p[2] = o_rot.z; p[3] = o_rot.w;
		} else if (visible) {
			// float atomics (RED, no return value): several views may be adding into the same bucket from
			// different streams at the same time
			// a NULL sink is a parameter the caller froze: its gradient is dropped
			if (a.dL_dmeans3D != nullptr) {
				float* p3 = a.dL_dmeans3D + 3 * (size_t)idx;
				atomicAdd(p3 + 0, o_mean3D.x);
				atomicAdd(p3 + 1, o_mean3D.y);
				atomicAdd(p3 + 2, o_mean3D.z);
			}
			if (a.dL_dopacity != nullptr)
				atomicAdd(a.dL_dopacity + idx, o_opacity);
			if (a.scales != nullptr) {
				if (a.dL_dscales != nullptr) {
					float* ps = a.dL_dscales + 3 * (size_t)idx;
					atomicAdd(ps + 0, o_scale.x);
					atomicAdd(ps + 1, o_scale.y);
					atomicAdd(ps + 2, o_scale.z);
				}
				if (a.dL_drotations != nullptr) {
					float* pr = a.dL_drotations + 4 * (size_t)idx;
					if ((reinterpret_cast<uintptr_t>(a.dL_drotations) & 15u) == 0) {
						red_add_v4(pr, o_rot.x, o_rot.y, o_rot.z, o_rot.w);
					} else { // a bucket slice that is not 16-byte aligned (odd P)
						atomicAdd(pr + 0, o_rot.x);
						atomicAdd(pr + 1, o_rot.y);
						atomicAdd(pr + 2, o_rot.z);
						atomicAdd(pr + 3, o_rot.w);
					}
				}
			} else if (a.dL_dcov3D != nullptr) {
				p = a.dL_dcov3D + 6 * (size_t)idx;
#pragma unroll
				for (int i = 0; i < 6; i++)
					atomicAdd(p + i, o_cov[i]);
			}
			if (a.shs == nullptr && a.dL_dcolors != nullptr) {
				p = a.dL_dcolors + 3 * (size_t)idx;
				atomicAdd(p + 0, o_color.x);
				atomicAdd(p + 1, o_color.y);
				atomicAdd(p + 2, o_color.z);
			}
		}
	}

	// ---- dL_dsh rows: dL_dsh[k] = fact[k] * dL_dRGB, written coalesced through shared memory ----
	if (sh_rows) {
		__syncwarp(); // the 32 rows of this warp in s_fact are complete
		const int rows = min(32, a.P - warp_first);
		const float* wfact = s_fact + 32 * warp * FACT_PITCH;
		if (VEC) {
			float4* dst = reinterpret_cast<float4*>(a.dL_dsh) + (size_t)warp_first * 12;
#pragma unroll
			for (int k = 0; k < 12; k++) {
				const int f = lane + 32 * k;
				const int row = f / 12, col = f - row * 12;
				if (row < rows && (!a.accumulate || ((vis_mask >> row) & 1u))) {
					const float* fr = wfact + row * FACT_PITCH;
					float o[4];
#pragma unroll
					for (int q = 0; q < 4; q++) {
						const int e = col * 4 + q;
						o[q] = fr[e / 3] * fr[16 + (e % 3)];
					}
					if (a.accumulate)
						red_add_v4(reinterpret_cast<float*>(dst + f), o[0], o[1], o[2], o[3]);
					else
						dst[f] = make_float4(o[0], o[1], o[2], o[3]);
				}
			}
		} else {
			float* dst = a.dL_dsh + (size_t)warp_first * row_f;
			const int total = rows * row_f;
			for (int f = lane; f < total; f += 32) {
				const int row = f / row_f, e = f - row * row_f;
				const float* fr = wfact + row * FACT_PITCH;
				const float val = fr[e / 3] * fr[16 + (e % 3)];
				if (!a.accumulate)
					dst[f] = val;
				else if ((vis_mask >> row) & 1u)
					atomicAdd(dst + f, val);
			}
		}
	}
}

} // namespace

cudaError_t launch_preprocess_backward(const PreprocessBwdArgs& a, cudaStream_t stream)
{
	if (a.P <= 0)
		return cudaSuccess;
	const int blocks = (a.P + PB_THREADS - 1) / PB_THREADS;
	// dynamic shared memory: [SH rows in | basis-factor rows out]
	size_t in = 0;
	bool vec = false;
	if (a.shs != nullptr) {
		vec = (a.M == 16) && ((reinterpret_cast<uintptr_t>(a.shs) & 15u) == 0) &&
		      ((reinterpret_cast<uintptr_t>(a.dL_dsh) & 15u) == 0);
		in = vec ? (size_t)PB_THREADS * 13 * sizeof(float4) : align_up((size_t)PB_THREADS * ((3 * a.M) | 1) * sizeof(float), 16);
	}
	PreprocessBwdArgs args = a;
	args.fact_offset = (int)(in / sizeof(float));
	const size_t smem = in + (size_t)PB_THREADS * FACT_PITCH * sizeof(float);
	if (vec)
		preprocess_backward_kernel<true><<<blocks, PB_THREADS, smem, stream>>>(args);
	else
		preprocess_backward_kernel<false><<<blocks, PB_THREADS, smem, stream>>>(args);
	count_launch();
	return cudaGetLastError();
}

} // namespace brs
