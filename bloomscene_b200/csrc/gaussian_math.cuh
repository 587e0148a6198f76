// Per-Gaussian maths shared by the forward and backward preprocess kernels.  The backward
// RECOMPUTES cov3D and the SH colour (for the `clamped` flags) instead of storing them in the
// forward, so both directions must run the exact same device functions.
#pragma once
#include "common.cuh"

namespace brs {

// reference forward.cu:118-152 (quaternion is NOT normalised: forward.cu:127)
__device__ __forceinline__ void compute_cov3d(const v3 scale, float mod, const float4 rot, float* cov3D)
{
	mat3 S = make_mat3(1.0f, 0.0f, 0.0f, 0.0f, 1.0f, 0.0f, 0.0f, 0.0f, 1.0f);
	S.c[0].x = mod * scale.x;
	S.c[1].y = mod * scale.y;
	S.c[2].z = mod * scale.z;

	float r = rot.x;
	float x = rot.y;
	float y = rot.z;
	float z = rot.w;

	mat3 R = make_mat3(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
	                   2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
	                   2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));

	mat3 M = mul(S, R);
	mat3 Sigma = mul(transpose(M), M);

	cov3D[0] = Sigma.c[0].x;
	cov3D[1] = Sigma.c[0].y;
	cov3D[2] = Sigma.c[0].z;
	cov3D[3] = Sigma.c[1].y;
	cov3D[4] = Sigma.c[1].z;
	cov3D[5] = Sigma.c[2].z;
}

// A Gaussian's SH row as it sits in shared memory (float4 pieces).  operator[] with a compile-time
// index is one LDS.128 (merged by the compiler with its neighbours) plus a register pick, so the row
// is pulled into registers a few coefficients at a time instead of living there as 48 floats.
struct ShRowView {
	const float4* row;
	__device__ __forceinline__ float operator[](int i) const
	{
		const float4 v = row[i >> 2];
		const int c = i & 3;
		return c == 0 ? v.x : (c == 1 ? v.y : (c == 2 ? v.z : v.w));
	}
};

// reference forward.cu:20-71.  `sh` indexes this Gaussian's coefficients (3 floats per coefficient,
// dense): a `const float*` or a ShRowView.  Returns the un-clamped colour; the caller clamps.
template <class SH>
__device__ __forceinline__ v3 eval_sh(int deg, const v3 pos, const v3 campos, const SH sh)
{
	v3 dir = pos - campos;
	dir = dir / length3(dir);

#define SHC(k) make_v3(sh[3 * (k)], sh[3 * (k) + 1], sh[3 * (k) + 2])
	v3 result = SH_C0 * SHC(0);

	if (deg > 0) {
		float x = dir.x;
		float y = dir.y;
		float z = dir.z;
		result = result - SH_C1 * y * SHC(1) + SH_C1 * z * SHC(2) - SH_C1 * x * SHC(3);

		if (deg > 1) {
			float xx = x * x, yy = y * y, zz = z * z;
			float xy = x * y, yz = y * z, xz = x * z;
			result = result + SH_C2[0] * xy * SHC(4) + SH_C2[1] * yz * SHC(5) +
			         SH_C2[2] * (2.0f * zz - xx - yy) * SHC(6) + SH_C2[3] * xz * SHC(7) +
			         SH_C2[4] * (xx - yy) * SHC(8);

			if (deg > 2) {
				result = result + SH_C3[0] * y * (3.0f * xx - yy) * SHC(9) + SH_C3[1] * xy * z * SHC(10) +
				         SH_C3[2] * y * (4.0f * zz - xx - yy) * SHC(11) +
				         SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * SHC(12) +
				         SH_C3[4] * x * (4.0f * zz - xx - yy) * SHC(13) + SH_C3[5] * z * (xx - yy) * SHC(14) +
				         SH_C3[6] * x * (xx - 3.0f * yy) * SHC(15);
			}
		}
	}
#undef SHC
	result = result + 0.5f;
	return result;
}

} // namespace brs
