// Minimal stand-in for the subset of GLM (g-truc/glm, header-only, MIT) that the reference
// rasterizer uses. TEST INFRASTRUCTURE ONLY: it exists so that the reference's own CUDA sources
// (which expect an un-vendored `third_party/glm` submodule, see the reference's setup.py:29) can
// be compiled unmodified into oracle/_ref/.  Nothing in the product path includes this file.
//
// GLM is absent from /root/reference (no .gitmodules in the snapshot, so the pinned commit is
// unknown; the upstream Inria rasterizer pins g-truc/glm 0.9.9.x).  What is restated here is
// GLM's published arithmetic for these operators, keeping its expression order, because with
// nvcc's default -fmad=true the order decides which products are fused:
//   * vec ops are component-wise, `v / s` is a true division per component,
//   * dot(a,b)   = (a*b).x + (a*b).y + (a*b).z,   length(v) = sqrt(dot(v,v)),
//   * mat3 is column-major (m[c][r]), mat3(9 scalars) fills column by column,
//   * (A*B)[c][r] = A[0][r]*B[c][0] + A[1][r]*B[c][1] + A[2][r]*B[c][2],
//   * s*M scales each column, transpose swaps indices, max(v,s) = (v<s)?s:v per component.
#pragma once
#include <cuda_runtime.h>
#include <cmath>

#define GLM_SHIM_FN __host__ __device__ inline

namespace glm {

struct vec3 {
	float x, y, z;
	GLM_SHIM_FN vec3() : x(0.f), y(0.f), z(0.f) {}
	GLM_SHIM_FN explicit vec3(float s) : x(s), y(s), z(s) {}
	template <typename A, typename B, typename C>
	GLM_SHIM_FN vec3(A a, B b, C c) : x(static_cast<float>(a)), y(static_cast<float>(b)), z(static_cast<float>(c)) {}
	GLM_SHIM_FN float& operator[](int i) { return (&x)[i]; }
	GLM_SHIM_FN const float& operator[](int i) const { return (&x)[i]; }
	GLM_SHIM_FN vec3& operator+=(const vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
	GLM_SHIM_FN vec3& operator+=(float s) { x += s; y += s; z += s; return *this; }
	GLM_SHIM_FN vec3& operator-=(const vec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
	GLM_SHIM_FN vec3& operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
	GLM_SHIM_FN vec3& operator*=(const vec3& o) { x *= o.x; y *= o.y; z *= o.z; return *this; }
	GLM_SHIM_FN vec3& operator/=(float s) { x /= s; y /= s; z /= s; return *this; }
};

struct vec4 {
	float x, y, z, w;
	GLM_SHIM_FN vec4() : x(0.f), y(0.f), z(0.f), w(0.f) {}
	GLM_SHIM_FN explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
	template <typename A, typename B, typename C, typename D>
	GLM_SHIM_FN vec4(A a, B b, C c, D d) : x(static_cast<float>(a)), y(static_cast<float>(b)), z(static_cast<float>(c)), w(static_cast<float>(d)) {}
	GLM_SHIM_FN float& operator[](int i) { return (&x)[i]; }
	GLM_SHIM_FN const float& operator[](int i) const { return (&x)[i]; }
};

GLM_SHIM_FN vec3 operator-(const vec3& v) { return vec3(-v.x, -v.y, -v.z); }
GLM_SHIM_FN vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
GLM_SHIM_FN vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
GLM_SHIM_FN vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
GLM_SHIM_FN vec3 operator+(const vec3& v, float s) { return vec3(v.x + s, v.y + s, v.z + s); }
GLM_SHIM_FN vec3 operator+(float s, const vec3& v) { return vec3(s + v.x, s + v.y, s + v.z); }
GLM_SHIM_FN vec3 operator-(const vec3& v, float s) { return vec3(v.x - s, v.y - s, v.z - s); }
GLM_SHIM_FN vec3 operator*(const vec3& v, float s) { return vec3(v.x * s, v.y * s, v.z * s); }
GLM_SHIM_FN vec3 operator*(float s, const vec3& v) { return vec3(s * v.x, s * v.y, s * v.z); }
GLM_SHIM_FN vec3 operator/(const vec3& v, float s) { return vec3(v.x / s, v.y / s, v.z / s); }

GLM_SHIM_FN float dot(const vec3& a, const vec3& b) { vec3 tmp(a * b); return tmp.x + tmp.y + tmp.z; }
GLM_SHIM_FN float length(const vec3& v) { return sqrtf(dot(v, v)); }
GLM_SHIM_FN vec3 max(const vec3& v, float s) { return vec3((v.x < s) ? s : v.x, (v.y < s) ? s : v.y, (v.z < s) ? s : v.z); }

struct mat3 {
	vec3 value[3];
	GLM_SHIM_FN mat3() { value[0] = vec3(1, 0, 0); value[1] = vec3(0, 1, 0); value[2] = vec3(0, 0, 1); }
	GLM_SHIM_FN explicit mat3(float s) { value[0] = vec3(s, 0, 0); value[1] = vec3(0, s, 0); value[2] = vec3(0, 0, s); }
	template <typename X1, typename Y1, typename Z1, typename X2, typename Y2, typename Z2, typename X3, typename Y3, typename Z3>
	GLM_SHIM_FN mat3(X1 x1, Y1 y1, Z1 z1, X2 x2, Y2 y2, Z2 z2, X3 x3, Y3 y3, Z3 z3)
	{
		value[0] = vec3(x1, y1, z1);
		value[1] = vec3(x2, y2, z2);
		value[2] = vec3(x3, y3, z3);
	}
	GLM_SHIM_FN mat3(const vec3& c0, const vec3& c1, const vec3& c2) { value[0] = c0; value[1] = c1; value[2] = c2; }
	GLM_SHIM_FN vec3& operator[](int i) { return value[i]; }
	GLM_SHIM_FN const vec3& operator[](int i) const { return value[i]; }
};

GLM_SHIM_FN mat3 operator*(const mat3& m1, const mat3& m2)
{
	const float SrcA00 = m1[0][0], SrcA01 = m1[0][1], SrcA02 = m1[0][2];
	const float SrcA10 = m1[1][0], SrcA11 = m1[1][1], SrcA12 = m1[1][2];
	const float SrcA20 = m1[2][0], SrcA21 = m1[2][1], SrcA22 = m1[2][2];
	const float SrcB00 = m2[0][0], SrcB01 = m2[0][1], SrcB02 = m2[0][2];
	const float SrcB10 = m2[1][0], SrcB11 = m2[1][1], SrcB12 = m2[1][2];
	const float SrcB20 = m2[2][0], SrcB21 = m2[2][1], SrcB22 = m2[2][2];
	mat3 Result(0.f);
	Result[0][0] = SrcA00 * SrcB00 + SrcA10 * SrcB01 + SrcA20 * SrcB02;
	Result[0][1] = SrcA01 * SrcB00 + SrcA11 * SrcB01 + SrcA21 * SrcB02;
	Result[0][2] = SrcA02 * SrcB00 + SrcA12 * SrcB01 + SrcA22 * SrcB02;
	Result[1][0] = SrcA00 * SrcB10 + SrcA10 * SrcB11 + SrcA20 * SrcB12;
	Result[1][1] = SrcA01 * SrcB10 + SrcA11 * SrcB11 + SrcA21 * SrcB12;
	Result[1][2] = SrcA02 * SrcB10 + SrcA12 * SrcB11 + SrcA22 * SrcB12;
	Result[2][0] = SrcA00 * SrcB20 + SrcA10 * SrcB21 + SrcA20 * SrcB22;
	Result[2][1] = SrcA01 * SrcB20 + SrcA11 * SrcB21 + SrcA21 * SrcB22;
	Result[2][2] = SrcA02 * SrcB20 + SrcA12 * SrcB21 + SrcA22 * SrcB22;
	return Result;
}

GLM_SHIM_FN mat3 operator*(float s, const mat3& m) { return mat3(m[0] * s, m[1] * s, m[2] * s); }
GLM_SHIM_FN mat3 operator*(const mat3& m, float s) { return mat3(m[0] * s, m[1] * s, m[2] * s); }

GLM_SHIM_FN mat3 transpose(const mat3& m)
{
	mat3 Result(0.f);
	Result[0][0] = m[0][0]; Result[0][1] = m[1][0]; Result[0][2] = m[2][0];
	Result[1][0] = m[0][1]; Result[1][1] = m[1][1]; Result[1][2] = m[2][1];
	Result[2][0] = m[0][2]; Result[2][1] = m[1][2]; Result[2][2] = m[2][2];
	return Result;
}

} // namespace glm
