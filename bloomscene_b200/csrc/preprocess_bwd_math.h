// Per-Gaussian backward maths of the preprocess stage, derived from the forward model in matrix form
// (host + device: tests/test_bwd_math.py compiles this header with g++ and checks every function against
// central finite differences of a float64 forward model).
//
// Forward model (reference forward.cu:74-152,186-237; conventions of the reference kept):
//   t      = Wr p + t0                      view-space position; Wr_ij = view[i + 4 j]
//   u_x    = clamp(t_x / t_z, +-1.3 tan_fovx) t_z      (likewise u_y)
//   J      = [ fx/t_z   0   -fx u_x / t_z^2 ;  0   fy/t_z   -fy u_y / t_z^2 ]          (2 x 3)
//   T      = J Wr                                                                      (2 x 3, rows T0, T1)
//   Sigma  = Rq diag(s^2) Rq^T,  s = mod * scale,  Rq = rotation matrix of the (unnormalised) quaternion
//   C2     = T Sigma T^T + 0.3 I = [a b; b c],   conic = C2^-1 = [A B; B C]
// The blend backward hands over G = [gx gy; gy gz], the gradient with respect to the SYMMETRIC conic matrix
// (the reference's dL_dconic2D {x, y, w}, whose y is half the derivative with respect to the scalar B).
//
// Matrix calculus used below:
//   d(C2^-1) = -C2^-1 dC2 C2^-1            =>  H  := dL/dC2   = -conic G conic = -(adj G adj) / det^2
//   C2 = T Sigma T^T                       =>  dL/dSigma = T^T H T,   dL/dT = 2 H T Sigma
//   Sigma = sum_i s_i^2 r_i r_i^T (r_i = columns of Rq)
//                                          =>  dL/ds_i = 2 s_i r_i^T G3 r_i,   K := dL/dRq = 2 G3 Rq diag(s^2)
//   Rq quadratic in (r, x, y, z)           =>  dL/dq from the symmetric / antisymmetric parts of K
// The reference's parity quirks are kept on purpose: det^2 is regularised by 1e-7, the clamp of u makes
// the x / y gradients of t vanish but u's dependence on t_z is ignored, dL/dscale is the gradient with
// respect to s = mod * scale, and the quaternion is not normalised.
#pragma once

#if defined(__CUDACC__)
#define BRS_HD __host__ __device__ __forceinline__
#else
#define BRS_HD inline
#endif

namespace brs {
namespace bwdmath {

struct V3 {
	float x, y, z;
};
BRS_HD V3 vec3(float x, float y, float z) { return V3{x, y, z}; }
BRS_HD float dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
BRS_HD V3 axpby(float a, const V3& x, float b, const V3& y) { return V3{a * x.x + b * y.x, a * x.y + b * y.y, a * x.z + b * y.z}; }

// symmetric 3 x 3: xx xy xz yy yz zz
struct Sym3 {
	float xx, xy, xz, yy, yz, zz;
};
BRS_HD V3 mul(const Sym3& m, const V3& v)
{
	return V3{m.xx * v.x + m.xy * v.y + m.xz * v.z, m.xy * v.x + m.yy * v.y + m.yz * v.z, m.xz * v.x + m.yz * v.y + m.zz * v.z};
}

// Rotation matrix of the quaternion (r, x, y, z), columns c0 c1 c2 (reference forward.cu:127-137 builds its transpose).
struct Rot {
	V3 c0, c1, c2;
};
BRS_HD Rot rotation(float r, float x, float y, float z)
{
	Rot R;
	R.c0 = V3{1.f - 2.f * (y * y + z * z), 2.f * (x * y + r * z), 2.f * (x * z - r * y)};
	R.c1 = V3{2.f * (x * y - r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z + r * x)};
	R.c2 = V3{2.f * (x * z + r * y), 2.f * (y * z - r * x), 1.f - 2.f * (x * x + y * y)};
	return R;
}

// Sigma = sum_i s_i^2 r_i r_i^T
BRS_HD Sym3 covariance(const Rot& R, const V3& s)
{
	const float a = s.x * s.x, b = s.y * s.y, c = s.z * s.z;
	Sym3 S;
	S.xx = a * R.c0.x * R.c0.x + b * R.c1.x * R.c1.x + c * R.c2.x * R.c2.x;
	S.xy = a * R.c0.x * R.c0.y + b * R.c1.x * R.c1.y + c * R.c2.x * R.c2.y;
	S.xz = a * R.c0.x * R.c0.z + b * R.c1.x * R.c1.z + c * R.c2.x * R.c2.z;
	S.yy = a * R.c0.y * R.c0.y + b * R.c1.y * R.c1.y + c * R.c2.y * R.c2.y;
	S.yz = a * R.c0.y * R.c0.z + b * R.c1.y * R.c1.z + c * R.c2.y * R.c2.z;
	S.zz = a * R.c0.z * R.c0.z + b * R.c1.z * R.c1.z + c * R.c2.z * R.c2.z;
	return S;
}

// ---- projected covariance: conic gradient -> dL/dSigma and the part of dL/dmean that flows through J ----
struct ProjectionGrad {
	Sym3 dSigma; // MATRIX gradient (off-diagonal entries are per matrix element, i.e. half the reference's dL_dcov3D)
	V3 dmean;
};
// view = 16 floats (memory = transpose of the maths matrix), mean = world position, Sigma = 3D covariance,
// g = (gx, gy, gz) of the symmetric conic gradient.
BRS_HD ProjectionGrad projection_backward(const float* view, const V3& mean, const Sym3& Sigma, float fx, float fy,
                                          float tan_fovx, float tan_fovy, float gx, float gy, float gz)
{
	const V3 W0 = vec3(view[0], view[4], view[8]), W1 = vec3(view[1], view[5], view[9]), W2 = vec3(view[2], view[6], view[10]);
	const float tx = dot(W0, mean) + view[12], ty = dot(W1, mean) + view[13], tz = dot(W2, mean) + view[14];
	const float limx = 1.3f * tan_fovx, limy = 1.3f * tan_fovy;
	const float rx = tx / tz, ry = ty / tz;
	const bool free_x = !(rx < -limx || rx > limx), free_y = !(ry < -limy || ry > limy);
	const float ux = fminf(limx, fmaxf(-limx, rx)) * tz, uy = fminf(limy, fmaxf(-limy, ry)) * tz;
	const float iz = 1.f / tz, iz2 = iz * iz;
	const float j00 = fx * iz, j02 = -fx * ux * iz2, j11 = fy * iz, j12 = -fy * uy * iz2;
	const V3 T0 = axpby(j00, W0, j02, W2), T1 = axpby(j11, W1, j12, W2);
	const V3 ST0 = mul(Sigma, T0), ST1 = mul(Sigma, T1);
	const float a = dot(T0, ST0) + 0.3f, b = dot(T0, ST1), c = dot(T1, ST1) + 0.3f;
	const float det = a * c - b * b;
	const float k = 1.f / (det * det + 0.0000001f);
	// H = -(adj G adj) k,  adj = [c -b; -b a]
	const float p00 = gx * c - gy * b, p01 = gy * a - gx * b, p10 = gy * c - gz * b, p11 = gz * a - gy * b; // G adj
	const float ha = -k * (c * p00 - b * p10), hb = -k * (c * p01 - b * p11), hc = -k * (a * p11 - b * p01);
	ProjectionGrad out;
	// dL/dSigma = T^T H T = T0 (x) U0 + T1 (x) U1 with U = H T
	const V3 U0 = axpby(ha, T0, hb, T1), U1 = axpby(hb, T0, hc, T1);
	out.dSigma.xx = T0.x * U0.x + T1.x * U1.x;
	out.dSigma.xy = T0.x * U0.y + T1.x * U1.y;
	out.dSigma.xz = T0.x * U0.z + T1.x * U1.z;
	out.dSigma.yy = T0.y * U0.y + T1.y * U1.y;
	out.dSigma.yz = T0.y * U0.z + T1.y * U1.z;
	out.dSigma.zz = T0.z * U0.z + T1.z * U1.z;
	// dL/dT = 2 H T Sigma: rows 2 (ha Sigma T0 + hb Sigma T1), 2 (hb Sigma T0 + hc Sigma T1)
	const V3 dT0 = axpby(2.f * ha, ST0, 2.f * hb, ST1), dT1 = axpby(2.f * hb, ST0, 2.f * hc, ST1);
	// T = J Wr: only j00, j02, j11, j12 are live
	const float d00 = dot(dT0, W0), d02 = dot(dT0, W2), d11 = dot(dT1, W1), d12 = dot(dT1, W2);
	const float dux = -fx * iz2 * d02, duy = -fy * iz2 * d12;
	const float dtx = free_x ? dux : 0.f, dty = free_y ? duy : 0.f;
	const float dtz = -iz2 * (fx * d00 + fy * d11) + 2.f * iz2 * iz * (fx * ux * d02 + fy * uy * d12);
	// t = Wr p + t0
	out.dmean = V3{W0.x * dtx + W1.x * dty + W2.x * dtz, W0.y * dtx + W1.y * dty + W2.y * dtz, W0.z * dtx + W1.z * dty + W2.z * dtz};
	return out;
}

// ---- pixel position: gradient of the 2D mean (already scaled to NDC units: x 0.5 W, x 0.5 H) -> dL/dmean ----
// ndc = (proj p)_xy / ((proj p)_w + 1e-7)   (reference forward.cu:196-198)
BRS_HD V3 pixel_backward(const float* proj, const V3& mean, float gndc_x, float gndc_y)
{
	const float hx = proj[0] * mean.x + proj[4] * mean.y + proj[8] * mean.z + proj[12];
	const float hy = proj[1] * mean.x + proj[5] * mean.y + proj[9] * mean.z + proj[13];
	const float hw = proj[3] * mean.x + proj[7] * mean.y + proj[11] * mean.z + proj[15];
	const float iw = 1.f / (hw + 0.0000001f);
	const float u = gndc_x * iw, v = gndc_y * iw;
	const float w = -(hx * u + hy * v) * iw; // through the division
	return V3{proj[0] * u + proj[1] * v + proj[3] * w, proj[4] * u + proj[5] * v + proj[7] * w, proj[8] * u + proj[9] * v + proj[11] * w};
}

// ---- 3D covariance: dL/dSigma (matrix gradient) -> dL/ds (s = mod * scale) and dL/dq ----
struct ShapeGrad {
	V3 dscale;
	float dr, dx, dy, dz;
};
BRS_HD ShapeGrad shape_backward(const Rot& R, const V3& s, float qr, float qx, float qy, float qz, const Sym3& G)
{
	const V3 g0 = mul(G, R.c0), g1 = mul(G, R.c1), g2 = mul(G, R.c2);
	ShapeGrad o;
	o.dscale = V3{2.f * s.x * dot(R.c0, g0), 2.f * s.y * dot(R.c1, g1), 2.f * s.z * dot(R.c2, g2)};
	// K = 2 G Rq diag(s^2): column i = 2 s_i^2 G r_i; K_ab = entry (row a, column b)
	const float wa = 2.f * s.x * s.x, wb = 2.f * s.y * s.y, wc = 2.f * s.z * s.z;
	const float k00 = wa * g0.x, k10 = wa * g0.y, k20 = wa * g0.z;
	const float k01 = wb * g1.x, k11 = wb * g1.y, k21 = wb * g1.z;
	const float k02 = wc * g2.x, k12 = wc * g2.y, k22 = wc * g2.z;
	// Rq = [1-2(yy+zz) 2(xy-rz) 2(xz+ry); 2(xy+rz) 1-2(xx+zz) 2(yz-rx); 2(xz-ry) 2(yz+rx) 1-2(xx+yy)]
	const float a01 = k10 - k01, a02 = k02 - k20, a12 = k21 - k12; // antisymmetric parts (signed as they enter d/dr)
	const float s01 = k01 + k10, s02 = k02 + k20, s12 = k12 + k21;
	o.dr = 2.f * (qz * a01 + qy * a02 + qx * a12);
	o.dx = 2.f * (qy * s01 + qz * s02 + qr * a12) - 4.f * qx * (k11 + k22);
	o.dy = 2.f * (qx * s01 + qr * a02 + qz * s12) - 4.f * qy * (k00 + k22);
	o.dz = 2.f * (qr * a01 + qx * s02 + qy * s12) - 4.f * qz * (k00 + k11);
	return o;
}

// ---- spherical harmonics ----
// Real SH basis of degree <= 3 in the reference's ordering and sign convention (forward.cu:20-71):
// colour = sum_k b_k(dir) sh_k + 0.5.  b[k] for k >= (deg + 1)^2 is set to 0.
BRS_HD void sh_basis(int deg, float x, float y, float z, float* b)
{
	const float c1 = 0.4886025119029199f;
	b[0] = 0.28209479177387814f;
#pragma unroll
	for (int k = 1; k < 16; k++)
		b[k] = 0.f;
	if (deg < 1)
		return;
	b[1] = -c1 * y;
	b[2] = c1 * z;
	b[3] = -c1 * x;
	if (deg < 2)
		return;
	const float xx = x * x, yy = y * y, zz = z * z;
	b[4] = 1.0925484305920792f * x * y;
	b[5] = -1.0925484305920792f * y * z;
	b[6] = 0.31539156525252005f * (2.f * zz - xx - yy);
	b[7] = -1.0925484305920792f * x * z;
	b[8] = 0.5462742152960396f * (xx - yy);
	if (deg < 3)
		return;
	b[9] = -0.5900435899266435f * y * (3.f * xx - yy);
	b[10] = 2.890611442640554f * x * y * z;
	b[11] = -0.4570457994644658f * y * (4.f * zz - xx - yy);
	b[12] = 0.3731763325901154f * z * (2.f * zz - 3.f * xx - 3.f * yy);
	b[13] = -0.4570457994644658f * x * (4.f * zz - xx - yy);
	b[14] = 1.445305721320277f * z * (xx - yy);
	b[15] = -0.5900435899266435f * x * (xx - 3.f * yy);
}

// dL/ddir = sum_k q_k grad b_k(dir), with q_k = sh_k . dL_dRGB already contracted over the colour channels
// (the gradients of the basis polynomials above, term by term).
BRS_HD V3 sh_direction_gradient(int deg, float x, float y, float z, const float* q)
{
	V3 g = {0.f, 0.f, 0.f};
	if (deg < 1)
		return g;
	const float c1 = 0.4886025119029199f;
	g.x = -c1 * q[3];
	g.y = -c1 * q[1];
	g.z = c1 * q[2];
	if (deg < 2)
		return g;
	const float q4 = 1.0925484305920792f * q[4], q5 = -1.0925484305920792f * q[5], q6 = 0.31539156525252005f * q[6],
	            q7 = -1.0925484305920792f * q[7], q8 = 0.5462742152960396f * q[8];
	g.x += q4 * y - 2.f * q6 * x + q7 * z + 2.f * q8 * x;
	g.y += q4 * x + q5 * z - 2.f * q6 * y - 2.f * q8 * y;
	g.z += q5 * y + 4.f * q6 * z + q7 * x;
	if (deg < 3)
		return g;
	const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
	const float q9 = -0.5900435899266435f * q[9], q10 = 2.890611442640554f * q[10], q11 = -0.4570457994644658f * q[11],
	            q12 = 0.3731763325901154f * q[12], q13 = -0.4570457994644658f * q[13], q14 = 1.445305721320277f * q[14],
	            q15 = -0.5900435899266435f * q[15];
	g.x += 6.f * q9 * xy + q10 * yz - 2.f * q11 * xy - 6.f * q12 * xz + q13 * (4.f * zz - 3.f * xx - yy) + 2.f * q14 * xz +
	       3.f * q15 * (xx - yy);
	g.y += 3.f * q9 * (xx - yy) + q10 * xz + q11 * (4.f * zz - xx - 3.f * yy) - 6.f * q12 * yz - 2.f * q13 * xy - 2.f * q14 * yz -
	       6.f * q15 * xy;
	g.z += q10 * xy + 8.f * q11 * yz + 3.f * q12 * (2.f * zz - xx - yy) + 8.f * q13 * xz + q14 * (xx - yy);
	return g;
}

// dir = v / |v|:  dL/dv = (g - dir (dir . g)) / |v|
BRS_HD V3 normalize_backward(const V3& dir, float inv_len, const V3& g)
{
	const float p = dot(dir, g);
	return V3{(g.x - dir.x * p) * inv_len, (g.y - dir.y * p) * inv_len, (g.z - dir.z * p) * inv_len};
}

} // namespace bwdmath
} // namespace brs
