// Host build of bloomscene_b200/csrc/preprocess_bwd_math.h for tests/test_bwd_math.py (plain C ABI over flat arrays).
#include <cmath>
using std::fmaxf;
using std::fminf;
#include "../../bloomscene_b200/csrc/preprocess_bwd_math.h"

using namespace brs::bwdmath;

extern "C" {

// out[9] = dSigma (xx xy xz yy yz zz, matrix gradient) + dmean
void shim_projection_backward(const float* view, const float* mean, const float* sigma6, float fx, float fy, float tan_fovx,
                              float tan_fovy, const float* g3, float* out)
{
	const Sym3 S{sigma6[0], sigma6[1], sigma6[2], sigma6[3], sigma6[4], sigma6[5]};
	const ProjectionGrad r = projection_backward(view, vec3(mean[0], mean[1], mean[2]), S, fx, fy, tan_fovx, tan_fovy, g3[0], g3[1], g3[2]);
	out[0] = r.dSigma.xx; out[1] = r.dSigma.xy; out[2] = r.dSigma.xz; out[3] = r.dSigma.yy; out[4] = r.dSigma.yz; out[5] = r.dSigma.zz;
	out[6] = r.dmean.x; out[7] = r.dmean.y; out[8] = r.dmean.z;
}

void shim_pixel_backward(const float* proj, const float* mean, float gx, float gy, float* out)
{
	const V3 r = pixel_backward(proj, vec3(mean[0], mean[1], mean[2]), gx, gy);
	out[0] = r.x; out[1] = r.y; out[2] = r.z;
}

// out[13] = Sigma(6) + dscale(3) + dq(4)
void shim_shape(const float* s, const float* q, const float* g6, float* out)
{
	const Rot R = rotation(q[0], q[1], q[2], q[3]);
	const V3 sv = vec3(s[0], s[1], s[2]);
	const Sym3 S = covariance(R, sv);
	out[0] = S.xx; out[1] = S.xy; out[2] = S.xz; out[3] = S.yy; out[4] = S.yz; out[5] = S.zz;
	const Sym3 G{g6[0], g6[1], g6[2], g6[3], g6[4], g6[5]};
	const ShapeGrad r = shape_backward(R, sv, q[0], q[1], q[2], q[3], G);
	out[6] = r.dscale.x; out[7] = r.dscale.y; out[8] = r.dscale.z;
	out[9] = r.dr; out[10] = r.dx; out[11] = r.dy; out[12] = r.dz;
}

// basis[16]; dv[3] = gradient with respect to the UNNORMALISED direction v, given q[16]
void shim_sh(int deg, const float* v, const float* q, float* basis, float* dv)
{
	const float len = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
	const V3 dir = vec3(v[0] / len, v[1] / len, v[2] / len);
	sh_basis(deg, dir.x, dir.y, dir.z, basis);
	const V3 g = sh_direction_gradient(deg, dir.x, dir.y, dir.z, q);
	const V3 r = normalize_backward(dir, 1.f / len, g);
	dv[0] = r.x; dv[1] = r.y; dv[2] = r.z;
}
}
