"""CPU check of the re-derived preprocess-backward maths (bloomscene_b200/csrc/preprocess_bwd_math.h): the header
is compiled for the host with g++ and every function is compared with central finite differences of a float64
forward model written here from the forward definitions (reference forward.cu:20-152,186-237)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "tests", "native", "bwd_math_shim.cpp")
OUT = os.path.join(ROOT, "tests", "native", "_build", "libbwdmath.so")


@pytest.fixture(scope="module")
def lib():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    hdr = os.path.join(ROOT, "bloomscene_b200", "csrc", "preprocess_bwd_math.h")
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(SHIM), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", SHIM, "-o", OUT], check=True)
    return C.CDLL(OUT)


def fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def f32(x):
    return np.ascontiguousarray(x, dtype=np.float32)


def rot(q):
    r, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)],
                     [2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)],
                     [2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)]])


def sym(s6):
    xx, xy, xz, yy, yz, zz = s6
    return np.array([[xx, xy, xz], [xy, yy, yz], [xz, yz, zz]])


def num_grad(f, x, h=1e-5):
    x = np.array(x, dtype=np.float64)
    g = np.zeros_like(x)
    for i in range(x.size):
        d = np.zeros_like(x)
        d.flat[i] = h
        g.flat[i] = (f(x + d) - f(x - d)) / (2 * h)
    return g


def conic_loss(view, mean, Sigma, fx, fy, tx_, ty_, g3):
    Wr = view.reshape(4, 4).T[:3, :3]
    t = Wr @ mean + view.reshape(4, 4)[3, :3]
    limx, limy = 1.3 * tx_, 1.3 * ty_
    ux = min(limx, max(-limx, t[0] / t[2])) * t[2]
    uy = min(limy, max(-limy, t[1] / t[2])) * t[2]
    J = np.array([[fx / t[2], 0, -fx * ux / t[2] ** 2], [0, fy / t[2], -fy * uy / t[2] ** 2]])
    T = J @ Wr
    C2 = T @ Sigma @ T.T + 0.3 * np.eye(2)
    M = np.linalg.inv(C2)
    return g3[0] * M[0, 0] + 2 * g3[1] * M[0, 1] + g3[2] * M[1, 1]


def test_projection_backward_matches_finite_differences(lib):
    rng = np.random.default_rng(0)
    for _ in range(20):
        ang = rng.uniform(0, 6.28)
        Rw = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
        view = np.eye(4)
        view[:3, :3] = Rw.T  # memory = transpose of the maths matrix
        view[3, :3] = [0.1, -0.2, 3.0]
        mean = rng.uniform(-0.6, 0.6, 3)
        A = rng.normal(size=(3, 3)) * 0.05
        Sigma = A @ A.T + 1e-3 * np.eye(3)
        s6 = np.array([Sigma[0, 0], Sigma[0, 1], Sigma[0, 2], Sigma[1, 1], Sigma[1, 2], Sigma[2, 2]])
        g3 = rng.normal(size=3)
        fx, fy, tx_, ty_ = 600.0, 580.0, 0.45, 0.3
        out = np.zeros(9, dtype=np.float32)
        lib.shim_projection_backward(fptr(f32(view)), fptr(f32(mean)), fptr(f32(s6)), C.c_float(fx), C.c_float(fy),
                                     C.c_float(tx_), C.c_float(ty_), fptr(f32(g3)), fptr(out))
        v64 = view.flatten()
        gm = num_grad(lambda m: conic_loss(v64, m, Sigma, fx, fy, tx_, ty_, g3), mean, 1e-6)
        # matrix gradient with respect to a symmetric perturbation of Sigma: off-diagonals count twice
        gs = num_grad(lambda s: conic_loss(v64, mean, sym(s), fx, fy, tx_, ty_, g3), s6, 1e-7)
        gs[[1, 2, 4]] *= 0.5
        assert np.allclose(out[6:9], gm, rtol=2e-3, atol=2e-3 * np.abs(gm).max()), (out[6:9], gm)
        assert np.allclose(out[:6], gs, rtol=2e-3, atol=2e-3 * np.abs(gs).max()), (out[:6], gs)


def test_projection_backward_clamp_quirk(lib):
    """Outside +-1.3 tan(fov) the reference zeroes the x / y gradient of t (backward.cu:195-196) and still
    differentiates u only through its explicit t_z factors; the mean gradient must lose exactly that term."""
    view = np.eye(4)
    view[3, :3] = [0.0, 0.0, 1.0]
    mean = np.array([2.0, 0.1, 0.5])  # t_x / t_z = 1.33 > 1.3 * 0.45
    Sigma = np.diag([0.01, 0.02, 0.015])
    s6 = np.array([0.01, 0, 0, 0.02, 0, 0.015])
    g3 = np.array([0.3, -0.2, 0.5])
    out = np.zeros(9, dtype=np.float32)
    lib.shim_projection_backward(fptr(f32(view)), fptr(f32(mean)), fptr(f32(s6)), C.c_float(600.0), C.c_float(580.0),
                                 C.c_float(0.45), C.c_float(0.3), fptr(f32(g3)), fptr(out))
    free = np.zeros(9, dtype=np.float32)
    lib.shim_projection_backward(fptr(f32(view)), fptr(f32(mean)), fptr(f32(s6)), C.c_float(600.0), C.c_float(580.0),
                                 C.c_float(4.5), C.c_float(3.0), fptr(f32(g3)), fptr(free))
    assert out[6] == 0.0 and free[6] != 0.0  # with the identity rotation dmean.x is exactly dL/dt_x
    assert np.isfinite(out).all()


def test_pixel_backward_matches_finite_differences(lib):
    rng = np.random.default_rng(1)
    for _ in range(10):
        proj = rng.normal(size=16)
        proj[15] = 3.0
        mean = rng.uniform(-1, 1, 3)
        gx, gy = rng.normal(size=2)

        def loss(m):
            P = proj.reshape(4, 4)
            h = np.array([P[:, k] @ np.append(m, 1.0) for k in range(4)])  # row-vector convention: h_k = sum_j m_j P[j][k]
            iw = 1.0 / (h[3] + 1e-7)
            return gx * h[0] * iw + gy * h[1] * iw

        out = np.zeros(3, dtype=np.float32)
        lib.shim_pixel_backward(fptr(f32(proj)), fptr(f32(mean)), C.c_float(gx), C.c_float(gy), fptr(out))
        g = num_grad(loss, mean, 1e-6)
        assert np.allclose(out, g, rtol=2e-3, atol=2e-3 * np.abs(g).max()), (out, g)


def test_shape_backward_matches_finite_differences(lib):
    rng = np.random.default_rng(2)
    for _ in range(20):
        s = np.exp(rng.normal(-1.0, 0.5, 3))
        q = rng.normal(size=4)
        q /= np.linalg.norm(q) * rng.uniform(0.8, 1.2)  # not normalised on purpose
        G6 = rng.normal(size=6)
        G = sym(G6)
        out = np.zeros(13, dtype=np.float32)
        lib.shim_shape(fptr(f32(s)), fptr(f32(q)), fptr(f32(G6)), fptr(out))
        Sigma = lambda s_, q_: rot(q_) @ np.diag(s_ ** 2) @ rot(q_).T
        assert np.allclose(sym(out[:6]), Sigma(s, q), rtol=1e-4, atol=1e-5)
        gs = num_grad(lambda s_: np.sum(G * Sigma(s_, q)), s)
        gq = num_grad(lambda q_: np.sum(G * Sigma(s, q_)), q)
        assert np.allclose(out[6:9], gs, rtol=2e-3, atol=2e-3 * np.abs(gs).max()), (out[6:9], gs)
        assert np.allclose(out[9:13], gq, rtol=2e-3, atol=2e-3 * np.abs(gq).max()), (out[9:13], gq)


def sh_basis64(deg, d):
    x, y, z = d
    C0, C1 = 0.28209479177387814, 0.4886025119029199
    C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
    C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
          1.445305721320277, -0.5900435899266435]
    b = np.zeros(16)
    b[0] = C0
    if deg > 0:
        b[1:4] = [-C1 * y, C1 * z, -C1 * x]
    if deg > 1:
        xx, yy, zz = x * x, y * y, z * z
        b[4:9] = [C2[0] * x * y, C2[1] * y * z, C2[2] * (2 * zz - xx - yy), C2[3] * x * z, C2[4] * (xx - yy)]
    if deg > 2:
        b[9:16] = [C3[0] * y * (3 * xx - yy), C3[1] * x * y * z, C3[2] * y * (4 * zz - xx - yy),
                   C3[3] * z * (2 * zz - 3 * xx - 3 * yy), C3[4] * x * (4 * zz - xx - yy), C3[5] * z * (xx - yy),
                   C3[6] * x * (xx - 3 * yy)]
    return b


@pytest.mark.parametrize("deg", [0, 1, 2, 3])
def test_sh_basis_and_direction_gradient(lib, deg):
    rng = np.random.default_rng(3 + deg)
    for _ in range(10):
        v = rng.normal(size=3) * rng.uniform(0.5, 4.0)
        qk = rng.normal(size=16)
        basis = np.zeros(16, dtype=np.float32)
        dv = np.zeros(3, dtype=np.float32)
        lib.shim_sh(deg, fptr(f32(v)), fptr(f32(qk)), fptr(basis), fptr(dv))
        want = sh_basis64(deg, v / np.linalg.norm(v))
        assert np.allclose(basis, want, rtol=1e-5, atol=1e-6)
        g = num_grad(lambda v_: qk @ sh_basis64(deg, v_ / np.linalg.norm(v_)), v, 1e-6)
        assert np.allclose(dv, g, rtol=3e-3, atol=3e-3 * max(np.abs(g).max(), 1e-6)), (dv, g)
