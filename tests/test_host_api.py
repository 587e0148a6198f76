"""CPU: the drop-in boundary — C-ABI exports, argument validation, Python surface, gradient plumbing."""
import ctypes as C
import inspect
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    return C.CDLL(os.path.join(ROOT, "bloomscene_b200", "libbloomrast.so"))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "bloomrast.h")).read()
    declared = set(re.findall(r"\b(brs_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"brs_alloc_fn"}
    assert {"brs_forward", "brs_backward", "brs_visible_filter", "brs_mark_visible", "brs_sort_pairs_u32"} <= declared
    lib = _lib()
    for name in sorted(declared):
        assert hasattr(lib, name), name


def test_version_sizes_and_error_strings():
    lib = _lib()
    assert lib.brs_version() == 200
    for f in ("brs_geom_bytes", "brs_binning_bytes", "brs_sort_scratch_bytes", "brs_backward_scratch_bytes"):
        getattr(lib, f).restype = C.c_size_t
    lib.brs_image_bytes.restype = C.c_size_t
    # pure functions of the sizes, monotone, 256-byte granular
    g1, g2 = lib.brs_geom_bytes(1000), lib.brs_geom_bytes(2000)
    assert 0 < g1 < g2 and g1 % 256 == 0
    assert lib.brs_geom_bytes(1000) == g1
    assert lib.brs_binning_bytes(10) >= 40 and lib.brs_binning_bytes(0) > 0
    assert lib.brs_image_bytes(1920, 1080) >= 1920 * 1080 * 8
    lib.brs_error_string.restype = C.c_char_p
    assert lib.brs_error_string(0) == b"ok"
    assert b"invalid" in lib.brs_error_string(-1)
    assert b"unknown" in lib.brs_error_string(-99)


def test_c_abi_rejects_bad_arguments_without_touching_the_gpu():
    lib = _lib()
    # NULL view / gaussians / state -> BRS_ERR_INVALID_ARG (-1), never a crash
    assert lib.brs_forward(None, None, None, None, None, None, None, None, None) == -1
    assert lib.brs_backward(None, None, None, None, None, None, None, None, None, None) == -1
    assert lib.brs_visible_filter(None, 10, None, None, 3, None, None, None, None) == -1
    assert lib.brs_mark_visible(-1, None, None, None, None, None) == -1
    assert lib.brs_mark_visible(0, None, None, None, None, None) == 0
    assert lib.brs_sort_pairs_u32(None, None, None, None, -5, 0, 32, None, None) == -1
    assert lib.brs_sort_pairs_u32(None, None, None, None, 0, 0, 32, None, None) == 0
    assert lib.brs_sort_pairs_u32(None, None, None, None, 8, 0, 33, None, None) == -1


class _View(C.Structure):
    _fields_ = [("image_width", C.c_int), ("image_height", C.c_int), ("tanfovx", C.c_float), ("tanfovy", C.c_float),
                ("scale_modifier", C.c_float), ("sh_degree", C.c_int), ("sh_coeffs", C.c_int), ("prefiltered", C.c_int),
                ("debug", C.c_int), ("bg", C.c_void_p), ("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p),
                ("campos", C.c_void_p)]


class _Gaussians(C.Structure):
    _fields_ = [("P", C.c_int)] + [(n, C.c_void_p) for n in ("means3D", "opacities", "shs", "colors_precomp", "scales",
                                                             "rotations", "cov3D_precomp")]


def test_forward_views_validates_the_stack_without_touching_the_gpu():
    """brs_forward_views: the views of a stack must agree on image size and SH layout; counts must fit 32-bit instance ids
    and 16-bit stacked tile rows.  All rejected before any CUDA call (the pointers are fake)."""
    lib = _lib()
    ALLOC = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_int, C.c_size_t)
    alloc = ALLOC(lambda ctx, which, n: None)
    fake = 0x1000

    def view(W=64, H=48, deg=0):
        return _View(W, H, 0.5, 0.5, 1.0, deg, 0, 0, 0, fake, fake, fake, fake)

    def call(views, g, n=None):
        arr = (_View * len(views))(*views)
        return lib.brs_forward_views(arr, len(views) if n is None else n, C.byref(g), C.c_void_p(fake), C.c_void_p(fake),
                                     C.c_void_p(fake), alloc, None, None, None, None)

    g = _Gaussians(100, fake, fake, None, fake, fake, fake, None)
    assert lib.brs_forward_views(None, 2, C.byref(g), None, None, None, alloc, None, None, None, None) == -1
    assert call([view()], g, n=0) == -1
    assert call([view(), view(W=80)], g) == -1            # different image sizes in one stack
    assert call([view(), view(deg=1)], g) == -1           # different SH degree
    both = _Gaussians(100, fake, fake, fake, fake, fake, fake, None)
    assert call([view(), view()], both) == -1             # shs AND colors_precomp (reference raises for it)
    big = _Gaussians(2**30, fake, fake, None, fake, fake, fake, None)
    assert call([view(), view(), view()], big) == -4      # 3 * 2^30 instances do not fit 32-bit ids: BRS_ERR_UNSUPPORTED
    assert call([view(H=16 * 40000), view(H=16 * 40000)], g) == -4   # 80 000 stacked tile rows > 16 bits


def test_state_layout_is_consistent():
    from bloomscene_b200 import _C

    lay = _C.state_layout(1000, 5000, 130, 70)
    assert lay["geom_records"] % 256 == 0 and lay["geom_depth_key"] >= lay["geom_records"] + 48 * 1000
    assert lay["geom_rect"] >= lay["geom_depth_key"] + 4000 and lay["geom_order"] >= lay["geom_rect"] + 8000
    ntiles = 9 * 5
    assert lay["image_final_T"] >= lay["image_ranges"] + 8 * ntiles
    assert lay["image_n_contrib"] >= lay["image_final_T"] + 4 * 130 * 70


def test_python_surface_matches_reference_package():
    import depth_diff_gaussian_rasterization as ddgr
    import diff_gaussian_rasterization as dgr

    S = ddgr.GaussianRasterizationSettings
    # field order is API (reference __init__.py:158-170)
    assert S._fields == ("image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix",
                         "projmatrix", "sh_degree", "campos", "prefiltered", "debug")
    R = ddgr.GaussianRasterizer
    assert list(inspect.signature(R.forward).parameters) == ["self", "means3D", "means2D", "opacities", "shs",
                                                              "colors_precomp", "scales", "rotations", "cov3D_precomp"]
    assert list(inspect.signature(R.visible_filter).parameters) == ["self", "means3D", "scales", "rotations",
                                                                     "cov3D_precomp"]
    assert hasattr(R, "markVisible") and callable(ddgr.rasterize_gaussians)
    assert dgr.GaussianRasterizer is R and dgr.GaussianRasterizationSettings is S
    # the four reference binding names (ext.cpp:15-20), including the reference's own spelling
    for fn in ("rasterize_gaussians", "rasterize_gaussians_backward", "rasterize_aussians_filter", "mark_visible"):
        assert hasattr(ddgr._C, fn)
    # extensions stay keyword-only additions: the batch entry point and its stacked pipeline
    assert "stack" in inspect.signature(ddgr.render_views).parameters and hasattr(ddgr._C, "rasterize_gaussians_views")


def _settings(S, W=32, H=32):
    eye = torch.eye(4)
    return S(image_height=H, image_width=W, tanfovx=0.5, tanfovy=0.5, bg=torch.zeros(3), scale_modifier=1.0,
             viewmatrix=eye, projmatrix=eye, sh_degree=0, campos=torch.zeros(3), prefiltered=False, debug=False)


def test_python_argument_errors_match_reference_messages():
    import depth_diff_gaussian_rasterization as ddgr

    r = ddgr.GaussianRasterizer(_settings(ddgr.GaussianRasterizationSettings))
    P = 4
    m, m2, o = torch.zeros(P, 3), torch.zeros(P, 3), torch.ones(P, 1)
    sh, col, s, q, cov = torch.zeros(P, 1, 3), torch.zeros(P, 3), torch.ones(P, 3), torch.zeros(P, 4), torch.zeros(P, 6)
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(m, m2, o, shs=None, colors_precomp=None, scales=s, rotations=q)
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(m, m2, o, shs=sh, colors_precomp=col, scales=s, rotations=q)
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair or precomputed 3D covariance"):
        r(m, m2, o, colors_precomp=col)
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair or precomputed 3D covariance"):
        r(m, m2, o, colors_precomp=col, scales=s, rotations=q, cov3D_precomp=cov)
    # no silent CPU fallback: CPU tensors are refused by the native binding
    with pytest.raises(Exception, match="CUDA"):
        r(m, m2, o, colors_precomp=col, scales=s, rotations=q)
    with pytest.raises(Exception, match=r"\(num_points, 3\)"):
        ddgr._C.rasterize_aussians_filter(torch.zeros(P, 4), s, q, 1.0, torch.Tensor([]), torch.eye(4), torch.eye(4),
                                          0.5, 0.5, 32, 32, False, False)


def test_autograd_plumbing_with_oracle_backend():
    """The shared Python wrapper routes gradients to the right inputs (reference __init__.py:144-154)."""
    from workload import synthetic
    from bloomscene_b200.rasterizer import GaussianRasterizationSettings, bind
    from oracle_backend import OracleBackend

    api = bind(OracleBackend())
    scene = synthetic.make_scene(400, "object", "sh1", -3.0, seed=3)
    cam = synthetic.orbit_camera(48, 32, 0.2)
    leaf = lambda t: t.clone().requires_grad_(True)
    means, sc, rot, op, sh = map(leaf, (scene.means3D, scene.scales, scene.rotations, scene.opacities, scene.shs))
    m2 = torch.zeros_like(means, requires_grad=True)
    rast = api.GaussianRasterizer(synthetic.raster_settings(cam, 1, torch.tensor([0.1, 0.2, 0.3]),
                                                            GaussianRasterizationSettings))
    color, radii, depth = rast(means, m2, op, shs=sh, scales=sc, rotations=rot)
    assert color.shape == (3, 32, 48) and depth.shape == (1, 32, 48) and radii.shape == (400,) and radii.dtype == torch.int32
    Wc, Wd = synthetic.loss_weights(48, 32)
    ((color * Wc).sum() + (depth * Wd).sum()).backward()
    for t, shape in ((means, (400, 3)), (m2, (400, 3)), (op, (400, 1)), (sc, (400, 3)), (rot, (400, 4)), (sh, (400, 4, 3))):
        assert t.grad is not None and tuple(t.grad.shape) == shape and torch.isfinite(t.grad).all()
    assert not m2.grad[:, 2].any() and m2.grad[:, :2].abs().sum() > 0
    # depth has no gradient in this fork (backward.cu:443-554): a depth-only loss gives all-zero gradients
    means.grad = None
    m2b = torch.zeros_like(means, requires_grad=True)
    color, radii, depth = rast(means, m2b, op, shs=sh, scales=sc, rotations=rot)
    (depth * Wd).sum().backward()
    assert not means.grad.any() and not m2b.grad.any()
    # finite-difference check of one opacity derivative (away from thresholds the map is smooth)
    with torch.no_grad():
        mid = ((op > 0.2) & (op < 0.8)).float().squeeze(1)  # stay clear of the 0.99 clamp the gradient ignores
        i = int(torch.argmax(op.grad.abs().squeeze(1) * mid))
        op_p, op_m = op.detach().clone(), op.detach().clone()
        op_p[i] += 5e-3
        op_m[i] -= 5e-3
        pert = (rast(means, m2, op_p, shs=sh, scales=sc, rotations=rot)[0].double() * Wc).sum().item()
        base = (rast(means, m2, op_m, shs=sh, scales=sc, rotations=rot)[0].double() * Wc).sum().item()
    op.grad = None
    color, _, _ = rast(means, m2, op, shs=sh, scales=sc, rotations=rot)
    (color * Wc).sum().backward()
    fd = (pert - base) / 1e-2
    assert abs(fd - op.grad[i].item()) <= 0.1 * abs(fd) + 1e-2


def test_render_views_equals_single_view_calls_with_oracle_backend():
    """render_views (forward-only batch entry point) returns exactly what one GaussianRasterizer call per view returns."""
    from workload import synthetic
    from bloomscene_b200.rasterizer import GaussianRasterizationSettings, bind
    from oracle_backend import OracleBackend

    api = bind(OracleBackend())
    scene = synthetic.make_scene(300, "object", "precomp", -3.0, seed=4)
    bg = torch.tensor([0.2, 0.1, 0.0])
    settings = [synthetic.raster_settings(synthetic.orbit_camera(40, 24, 0.7 * k), 0, bg, GaussianRasterizationSettings)
                for k in range(3)]
    color, depth, radii = api.render_views(settings, scene.means3D, scene.opacities, colors_precomp=scene.colors_precomp,
                                           scales=scene.scales, rotations=scene.rotations, keep_radii=True)
    assert color.shape == (3, 3, 24, 40) and depth.shape == (3, 1, 24, 40) and len(radii) == 3
    for k, rs in enumerate(settings):
        c, r, d = api.GaussianRasterizer(rs)(scene.means3D, torch.zeros_like(scene.means3D), scene.opacities,
                                             colors_precomp=scene.colors_precomp, scales=scene.scales,
                                             rotations=scene.rotations)
        assert torch.equal(c, color[k]) and torch.equal(d, depth[k]) and torch.equal(r, radii[k])
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        api.render_views(settings, scene.means3D, scene.opacities, scales=scene.scales, rotations=scene.rotations)
    with pytest.raises(Exception, match="share one resolution"):
        other = synthetic.raster_settings(synthetic.orbit_camera(32, 24, 0.1), 0, bg, GaussianRasterizationSettings)
        api.render_views(settings + [other], scene.means3D, scene.opacities, colors_precomp=scene.colors_precomp,
                         scales=scene.scales, rotations=scene.rotations)


def test_opt_in_depth_gradient_with_oracle_backend():
    """depth_gradient=True (extension, default off): a depth-only loss reaches the parameters, and the
    oracle's derivative agrees with central finite differences of its own forward."""
    from workload import synthetic
    from bloomscene_b200.rasterizer import GaussianRasterizationSettings, bind
    from oracle_backend import OracleBackend

    api = bind(OracleBackend())
    scene = synthetic.make_scene(400, "object", "sh1", -3.0, seed=3)
    cam = synthetic.orbit_camera(48, 32, 0.2)
    settings = synthetic.raster_settings(cam, 1, torch.tensor([0.1, 0.2, 0.3]), GaussianRasterizationSettings)
    _, Wd = synthetic.loss_weights(48, 32)
    rast = api.GaussianRasterizer(settings, depth_gradient=True)
    assert api.GaussianRasterizer.supports_depth_gradient

    def depth_loss(means, op):
        return (rast(means, torch.zeros_like(means), op, shs=scene.shs, scales=scene.scales,
                     rotations=scene.rotations)[2].double() * Wd).sum()

    means, op = scene.means3D.clone().requires_grad_(True), scene.opacities.clone().requires_grad_(True)
    depth_loss(means, op).backward()
    assert means.grad.abs().sum() > 0 and op.grad.abs().sum() > 0

    def fd(t, index, eps):
        with torch.no_grad():
            plus, minus = t.detach().clone(), t.detach().clone()
            plus[index] += eps
            minus[index] -= eps
            args = (plus, op) if t is means else (means, plus)
            args_m = (minus, op) if t is means else (means, minus)
            return (depth_loss(*args).item() - depth_loss(*args_m).item()) / (2 * eps)

    # the map has jumps (acc > 0.5 gate, alpha thresholds, radius changes): take the best-agreeing three of the
    # six largest gradient entries of each tensor, all three must match to 2 %
    for t, eps in ((means, 2e-4), (op, 5e-3)):
        g = t.grad.reshape(-1)
        if t is op:
            g = g * ((op.detach().reshape(-1) > 0.2) & (op.detach().reshape(-1) < 0.8))  # clear of the ignored 0.99 clamp
        top = torch.argsort(-g.abs())[:6]
        errs = []
        for i in top.tolist():
            index = tuple(int(v) for v in np.unravel_index(i, t.shape))
            errs.append(abs(fd(t, index, eps) - t.grad[index].item()) / abs(t.grad[index].item()))
        assert sorted(errs)[2] <= 0.02, errs


def test_oracle_double_precision_sums_agree_with_float_sums_on_a_small_scene():
    """Oracle.backward(f64_sums=True) — the referee of the screen-filling-Gaussians gradient test — changes only the
    ACCUMULATION of the per-Gaussian sums: on a small scene (short sums) it agrees with the fp32 sums to float rounding,
    and it leaves no state behind."""
    from oracle import oracle
    from workload import synthetic

    scene = synthetic.make_scene(300, "object", "sh1", -2.8, seed=4)
    cam = synthetic.orbit_camera(48, 40, 0.3)
    Wc, _ = synthetic.loss_weights(48, 40, seed=2)
    o = oracle.run_scene(scene, cam, torch.tensor([0.1, 0.2, 0.3]))["oracle"]
    a = o.backward(Wc.numpy())
    b = o.backward(Wc.numpy(), f64_sums=True)
    c = o.backward(Wc.numpy())
    for k in a:
        na = np.linalg.norm(a[k])
        assert np.linalg.norm(a[k] - b[k]) <= 2e-6 * max(na, 1e-30), k
        assert np.linalg.norm(a[k] - c[k]) <= 2e-6 * max(na, 1e-30), k   # fp32 again (OpenMP order noise only)
    assert np.abs(a["dL_dopacity"]).sum() > 0
