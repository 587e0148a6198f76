"""GPU: the C-ABI on its own — brs_forward / brs_backward / brs_visible_filter / brs_mark_visible driven through
ctypes exactly as a non-torch host (the cgo / JNI / ctypes stub of INTEGRATION.md) would drive them: plain
structs of device pointers, a caller-side allocator callback, status codes.  torch only supplies device memory.
Results must equal what the torch binding (`_C`, which wraps the same entry points) returns, bit for bit for the
forward and within float-atomic noise for the gradients."""
import ctypes as C
import os

import pytest
import torch

import parity_lib as pl
from workload import synthetic

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV = "cuda:0"
FP = C.POINTER(C.c_float)


class View(C.Structure):  # brs_view
    _fields_ = [("image_width", C.c_int), ("image_height", C.c_int), ("tanfovx", C.c_float), ("tanfovy", C.c_float),
                ("scale_modifier", C.c_float), ("sh_degree", C.c_int), ("sh_coeffs", C.c_int), ("prefiltered", C.c_int),
                ("debug", C.c_int), ("bg", C.c_void_p), ("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p),
                ("campos", C.c_void_p)]


class Gaussians(C.Structure):  # brs_gaussians
    _fields_ = [("P", C.c_int), ("means3D", C.c_void_p), ("opacities", C.c_void_p), ("shs", C.c_void_p),
                ("colors_precomp", C.c_void_p), ("scales", C.c_void_p), ("rotations", C.c_void_p),
                ("cov3D_precomp", C.c_void_p)]


class FwdState(C.Structure):  # brs_fwd_state
    _fields_ = [("geom", C.c_void_p), ("geom_bytes", C.c_size_t), ("binning", C.c_void_p), ("binning_bytes", C.c_size_t),
                ("image", C.c_void_p), ("image_bytes", C.c_size_t), ("num_rendered", C.c_int)]


class Grads(C.Structure):  # brs_grads
    _fields_ = [("dL_dmeans2D", C.c_void_p), ("dL_dcolors", C.c_void_p), ("dL_dopacity", C.c_void_p),
                ("dL_dmeans3D", C.c_void_p), ("dL_dcov3D", C.c_void_p), ("dL_dsh", C.c_void_p), ("dL_dscales", C.c_void_p),
                ("dL_drotations", C.c_void_p), ("accumulate", C.c_int), ("depth_gradient", C.c_int),
                ("out_depth", C.c_void_p)]


ALLOC = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_int, C.c_size_t)  # brs_alloc_fn


def _lib():
    lib = C.CDLL(os.path.join(ROOT, "bloomscene_b200", "libbloomrast.so"))
    lib.brs_forward.argtypes = [C.POINTER(View), C.POINTER(Gaussians), C.c_void_p, C.c_void_p, C.c_void_p, ALLOC,
                                C.c_void_p, C.POINTER(FwdState), C.c_void_p]
    lib.brs_backward.argtypes = [C.POINTER(View), C.POINTER(Gaussians), C.c_void_p, C.POINTER(FwdState), C.c_void_p,
                                 C.c_void_p, C.POINTER(Grads), ALLOC, C.c_void_p, C.c_void_p]
    lib.brs_visible_filter.argtypes = [C.POINTER(View), C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p]
    lib.brs_mark_visible.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.brs_forward_views.argtypes = [C.POINTER(View), C.c_int, C.POINTER(Gaussians), C.c_void_p, C.c_void_p, C.c_void_p,
                                      ALLOC, C.c_void_p, C.POINTER(C.c_longlong), C.c_void_p, C.c_void_p]
    lib.brs_error_string.restype = C.c_char_p
    return lib


class Arena:
    """The caller's side of brs_alloc_fn: hands out torch byte tensors and keeps them alive."""

    def __init__(self):
        self.buffers = {}
        self.scratch = []

        def alloc(ctx, which, nbytes):
            t = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=DEV)
            if which == 3:
                self.scratch.append(t)
            else:
                self.buffers[which] = t
            return t.data_ptr()

        self.fn = ALLOC(alloc)


def _ptr(t):
    return None if t is None else t.data_ptr()


@pytest.mark.parametrize("color", ["sh2", "precomp"])
def test_forward_backward_through_ctypes_only(color):
    lib = _lib()
    api = pl.ours()
    W, H = 150, 94
    scene = synthetic.make_scene(5000, "object", color, -3.4, seed=31).to(DEV)
    cam = synthetic.orbit_camera(W, H, 0.6).to(DEV)
    bg = torch.tensor([0.2, 0.4, 0.1], device=DEV)
    M = 0 if scene.shs is None else scene.shs.shape[1]
    P = scene.P

    view = View(W, H, cam.tanfovx, cam.tanfovy, 1.0, scene.sh_degree, M, 0, 0, _ptr(bg), _ptr(cam.viewmatrix),
                _ptr(cam.projmatrix), _ptr(cam.campos))
    g = Gaussians(P, _ptr(scene.means3D), _ptr(scene.opacities), _ptr(scene.shs), _ptr(scene.colors_precomp),
                  _ptr(scene.scales), _ptr(scene.rotations), None)
    color_img = torch.empty(3, H, W, device=DEV)
    depth_img = torch.empty(1, H, W, device=DEV)
    radii = torch.empty(P, dtype=torch.int32, device=DEV)
    arena, state = Arena(), FwdState()
    stream = torch.cuda.current_stream().cuda_stream
    rc = lib.brs_forward(C.byref(view), C.byref(g), color_img.data_ptr(), depth_img.data_ptr(), radii.data_ptr(), arena.fn,
                         None, C.byref(state), stream)
    assert rc == 0, lib.brs_error_string(rc)
    torch.cuda.synchronize()
    assert state.num_rendered > 0 and state.geom == arena.buffers[0].data_ptr() and state.image == arena.buffers[2].data_ptr()

    ref = api._C.rasterize_gaussians(*pl.forward_args(scene, cam, bg))
    assert int(ref[0]) == state.num_rendered
    assert torch.equal(ref[1], color_img) and torch.equal(ref[2], depth_img) and torch.equal(ref[3], radii)

    # backward: every gradient tensor is fully written by the library (torch.empty, no zero-fill)
    Wc, _ = synthetic.loss_weights(W, H)
    Wc = Wc.to(DEV).contiguous()
    shapes = {"dL_dmeans2D": (P, 3), "dL_dcolors": (P, 3), "dL_dopacity": (P, 1), "dL_dmeans3D": (P, 3), "dL_dcov3D": (P, 6),
              "dL_dsh": (P, max(M, 1), 3), "dL_dscales": (P, 3), "dL_drotations": (P, 4)}
    out = {k: torch.full(s, float("nan"), device=DEV) for k, s in shapes.items()}
    grads = Grads(*[out[k].data_ptr() if (k != "dL_dsh" or M > 0) else None for k in shapes], 0, 0, None)
    rc = lib.brs_backward(C.byref(view), C.byref(g), radii.data_ptr(), C.byref(state), Wc.data_ptr(), None, C.byref(grads),
                          arena.fn, None, stream)
    assert rc == 0, lib.brs_error_string(rc)
    torch.cuda.synchronize()
    e = torch.Tensor([])
    want = api._C.rasterize_gaussians_backward(
        bg, scene.means3D, radii, scene.colors_precomp if scene.colors_precomp is not None else e, scene.scales,
        scene.rotations, 1.0, e, cam.viewmatrix, cam.projmatrix, cam.tanfovx, cam.tanfovy, Wc, e,
        scene.shs if scene.shs is not None else e, scene.sh_degree, cam.campos, ref[4], int(ref[0]), ref[5], ref[6], False)
    names = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drotations"]
    for k, w in zip(names, want):
        if k == "dL_dsh" and M == 0:
            continue
        assert torch.isfinite(out[k]).all(), k  # no element left unwritten
        assert pl.rel_l2(out[k].reshape(w.shape), w) <= 1e-5, k

    # state mismatch and missing outputs are reported as status codes, not crashes
    bad = FwdState(state.geom, 16, state.binning, state.binning_bytes, state.image, state.image_bytes, state.num_rendered)
    assert lib.brs_backward(C.byref(view), C.byref(g), radii.data_ptr(), C.byref(bad), Wc.data_ptr(), None, C.byref(grads),
                            arena.fn, None, stream) == -5
    no_out = Grads()
    assert lib.brs_backward(C.byref(view), C.byref(g), radii.data_ptr(), C.byref(state), Wc.data_ptr(), None,
                            C.byref(no_out), arena.fn, None, stream) == -1
    # depth gradient requested without the forward's depth image
    need_depth = Grads(*[out[k].data_ptr() if (k != "dL_dsh" or M > 0) else None for k in shapes], 0, 1, None)
    assert lib.brs_backward(C.byref(view), C.byref(g), radii.data_ptr(), C.byref(state), Wc.data_ptr(), Wc.data_ptr(),
                            C.byref(need_depth), arena.fn, None, stream) == -1


def test_filter_and_mark_visible_through_ctypes_only():
    lib = _lib()
    api = pl.ours()
    W, H = 96, 64
    scene = synthetic.make_scene(3000, "band", "precomp", -3.6, seed=32).to(DEV)
    cam = synthetic.yaw_camera(W, H, 0.9).to(DEV)
    view = View(W, H, cam.tanfovx, cam.tanfovy, 1.0, 0, 0, 0, 0, None, _ptr(cam.viewmatrix), _ptr(cam.projmatrix), None)
    stream = torch.cuda.current_stream().cuda_stream
    # scales handed in as a [:, :3] view of a [P, 6] tensor through the row stride (gaussian_renderer/__init__.py:344)
    wide = torch.cat([scene.scales, torch.rand_like(scene.scales)], dim=1).contiguous()
    radii = torch.empty(scene.P, dtype=torch.int32, device=DEV)
    rc = lib.brs_visible_filter(C.byref(view), scene.P, _ptr(scene.means3D), wide.data_ptr(), 6, _ptr(scene.rotations),
                                None, radii.data_ptr(), stream)
    assert rc == 0, lib.brs_error_string(rc)
    rs = synthetic.raster_settings(cam, 0, torch.zeros(3, device=DEV), api.GaussianRasterizationSettings)
    want = api.GaussianRasterizer(rs).visible_filter(scene.means3D, scales=wide[:, :3], rotations=scene.rotations)
    torch.cuda.synchronize()
    assert torch.equal(radii, want) and (radii > 0).any() and (radii == 0).any()
    present = torch.empty(scene.P, dtype=torch.uint8, device=DEV)
    assert lib.brs_mark_visible(scene.P, _ptr(scene.means3D), _ptr(cam.viewmatrix), _ptr(cam.projmatrix),
                                present.data_ptr(), stream) == 0
    torch.cuda.synchronize()
    assert torch.equal(present.bool(), api.GaussianRasterizer(rs).markVisible(scene.means3D))


def test_forward_views_through_ctypes_only():
    """brs_forward_views: a stack of views with DIFFERENT fields of view and scale modifiers as one pipeline, driven
    with plain structs; every view equals a brs_forward of its own, and num_rendered is the sum."""
    lib = _lib()
    W, H = 150, 94
    scene = synthetic.make_scene(6000, "object", "sh2", -3.4, seed=33).to(DEV)
    bg = torch.tensor([0.2, 0.4, 0.1], device=DEV)
    M, P = scene.shs.shape[1], scene.P
    cams = [synthetic.orbit_camera(W, H, 0.5 * k).to(DEV) for k in range(5)]
    mods = [1.0, 0.7, 1.0, 1.6, 0.4]
    fovs = [1.0, 1.2, 0.8, 1.0, 1.5]  # tan-fov factors: the projection matrix stays, the focal length changes
    views = (View * len(cams))(*[
        View(W, H, c.tanfovx * f, c.tanfovy * f, m, scene.sh_degree, M, 0, 0, _ptr(bg), _ptr(c.viewmatrix), _ptr(c.projmatrix),
             _ptr(c.campos)) for c, m, f in zip(cams, mods, fovs)])
    g = Gaussians(P, _ptr(scene.means3D), _ptr(scene.opacities), _ptr(scene.shs), None, _ptr(scene.scales),
                  _ptr(scene.rotations), None)
    n = len(cams)
    color = torch.empty(n, 3, H, W, device=DEV)
    depth = torch.empty(n, 1, H, W, device=DEV)
    radii = torch.empty(n, P, dtype=torch.int32, device=DEV)
    stream = torch.cuda.current_stream().cuda_stream
    arena, total = Arena(), C.c_longlong(-7)
    rc = lib.brs_forward_views(views, n, C.byref(g), color.data_ptr(), depth.data_ptr(), radii.data_ptr(), arena.fn, None,
                               C.byref(total), None, stream)
    assert rc == 0, lib.brs_error_string(rc)
    torch.cuda.synchronize()
    R_sum = 0
    for k in range(n):
        c1 = torch.empty(3, H, W, device=DEV)
        d1 = torch.empty(1, H, W, device=DEV)
        r1 = torch.empty(P, dtype=torch.int32, device=DEV)
        a1, st = Arena(), FwdState()
        one = View(*[getattr(views[k], f) for f, _ in View._fields_])
        assert lib.brs_forward(C.byref(one), C.byref(g), c1.data_ptr(), d1.data_ptr(), r1.data_ptr(), a1.fn, None, C.byref(st),
                               stream) == 0
        torch.cuda.synchronize()
        R_sum += st.num_rendered
        assert torch.equal(r1, radii[k]) and torch.equal(c1, color[k]) and torch.equal(d1, depth[k]), k
    assert total.value == R_sum > 0
    # one view of the stack with another image size: rejected
    views[2].image_width = W + 16
    assert lib.brs_forward_views(views, n, C.byref(g), color.data_ptr(), depth.data_ptr(), radii.data_ptr(), arena.fn, None,
                                 C.byref(total), None, stream) == -1
