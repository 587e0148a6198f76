"""Build the reference's OWN CUDA rasterizer, unmodified, into oracle/_ref/ (test infrastructure).

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (bloomscene_b200/, the drop-in packages)
imports, links or executes anything under oracle/.  Only tests/, __graft_entry__.smoke() and
bench.py's reference / cpu_baseline legs may.

What this does: compiles, with nvcc for sm_100a, the reference sources *where they lie* under
/root/reference/submodules/depth-diff-gaussian-rasterization (rasterizer_impl.cu, forward.cu,
backward.cu, rasterize_points.cu, ext.cpp) into one Python extension `oracle/_ref/_ref_C.so`.
No reference source is copied into this repo; outputs go only into oracle/_ref/ (git-ignored, but
NOT gpurun-ignored, so the .so travels to the GPU box where /root/reference does not exist).

Two aids that do not touch the reference sources:
  * -I oracle/glm_shim : GLM is an un-vendored submodule in the reference (setup.py:29 points at
    third_party/glm, which is absent), so a header restating the GLM subset it uses is supplied.
  * -include cstdint   : gcc 13 no longer pulls <cstdint> in transitively
    (cuda_rasterizer/rasterizer_impl.h:24,40 use std::uintptr_t / uint32_t).

Flags are those torch's BuildExtension would pass for `python setup.py install` of the reference
(setup.py passes no arch / fast-math flags): nvcc defaults -O3, -fmad=true, -prec-div=true,
-prec-sqrt=true, -ftz=false.  The module name is `_ref_C` so it can never shadow the product `_C`.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_RAST = Path("/root/reference/submodules/depth-diff-gaussian-rasterization")
OUT = HERE / "_ref"
MODULE = "_ref_C"
SOURCES = [
    "cuda_rasterizer/rasterizer_impl.cu",
    "cuda_rasterizer/forward.cu",
    "cuda_rasterizer/backward.cu",
    "rasterize_points.cu",
    "ext.cpp",
]


def _torch_paths():
    import torch
    from torch.utils import cpp_extension as ce

    inc = ce.include_paths("cuda") if hasattr(ce, "include_paths") else []
    try:
        inc = ce.include_paths(device_type="cuda")
    except TypeError:
        pass
    lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    return inc, lib


def _fingerprint() -> str:
    h = hashlib.sha256()
    for s in SOURCES:
        h.update((REF_RAST / s).read_bytes())
    for extra in sorted((REF_RAST / "cuda_rasterizer").glob("*.h")):
        h.update(extra.read_bytes())
    h.update((HERE / "glm_shim/glm/glm.hpp").read_bytes())
    h.update(Path(__file__).read_bytes())
    return h.hexdigest()[:16]


def so_path() -> Path:
    return OUT / f"{MODULE}.so"


PKG = "ref_depth_diff_gaussian_rasterization"


def write_reference_package() -> Path | None:
    """Place the reference's OWN Python wrapper (depth_diff_gaussian_rasterization/__init__.py, every
    line as it is in /root/reference) into the git-ignored oracle/_ref/pkg/, with its one import line
    `from . import _C` redirected to oracle/_ref/_ref_C.so.  bench.py --impl reference imports this
    package and nothing of the product, so the reference arm runs the reference's stock code path."""
    src = REF_RAST / "depth_diff_gaussian_rasterization" / "__init__.py"
    dst = OUT / "pkg" / PKG / "__init__.py"
    if not src.exists():
        return dst if dst.exists() else None
    text = src.read_text()
    needle = "from . import _C\n"
    assert text.count(needle) == 1, "reference wrapper changed: expected exactly one `from . import _C`"
    loader = (
        "import importlib.util as _ilu, os as _os\n"
        f"_spec = _ilu.spec_from_file_location('{MODULE}', _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), '..', '..', '{MODULE}.so'))\n"
        "_C = _ilu.module_from_spec(_spec)\n"
        "_spec.loader.exec_module(_C)\n"
    )
    dst.parent.mkdir(parents=True, exist_ok=True)
    dst.write_text(text.replace(needle, loader))
    return dst


def build(verbose: bool = True) -> Path | None:
    """Build oracle/_ref/_ref_C.so if /root/reference is present; return its path (or None)."""
    if not REF_RAST.exists():
        # GPU box: only the prebuilt file is used.
        return so_path() if so_path().exists() else None
    OUT.mkdir(parents=True, exist_ok=True)
    stamp = OUT / "stamp.txt"
    fp = _fingerprint()
    write_reference_package()
    if so_path().exists() and stamp.exists() and stamp.read_text().strip() == fp:
        return so_path()

    inc, torch_lib = _torch_paths()
    pyinc = sysconfig.get_paths()["include"]
    common = [
        f"-I{HERE / 'glm_shim'}",
        f"-I{REF_RAST}",
        f"-I{REF_RAST / 'cuda_rasterizer'}",
        *[f"-I{p}" for p in inc],
        f"-I{pyinc}",
        f"-DTORCH_EXTENSION_NAME={MODULE}",
        "-DTORCH_API_INCLUDE_EXTENSION_H",
        "-D_GLIBCXX_USE_CXX11_ABI=1",
        "-std=c++17",
    ]
    nvcc_flags = [
        "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__",
        "-D__CUDA_NO_BFLOAT16_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__",
        "--expt-relaxed-constexpr",
        "-gencode", "arch=compute_100a,code=sm_100a",
        "--compiler-options", "-fPIC",
        "-include", "cstdint",
    ]
    objdir = OUT / "obj"
    objdir.mkdir(exist_ok=True)

    def compile_one(src: str) -> Path:
        obj = objdir / (src.replace("/", "_") + ".o")
        if src.endswith(".cu"):
            cmd = ["nvcc", *common, *nvcc_flags, "-c", str(REF_RAST / src), "-o", str(obj)]
        else:
            cmd = ["g++", *common, "-O2", "-fPIC", "-include", "cstdint", "-c", str(REF_RAST / src), "-o", str(obj)]
        if verbose:
            print("[oracle/_ref]", " ".join(cmd[:1]), src, flush=True)
        subprocess.run(cmd, check=True)
        return obj

    with ThreadPoolExecutor(max_workers=5) as ex:
        objs = list(ex.map(compile_one, SOURCES))

    link = [
        "g++", "-shared", *map(str, objs), "-o", str(so_path()),
        f"-L{torch_lib}", "-L/usr/local/cuda/lib64",
        "-lc10", "-ltorch", "-ltorch_cpu", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda", "-lcudart",
        f"-Wl,-rpath,{torch_lib}", "-Wl,-rpath,/usr/local/cuda/lib64",
    ]
    subprocess.run(link, check=True)
    stamp.write_text(fp)
    if verbose:
        print("[oracle/_ref] built", so_path(), flush=True)
    return so_path()


def load_reference_package():
    """Import the reference's own Python package from oracle/_ref/pkg (see write_reference_package);
    None if it was never built."""
    init = OUT / "pkg" / PKG / "__init__.py"
    if not init.exists() or not so_path().exists():
        return None
    import importlib.util

    import torch  # noqa: F401

    spec = importlib.util.spec_from_file_location(PKG, str(init), submodule_search_locations=[str(init.parent)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[PKG] = mod
    spec.loader.exec_module(mod)
    return mod


def load():
    """Import the built reference extension (needs `import torch` first). Returns the module or None."""
    p = so_path()
    if not p.exists():
        return None
    import importlib.util

    import torch  # noqa: F401  (registers libtorch symbols before dlopen)

    spec = importlib.util.spec_from_file_location(MODULE, str(p))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build()
    print(p)
    sys.exit(0 if p else 1)
