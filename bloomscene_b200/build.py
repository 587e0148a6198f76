"""In-tree build of the native code (no JIT cache: the built .so files travel with the repo).

  bloomscene_b200/libbloomrast.so  <- csrc/*.cu   nvcc, sm_100a only, no torch headers (C-ABI, include/bloomrast.h)
  bloomscene_b200/_C.so            <- csrc/torch_ext.cpp   g++ only, links libbloomrast.so + libtorch

Floating-point flags are nvcc's defaults on purpose (-fmad=true, IEEE div/sqrt, no fast-math): the
integer outputs must be bit-identical to the reference, which is built with the same defaults.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
BUILD = PKG / "_build"
LIB = PKG / "libbloomrast.so"
EXT = PKG / "_C.so"

CU_SOURCES = ["preprocess.cu", "binning.cu", "blend_fwd.cu", "blend_bwd.cu", "preprocess_bwd.cu", "loss.cu", "neural.cu", "measure.cu", "api.cu"]
NVCC_FLAGS = [
    "-std=c++17", "-O3",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "--compiler-options", "-fPIC",
    "-Xptxas", "-v",
]


def _hash(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()[:16]


def _native_inputs():
    return list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [ROOT / "include" / "bloomrast.h"]


def build_lib(verbose: bool = False) -> Path:
    BUILD.mkdir(exist_ok=True)
    stamp = BUILD / "lib.stamp"
    fp = _hash(_native_inputs())
    if LIB.exists() and stamp.exists() and stamp.read_text() == fp:
        return LIB

    def compile_one(src: str) -> Path:
        obj = BUILD / (src + ".o")
        log = BUILD / (src + ".log")
        cmd = ["nvcc", *NVCC_FLAGS, f"-I{ROOT / 'include'}", "-c", str(CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log.write_text(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(f"[build] nvcc {src}", flush=True)
        return obj

    with ThreadPoolExecutor(max_workers=len(CU_SOURCES)) as ex:
        objs = list(ex.map(compile_one, CU_SOURCES))
    cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", *map(str, objs), "-o", str(LIB),
           "-cudart", "shared"]
    subprocess.run(cmd, check=True)
    stamp.write_text(fp)
    return LIB


def build_ext(verbose: bool = False) -> Path:
    import torch
    from torch.utils import cpp_extension as ce

    BUILD.mkdir(exist_ok=True)
    stamp = BUILD / "ext.stamp"
    fp = _hash([CSRC / "torch_ext.cpp", ROOT / "include" / "bloomrast.h"]) + torch.__version__
    if EXT.exists() and stamp.exists() and stamp.read_text() == fp:
        return EXT
    try:
        inc = ce.include_paths(device_type="cuda")
    except TypeError:
        inc = ce.include_paths(True)
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    obj = BUILD / "torch_ext.o"
    cmd = [
        "g++", "-std=c++17", "-O2", "-fPIC", "-DTORCH_EXTENSION_NAME=_C", "-DTORCH_API_INCLUDE_EXTENSION_H",
        "-D_GLIBCXX_USE_CXX11_ABI=1",
        *[f"-I{p}" for p in inc], f"-I{sysconfig.get_paths()['include']}", f"-I{ROOT / 'include'}",
        "-c", str(CSRC / "torch_ext.cpp"), "-o", str(obj),
    ]
    if verbose:
        print("[build] g++ torch_ext.cpp", flush=True)
    subprocess.run(cmd, check=True)
    link = [
        "g++", "-shared", str(obj), "-o", str(EXT),
        f"-L{PKG}", "-lbloomrast", f"-L{torch_lib}", "-L/usr/local/cuda/lib64",
        "-lc10", "-ltorch", "-ltorch_cpu", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda", "-lcudart",
        "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{torch_lib}", "-Wl,-rpath,/usr/local/cuda/lib64",
    ]
    subprocess.run(link, check=True)
    stamp.write_text(fp)
    return EXT


def build_tools(verbose: bool = False):
    """Stand-alone measurement programs under tools/ (they link libbloomrast.so through the C-ABI only)."""
    out = []
    for name in ("sort_vs_cub", "probe_ffma2"):
        src = ROOT / "tools" / f"{name}.cu"
        exe = BUILD / name
        if not src.exists() or (exe.exists() and exe.stat().st_mtime > max(src.stat().st_mtime, LIB.stat().st_mtime)):
            continue
        cmd = ["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", str(src), f"-I{ROOT / 'include'}",
               f"-L{PKG}", "-lbloomrast", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/..", "-o", str(exe)]
        if verbose:
            print(f"[build] nvcc tools/{name}.cu", flush=True)
        subprocess.run(cmd, check=True)
        out.append(exe)
    return out


def build_all(verbose: bool = False):
    lib = build_lib(verbose)
    ext = build_ext(verbose)
    build_tools(verbose)
    return lib, ext


if __name__ == "__main__":
    print(build_all(verbose=True))
    sys.exit(0)
