"""Builds ``torch_renderer_b200/libtrb.so`` in-tree with nvcc for sm_100a.

The library is a plain C-ABI shared object (``include/trb.h``): no torch, no pybind11, so a full
rebuild is a few seconds.  ``python -m torch_renderer_b200.build`` or ``__graft_entry__.build()``.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
_CSRC = os.path.join(_PKG, "csrc")
# TRB_BUILD_TAG=<tag> builds a variant beside the product library (libtrb_<tag>.so, objects in build_<tag>/):
# diagnostic builds (-DTRB_KN_STATS) and same-box A/B of kernel variants, selected at run time with TRB_LIB_PATH
_TAG = os.environ.get("TRB_BUILD_TAG", "")
_OBJ = os.path.join(_PKG, "build" + (f"_{_TAG}" if _TAG else ""))
LIB_PATH = os.path.join(_PKG, f"libtrb_{_TAG}.so" if _TAG else "libtrb.so")

SOURCES = ["api.cu", "raster.cu", "shade.cu", "transform.cu", "render.cu", "render_kn.cu", "render_stages.cu", "allreduce.cu", "points.cu", "clip.cu", "points_render.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libtrb.so cannot be built")


def _host_compiler_flags():
    # the image exports CC/CXX pointing at a gcc wrapper; use the system g++ explicitly
    for cand in ("/usr/bin/g++", shutil.which("g++")):
        if cand and os.path.exists(cand):
            return ["-ccbin", cand]
    return []


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, ptxas_info: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(_OBJ, exist_ok=True)
    headers = [os.path.join(_CSRC, h) for h in os.listdir(_CSRC) if h.endswith((".cuh", ".h"))]
    headers.append(os.path.join(_ROOT, "include", "trb.h"))
    headers.append(os.path.abspath(__file__))
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(_CSRC, s))]
    flags = list(NVCC_FLAGS) + os.environ.get("TRB_EXTRA_NVCC_FLAGS", "").split()  # e.g. -DTRB_KN_STATS (diagnostics)
    if ptxas_info:
        flags += ["-Xptxas", "-v"]

    def compile_one(src):
        obj = os.path.join(_OBJ, src.replace(".cu", ".o"))
        path = os.path.join(_CSRC, src)
        if force or _stale(obj, [path] + headers):
            cmd = [nvcc, *ARCH, *flags, *_host_compiler_flags(), "-I", os.path.join(_ROOT, "include"),
                   "-I", _CSRC, "-c", path, "-o", obj]
            if verbose:
                print(" ".join(cmd), flush=True)
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
            if (verbose or ptxas_info) and (r.stdout or r.stderr):
                print(r.stdout + r.stderr, flush=True)
            return obj, True
        return obj, False

    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(compile_one, srcs))
    objs = [o for o, _ in results]
    if force or any(changed for _, changed in results) or _stale(LIB_PATH, objs):
        cmd = [nvcc, *ARCH, "-shared", *_host_compiler_flags(), "-o", LIB_PATH, *objs]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, ptxas_info="--ptxas" in sys.argv))
