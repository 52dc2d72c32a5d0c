"""Build libsd_fusion.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

``python -m semantic_depth_b200.build`` or ``__graft_entry__.build()``.  The library is written to
``semantic_depth_b200/libsd_fusion.so`` so that it travels with the repo snapshot to the GPU box.
-fmad=false is part of the numerical contract (see csrc/sd_common.cuh).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsd_fusion.so")
OPS_LIB = os.path.join(HERE, "libsd_torch_ops.so")       # TORCH_LIBRARY op layer over the C ABI (host C++ only)
OPS_SRC = os.path.join(CSRC, "sd_torch_ops.cpp")
SOURCES = ["sd_api.cu", "sd_pixel.cu", "sd_select.cu", "sd_compact.cu", "sd_plane.cu", "sd_knn.cu", "sd_ransac.cu", "sd_resize.cu", "sd_ply.cu", "sd_overlay.cu", "sd_fcn_head.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(OPS_LIB):
        return True
    t = min(os.path.getmtime(LIB), os.path.getmtime(OPS_LIB))
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "sd_fusion.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *os.environ.get("SD_EXTRA_NVCC_FLAGS", "").split(), "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {src}\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-lcudart"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    build_ops(verbose)
    return LIB


def build_ops(verbose: bool = False) -> str:
    """libsd_torch_ops.so: csrc/sd_torch_ops.cpp against this interpreter's PyTorch, linked to libsd_fusion.so next to it."""
    import torch
    from torch.utils.cpp_extension import include_paths, library_paths
    cxx = os.environ.get("CXX") or shutil.which("g++") or "g++"
    cuda_inc = os.path.join(os.path.dirname(os.path.dirname(_nvcc())), "include")
    cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
           "-I", os.path.join(HERE, "..", "include"), "-I", cuda_inc]
    for inc in include_paths():
        cmd += ["-isystem", inc]
    cmd += [OPS_SRC, "-o", OPS_LIB]
    for lp in library_paths():
        cmd += ["-L", lp, f"-Wl,-rpath,{lp}"]
    cmd += ["-L", HERE, "-lsd_fusion", "-Wl,-rpath,$ORIGIN", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0 or verbose:
        sys.stderr.write(f"--- {' '.join(cmd)}\n{r.stdout}\n")
    if r.returncode != 0:
        raise RuntimeError("building libsd_torch_ops.so failed")
    return OPS_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
