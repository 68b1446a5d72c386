"""Build recipes for the native parts of amcl3d_b200 (run HERE without a GPU: nvcc cross-compiles sm_100a).

* ``build_cuda()`` -> amcl3d_b200/lib/libamcl3d_cuda.so  : the CUDA kernels + the C-ABI of include/amcl3d_cuda.h
* ``build_host()`` -> amcl3d_b200/lib/libamcl3d_host.so  : the host-side C++ classes (Grid3d, ParticleFilter,
  PointCloudTools with the reference's class API) + the extern "C" test harness, linked against the CUDA library

Both are built in-tree so that they travel to the GPU box with the repository snapshot.
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "amcl3d_b200")
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")
LIB = os.path.join(PKG, "lib")

CUDA_SOURCES = ["api.cu", "weight.cu", "filter.cu", "filter_exact.cu", "cloud.cu", "distance_field.cu", "comm.cu", "probe.cu", "order.cu", "voxel_grid.cu"]
HOST_SOURCES = ["Grid3d.cpp", "ParticleFilter.cpp", "PointCloudTools.cpp"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",          # never contract a*b+c: the reference's float expressions are not fused
    "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _host_cxx():
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    return r.stdout + r.stderr


def cuda_lib_path():
    return os.path.join(LIB, "libamcl3d_cuda.so")


def host_lib_path():
    return os.path.join(LIB, "libamcl3d_host.so")


def build_cuda(force=False, verbose=False, extra_flags=()):
    os.makedirs(LIB, exist_ok=True)
    out = cuda_lib_path()
    srcs = [os.path.join(CSRC, s) for s in CUDA_SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith((".cuh", ".h"))] + [
        os.path.join(ROOT, "include", "amcl3d_cuda.h")]
    if not force and not _stale(out, deps):
        return out
    objs = []
    for s in srcs:
        o = os.path.join(LIB, os.path.basename(s) + ".o")
        if force or _stale(o, deps):
            _run([_nvcc(), "-ccbin", _host_cxx()] + NVCC_FLAGS + list(extra_flags) + ["-c", s, "-o", o], verbose)
        objs.append(o)
    _run([_nvcc(), "-ccbin", _host_cxx(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC",
          "-o", out] + objs + ["-ldl"], verbose)
    return out


def build_host(force=False, verbose=False):
    os.makedirs(LIB, exist_ok=True)
    out = host_lib_path()
    cuda = build_cuda(force=False, verbose=verbose)
    srcs = [os.path.join(HOST, s) for s in HOST_SOURCES] + [os.path.join(ROOT, "tests", "harness", "class_harness.cpp")]
    deps = srcs + [os.path.join(HOST, h) for h in os.listdir(HOST) if h.endswith(".h")] + [cuda]
    if not force and not _stale(out, deps):
        return out
    cmd = [_host_cxx(), "-std=c++11", "-O2", "-fPIC", "-shared", "-Wall",
           "-I" + os.path.join(HOST, "compat"), "-I" + HOST, "-I" + os.path.join(ROOT, "include")] + srcs + [
        "-L" + LIB, "-lamcl3d_cuda", "-Wl,-rpath,$ORIGIN", "-o", out]
    _run(cmd, verbose)
    return out


def build_all(force=False, verbose=False):
    paths = [build_cuda(force=force, verbose=verbose)]
    if all(os.path.exists(os.path.join(HOST, s)) for s in HOST_SOURCES):
        paths.append(build_host(force=force, verbose=verbose))
    return paths


if __name__ == "__main__":
    for p in build_all(force="--force" in sys.argv, verbose=True):
        print("built", p)
