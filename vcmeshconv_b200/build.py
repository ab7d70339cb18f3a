"""In-tree build of the two shared libraries (no JIT cache, no pip install).

  libvcmesh.so           hand-written sm_100a CUDA kernels behind include/vcmesh.h
  libvcmesh_sampling.so  host C++ GraphSampling behind include/vcmesh_sampling.h

nvcc cross-compiles for sm_100a without a GPU; the built .so files are git-
ignored but travel to the GPU box with the gpurun snapshot.
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_CUDA = os.path.join(PKG, "libvcmesh.so")
LIB_SAMPLING = os.path.join(PKG, "libvcmesh_sampling.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
    "-shared",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("build failed: " + cmd[0])
    return r.stdout + r.stderr


def cuda_sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def cuda_deps():
    hdr = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdr += [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
    return cuda_sources() + hdr


def build_sampling(force=False):
    src = os.path.join(CSRC, "graph_sampling.cpp")
    deps = [src, os.path.join(ROOT, "include", "vcmesh_sampling.h")]
    if force or _newer(LIB_SAMPLING, deps):
        _run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden",
              "-DVCM_BUILDING", src, "-o", LIB_SAMPLING])
    return LIB_SAMPLING


def build_cuda(force=False, verbose=False):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if force or _newer(LIB_CUDA, cuda_deps()):
        if not os.path.exists(nvcc):
            raise RuntimeError("nvcc not found and %s is missing or stale" % LIB_CUDA)
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
              ["-I", os.path.join(ROOT, "include"), "-DVCM_BUILDING"] + cuda_sources() + \
              ["-o", LIB_CUDA, "-lcuda"]
        out = _run(cmd)
        if verbose:
            print(out)
    return LIB_CUDA


def build_all(force=False, verbose=False):
    return build_sampling(force), build_cuda(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built", LIB_SAMPLING, LIB_CUDA)
