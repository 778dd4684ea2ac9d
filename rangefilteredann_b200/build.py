"""Builds the two native artefacts of the package, in-tree:

  libwsann_cuda.so                 nvcc, sm_100a only  (csrc/wsann.cu: kernels + C ABI; csrc/ws_gemm.cu: the
                                   tcgen05 prefilter kernels — separate objects under build/)
  _window_ann_b200.cpython-*.so    g++ + pybind11      (csrc/host/python_bindings.cpp; re-exported as
                                   `window_ann` by window_ann.py)

Both land next to this file so they travel with the repo snapshot.  nvcc cross-compiles
without a GPU.
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libwsann_cuda.so")
EXT = os.path.join(HERE, "_window_ann_b200" + sysconfig.get_config_var("EXT_SUFFIX"))

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _sources(*dirs: str) -> list[str]:
    out = []
    for d in dirs:
        for root, _, files in os.walk(d):
            out += [os.path.join(root, f) for f in files if f.endswith((".cu", ".cuh", ".h", ".hpp", ".cpp"))]
    return out


CUDA_UNITS = ("wsann.cu", "ws_gemm.cu")  # one object each, compiled in parallel, linked into LIB
OBJ_DIR = os.path.join(HERE, "build")


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    srcs = _sources(CSRC, os.path.join(os.path.dirname(HERE), "include"))
    headers = [s for s in srcs if not s.endswith(".cu")]
    lib_out = os.environ.get("WSANN_LIB_OUT", LIB)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("WSANN_NVCC_EXTRA", "").split()
    os.makedirs(OBJ_DIR, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "-shared"]
    procs, objs = [], []
    for unit in CUDA_UNITS:
        src = os.path.join(CSRC, unit)
        obj = os.path.join(OBJ_DIR, unit.replace(".cu", ".o"))
        objs.append(obj)
        if not force and not extra and _newer(obj, [src, *headers]):
            continue
        cmd = [nvcc, *flags, *extra, "-c", "-o", obj, src]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((cmd, subprocess.Popen(cmd)))
    for cmd, p in procs:
        if p.wait() != 0:
            raise subprocess.CalledProcessError(p.returncode, cmd)
    if procs or force or not _newer(lib_out, objs):
        subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib_out, *objs], check=True)
    return LIB


def build_bindings(force: bool = False) -> str:
    import pybind11
    srcs = _sources(os.path.join(CSRC, "host"), os.path.join(os.path.dirname(HERE), "include")) + [LIB]
    if not force and _newer(EXT, srcs):
        return EXT
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden",
           "-I" + pybind11.get_include(), "-I" + sysconfig.get_paths()["include"],
           os.path.join(CSRC, "host", "python_bindings.cpp"), "-o", EXT,
           "-L" + HERE, "-lwsann_cuda", "-Wl,-rpath,$ORIGIN"]
    subprocess.run(cmd, check=True)
    return EXT


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_cuda(force, verbose)
    build_bindings(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built", LIB, EXT)
