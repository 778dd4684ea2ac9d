"""Builds the two native artefacts of the package, in-tree:

  libwsann_cuda.so                 nvcc, sm_100a only  (csrc/wsann.cu: arena + orchestration + C ABI; csrc/ws_k_*.cu:
                                   kernel families, one object per metric / beam capacity; csrc/ws_gemm.cu: the
                                   tcgen05 prefilter kernels — separate objects under build/, compiled in parallel)
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
              "-diag-suppress", "177", "-shared", "-Xcompiler", "-fPIC"]


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


# (object name, source, extra defines): one object each, compiled in parallel, linked into LIB.  The templated
# kernel families are instantiated per metric (and the warp beam kernel per beam capacity) so that no single
# translation unit dominates the build.
CUDA_UNITS = (
    [("wsann", "wsann.cu", []), ("ws_gemm", "ws_gemm.cu", []), ("ws_k_misc", "ws_k_misc.cu", [])]
    + [(f"ws_k_beam_warp_m{m}_{cs}", "ws_k_beam_warp.cu", [f"-DWSK_METRIC={m}", f"-DWSK_CS={cs}"])
       for m in (0, 1) for cs in (7, 8, 9, 10)]
    + [(f"{name}_m{m}", f"{name}.cu", [f"-DWSK_METRIC={m}"])
       for name in ("ws_k_beam_cta", "ws_k_scan", "ws_k_build") for m in (0, 1)]
)
OBJ_DIR = os.path.join(HERE, "build")


def build_cuda(force: bool = False, verbose: bool = False, jobs: int | None = None) -> str:
    srcs = _sources(CSRC, os.path.join(os.path.dirname(HERE), "include"))
    headers = [s for s in srcs if not s.endswith(".cu")]
    lib_out = os.environ.get("WSANN_LIB_OUT", LIB)
    obj_dir = os.environ.get("WSANN_OBJ_DIR", OBJ_DIR)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("WSANN_NVCC_EXTRA", "").split()
    jobs = jobs or int(os.environ.get("WSANN_BUILD_JOBS", os.cpu_count() or 4))
    os.makedirs(obj_dir, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "-shared"]
    pending, objs = [], []
    for name, unit, defines in CUDA_UNITS:
        src = os.path.join(CSRC, unit)
        obj = os.path.join(obj_dir, name + ".o")
        objs.append(obj)
        if not force and not extra and _newer(obj, [src, *headers]):
            continue
        cmd = [nvcc, *flags, *defines, *extra, "-c", "-o", obj, src]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        pending.append(cmd)
    compiled = bool(pending)
    running = []
    while pending or running:
        while pending and len(running) < jobs:
            cmd = pending.pop(0)
            running.append((cmd, subprocess.Popen(cmd)))
        cmd, p = running.pop(0)
        if p.wait() != 0:
            for _, q in running:
                q.kill()
            raise subprocess.CalledProcessError(p.returncode, cmd)
    if compiled or force or not _newer(lib_out, objs):
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
