"""rangefilteredann_b200 — B200-native window-search engine behind RangeFilteredANN's
`window_ann` boundary.

Layout (only what the query hot path needs):
  csrc/wsann.cu, ws_kernels.cuh, ws_device.cuh, ws_decompose.h   sm_100a kernels + C ABI
  csrc/host/window_index.hpp, python_bindings.cpp                 host index classes + pybind11
  wrapper.py    mirror of the reference's experiments/wrapper.py helpers
  capi.py       ctypes view of include/wsann.h (tests / bench plumbing)
  synth.py      synthetic datasets of the BASELINE.json shapes
  build.py      in-tree build of the two shared objects

There is no CPU fallback: importing the engine without its compiled extension raises.
"""
from __future__ import annotations

import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_ENGINE = None


def _ext_path() -> str | None:
    for f in sorted(os.listdir(_HERE)):
        if f.startswith("_window_ann_b200") and f.endswith(".so"):
            return os.path.join(_HERE, f)
    return None


def load_engine():
    """Returns this repo's compiled `window_ann` module. Raises (loudly) when the CUDA
    extension has not been built — there is nothing to fall back to."""
    global _ENGINE
    if _ENGINE is not None:
        return _ENGINE
    path = _ext_path()
    lib = os.path.join(_HERE, "libwsann_cuda.so")
    if path is None or not os.path.exists(lib):
        raise ImportError(
            "rangefilteredann_b200: native extension missing (libwsann_cuda.so / _window_ann_b200*.so). "
            "Run `python -m rangefilteredann_b200.build` (needs nvcc); there is no CPU fallback.")
    # its own extension name: it must be able to coexist with the reference's `window_ann`
    # extension in one interpreter (tests and bench load both)
    spec = importlib.util.spec_from_file_location("_window_ann_b200", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _ENGINE = mod
    return mod
