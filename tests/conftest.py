import contextlib
import importlib.util
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
DATA_CACHE = os.path.join(ROOT, "data_cache")
REF_CACHE = os.path.join(ROOT, "ref_cache")  # graph caches written by the UNMODIFIED reference builder (oracle/build_ref_cache.py)

# A graph-cache miss must never be answered by the device-side builder behind a test's back: it would write
# engine-built graphs into the committed golden directories under the reference's file names and turn "searched on
# identical reference graphs" into self-comparison.  Tests that WANT a device build say so (device_graph_build()).
os.environ["WSANN_GRAPH_BUILD"] = "0"


@contextlib.contextmanager
def device_graph_build():
    old = os.environ.get("WSANN_GRAPH_BUILD")
    os.environ["WSANN_GRAPH_BUILD"] = "1"
    try:
        yield
    finally:
        if old is None:
            os.environ.pop("WSANN_GRAPH_BUILD", None)
        else:
            os.environ["WSANN_GRAPH_BUILD"] = old


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _load_ext(path, name="window_ann"):
    """Load a pybind11 module called `window_ann` from an explicit file without leaving it
    in sys.modules (the reference and this engine both use that module name)."""
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules.pop(name, None)
    return mod


def find_ext(dirname):
    if not os.path.isdir(dirname):
        return None
    for f in sorted(os.listdir(dirname)):
        if f.startswith("window_ann") and f.endswith(".so"):
            return os.path.join(dirname, f)
    return None


@pytest.fixture(scope="session")
def engine():
    """This repo's window_ann module (C++ host classes over libwsann_cuda.so)."""
    from rangefilteredann_b200 import load_engine
    return load_engine()


@pytest.fixture(scope="session")
def ref():
    """The UNMODIFIED reference module compiled into oracle/_ref (checker only)."""
    path = find_ext(os.path.join(ROOT, "oracle", "_ref"))
    if path is None:
        pytest.skip("oracle/_ref not built")
    os.environ.setdefault("PARLAY_NUM_THREADS", str(os.cpu_count()))
    try:
        return _load_ext(path)
    except Exception as e:  # e.g. built for another CPU
        pytest.skip(f"oracle/_ref not loadable here: {e}")


def has_gpu():
    try:
        from rangefilteredann_b200 import load_engine
        return load_engine().device_count() > 0
    except Exception:
        return False
