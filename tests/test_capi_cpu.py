"""CPU: the C-ABI library loads and exports every symbol include/wsann.h declares; argument
validation and the no-CPU-fallback rule hold without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from rangefilteredann_b200 import capi


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "wsann.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ws_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    L = capi.lib()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/wsann.h but not exported"
    assert L.ws_abi_version() == 1


def _host_index(n=1000, dim=4):
    labels = np.arange(n, dtype=np.float32)
    h = C.c_void_p()
    capi.check(capi.lib().ws_index_create(-1, 0, n, dim, None, capi.ptr(labels), None, 1, C.byref(h)))
    return capi.Handle(h.value, True), labels


def test_no_cpu_fallback():
    idx, labels = _host_index()
    capi.check(capi.lib().ws_index_finalize(idx.raw))
    q = np.zeros((2, 4), np.float32)
    w = np.array([[0, 10], [5, 50]], np.float32)
    ids = np.zeros((2, 10), np.uint32)
    d = np.zeros((2, 10), np.float32)
    with pytest.raises(capi.WsError, match="no CPU fallback"):
        idx.prefilter_batch(q, w, 2, 10, ids, d)
    qp = capi.query_params()
    with pytest.raises(capi.WsError):
        idx.tree_batch("fenwick", q, w, 2, qp, ids, d)


def test_create_validation():
    L = capi.lib()
    h = C.c_void_p()
    labels = np.array([3, 2, 1], np.float32)
    assert L.ws_index_create(-1, 0, 3, 4, None, capi.ptr(labels), None, 1, C.byref(h)) == -1  # unsorted labels
    assert b"not sorted" in L.ws_last_error()
    assert L.ws_index_create(-1, 7, 3, 4, None, capi.ptr(labels), None, 0, C.byref(h)) == -1  # bad metric
    assert L.ws_index_create(-1, 0, 0, 4, None, capi.ptr(labels), None, 0, C.byref(h)) == -1  # empty
    v = np.zeros((3, 4), np.float32)
    n = C.c_int()
    L.ws_device_count(C.byref(n))
    rc = L.ws_index_create(0, 0, 3, 4, capi.ptr(v), capi.ptr(np.sort(labels)), None, 1, C.byref(h))
    if n.value == 0:
        assert rc == -2 and b"no CPU fallback" in L.ws_last_error()  # WS_ERR_CUDA, loudly
    else:
        assert rc == 0
        L.ws_index_destroy(h)


def test_engine_module_surface(engine):
    # every index class x variant the reference registers (python_bindings.cpp:111-157,231-236)
    for sfx in ("FloatEuclidian", "FloatMips", "UInt8Euclidian", "UInt8Mips", "Int8Euclidian", "Int8Mips"):
        for cls in ("PrefilterIndex", "PostfilterVamanaIndex", "RangeFilterTreeIndex", "VamanaRangeFilterTreeIndex",
                    "SuperOptimizedPostfilterTreeIndex"):
            assert hasattr(engine, cls + sfx)
    qp = engine.QueryParams(10, 20, 1.35, 10_000_000, 10_000, 1, 10000, None, False)
    assert qp is not None
    assert engine.BuildParams(64, 500, 1.0, "x/") is not None
    with pytest.raises(RuntimeError, match="2-dimensional"):
        engine.PrefilterIndexFloatEuclidian(np.zeros(4, np.float32), np.zeros(4, np.float32))
    # 8-bit variants: dimensions beyond the exactly-representable integer range are refused up front
    with pytest.raises(RuntimeError, match="8-bit variants are supported up to 258"):
        engine.PrefilterIndexUInt8Euclidian(np.zeros((4, 300), np.uint8), np.arange(4, dtype=np.float32))
    with pytest.raises(RuntimeError, match="up to 1023"):
        engine.PrefilterIndexInt8Mips(np.zeros((4, 1024), np.int8), np.arange(4, dtype=np.float32))


def test_group_and_snapshot_entry_points_validate_arguments(tmp_path):
    """ws_group / NCCL / snapshot entry points reject bad arguments without a device (no compute)."""
    L = capi.lib()
    out = C.c_void_p()
    assert L.ws_group_create(None, 0, 0, C.byref(out)) == -1 and out.value is None
    assert L.ws_group_tree_batch(None, 0, None, None, 0, None, None, None) == -1
    assert L.ws_index_load(str(tmp_path / "missing.wsann").encode(), 0, C.byref(out)) == -1
    assert b"cannot open" in L.ws_last_error()
    bad = tmp_path / "bad.wsann"
    bad.write_bytes(b"garbage garbage garbage")
    assert L.ws_index_load(str(bad).encode(), 0, C.byref(out)) == -1
    assert b"not an arena snapshot" in L.ws_last_error()
    assert L.ws_index_comm_init(None, 2, 0, None) == -1
    assert L.ws_allgather_merge(None, None, None, 0, 10, 0, None, None) == -1


def test_staging_helper_pool_copies_correctly_and_the_process_exits():
    """The helper threads that stage host batches into pinned memory (csrc/wsann.cu WsCopyPool): hundreds of jobs,
    with pauses that let the helpers fall asleep, in a child process that must also EXIT (a pool with waiting threads
    must not block interpreter shutdown)."""
    import subprocess
    import sys
    code = ("import ctypes as C, sys; sys.path.insert(0, %r); from rangefilteredann_b200 import capi; L = capi.lib(); "
            "L.ws_debug_copy_pool_selftest.argtypes = [C.c_uint64, C.c_uint32]; "
            "rc = L.ws_debug_copy_pool_selftest(5_300_000, 300); rc2 = L.ws_debug_copy_pool_selftest(100, 20); print('rc', rc, rc2)") % ROOT
    for threads in ("3", "0", "8"):
        env = dict(os.environ, WSANN_COPY_THREADS=threads)
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120, env=env)
        assert out.returncode == 0 and "rc 0 0" in out.stdout, (threads, out.stdout, out.stderr)
