"""ctypes loader for oracle/libwsann_oracle.so (TEST INFRASTRUCTURE — the checker)."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "libwsann_oracle.so")
KINDS = {"prefilter": 0, "flat": 1, "wst": 2, "super": 3, "pretree": 4}
METHODS = {"fenwick": 0, "optimized_postfilter": 1, "three_split": 2, "super": 3, "prefilter": 10, "flat": 11}

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB)
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p,
                                    C.c_int32, C.c_float, C.c_float, C.c_long, C.c_long, C.c_double, C.c_char_p]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_long, C.c_long,
                                   C.c_long, C.c_long, C.c_int, C.c_float, C.c_uint32, C.c_int, C.c_void_p,
                                   C.c_void_p, C.c_void_p]
        L.oracle_decompose.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_long, C.c_long, C.c_int,
                                       C.c_float, C.c_uint32, C.c_void_p]
        _lib = L
    return _lib


class Oracle:
    def __init__(self, kind, data, labels, cache_path, metric=0, dist_mode=0, cutoff=1000, split=2.0, shift=0.5,
                 L=500, R=64, alpha=1.0):
        data = np.ascontiguousarray(data, dtype=np.float32)
        labels = np.ascontiguousarray(labels, dtype=np.float32)
        self.dim = data.shape[1]
        self.kind = kind
        cp = None if cache_path is None else cache_path.encode()
        self.h = lib().oracle_create(KINDS[kind], metric, dist_mode, data.shape[0], data.shape[1], data.ctypes.data,
                                     labels.ctypes.data, cutoff, split, shift, L, R, alpha, cp)
        if not self.h:
            raise RuntimeError("oracle_create: " + lib().oracle_last_error().decode())

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle_destroy(self.h)
            self.h = None

    def batch(self, method, queries, windows, k=10, beam=10, mult=1, max_beam=10000, ratio=None, pad_id=0,
              threads=None, stats=False):
        queries = np.ascontiguousarray(queries, dtype=np.float32)
        windows = np.ascontiguousarray(windows, dtype=np.float32)
        nq = len(windows)
        ids = np.empty((nq, k), dtype=np.uint32)
        dists = np.empty((nq, k), dtype=np.float32)
        st = np.zeros(3, dtype=np.uint64)
        threads = threads or os.cpu_count()
        rc = lib().oracle_batch(self.h, METHODS[method], queries.ctypes.data, windows.ctypes.data, nq, k, beam, mult,
                                max_beam, 0 if ratio is None else 1, 0.0 if ratio is None else ratio, pad_id, threads,
                                ids.ctypes.data, dists.ctypes.data, st.ctypes.data)
        if rc != 0:
            raise RuntimeError("oracle_batch: " + lib().oracle_last_error().decode())
        if stats:
            return ids, dists, dict(visited=int(st[0]), dist_cmps=int(st[1]), graph_searches=int(st[2]))
        return ids, dists

    def decompose(self, method, lo, hi, beam=10, mult=2, ratio=None, cap=4096):
        out = np.full((cap, 4), -2, dtype=np.int64)
        n = lib().oracle_decompose(self.h, METHODS[method], lo, hi, beam, mult, 0 if ratio is None else 1,
                                   0.0 if ratio is None else ratio, cap, out.ctypes.data)
        if n < 0:
            raise RuntimeError("oracle_decompose failed: " + lib().oracle_last_error().decode())
        return out[:n]
