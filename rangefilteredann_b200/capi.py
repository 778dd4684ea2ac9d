"""ctypes view of include/wsann.h — plumbing for tests and bench.py (device-resident
batches, CUDA-event timers, counters).  The query path itself lives in libwsann_cuda.so;
nothing here computes."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwsann_cuda.so")

WS_OK = 0
WS_FLAG_DEVICE_PTRS = 1
METHODS = {"fenwick": 0, "optimized_postfilter": 1, "three_split": 2, "super": 3}
MODE_PREFILTER = 10


class QueryParamsC(C.Structure):
    _fields_ = [("k", C.c_int64), ("beam_size", C.c_int64), ("cut", C.c_double), ("limit", C.c_int64),
                ("degree_limit", C.c_int64), ("final_beam_multiply", C.c_int64),
                ("postfiltering_max_beam", C.c_int64), ("min_query_to_bucket_ratio", C.c_float),
                ("has_min_query_to_bucket_ratio", C.c_int32), ("verbose", C.c_int32)]


class StatsC(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("graph_searches", "visited", "dist_cmps", "scan_points", "graph_tasks",
                                           "scan_tasks", "escalated_tasks", "beam_sum")]


def query_params(k=10, beam=10, final_multiply=1, max_beam=10000, ratio=None, cut=1.35, limit=10_000_000,
                 degree_limit=10_000) -> QueryParamsC:
    return QueryParamsC(k, beam, cut, limit, degree_limit, final_multiply, max_beam,
                        0.0 if ratio is None else float(ratio), 0 if ratio is None else 1, 0)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} missing: run `python -m rangefilteredann_b200.build` (no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.ws_last_error.restype = C.c_char_p
        vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int32
        L.ws_index_create.argtypes = [C.c_int, C.c_int, u64, u32, vp, vp, vp, C.c_int, C.POINTER(vp)]
        L.ws_index_destroy.argtypes = [vp]
        L.ws_index_destroy.restype = None
        L.ws_index_add_graph.argtypes = [vp, u64, u64, u32, vp, vp, C.POINTER(i32)]
        L.ws_index_set_wst.argtypes = [vp, u32, u32, i32, vp, vp, vp]
        L.ws_index_set_super.argtypes = [vp, u32, i32, vp, vp, vp, vp]
        L.ws_index_finalize.argtypes = [vp]
        L.ws_index_set_decode.argtypes = [vp, vp]
        L.ws_index_save.argtypes = [vp, C.c_char_p]
        L.ws_index_load.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp)]
        L.ws_index_shape.argtypes = [vp, C.POINTER(u64), C.POINTER(u32), C.POINTER(C.c_int), C.POINTER(u64)]
        L.ws_prefilter_batch.argtypes = [vp, vp, vp, u64, u32, vp, vp, u32]
        L.ws_postfilter_batch.argtypes = [vp, i32, vp, vp, u64, C.POINTER(QueryParamsC), C.c_int, vp, vp, u32]
        L.ws_tree_batch.argtypes = [vp, C.c_int, vp, vp, u64, C.POINTER(QueryParamsC), vp, vp, u32]
        L.ws_merge_partial_topk.argtypes = [vp, vp, vp, u32, u64, u32, u32, vp, vp]
        L.ws_index_sync.argtypes = [vp]
        L.ws_device_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
        L.ws_device_free.argtypes = [vp, vp]
        L.ws_host_alloc_pinned.argtypes = [C.c_size_t, C.POINTER(vp)]
        L.ws_host_free_pinned.argtypes = [vp]
        L.ws_copy_h2d.argtypes = [vp, vp, vp, C.c_size_t]
        L.ws_copy_d2h.argtypes = [vp, vp, vp, C.c_size_t]
        L.ws_timer_start.argtypes = [vp]
        L.ws_timer_stop.argtypes = [vp, C.POINTER(C.c_float)]
        L.ws_flush_l2.argtypes = [vp]
        L.ws_index_get_stats.argtypes = [vp, C.POINTER(StatsC)]
        L.ws_index_reset_stats.argtypes = [vp]
        L.ws_index_launch_count.argtypes = [vp, C.POINTER(u64)]
        L.ws_index_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
        L.ws_index_kernel_times.argtypes = [vp, vp, vp, C.c_int]
        L.ws_index_hbm_bytes.argtypes = [vp, C.POINTER(u64)]
        L.ws_index_task_capacity.argtypes = [vp, C.c_int, C.POINTER(u32)]
        L.ws_debug_decompose_host.argtypes = [vp, C.c_int, vp, u64, C.POINTER(QueryParamsC), u32, vp, vp]
        L.ws_device_count.argtypes = [C.POINTER(C.c_int)]
        L.ws_index_device.argtypes = [vp, C.POINTER(C.c_int)]
        # multi-GPU (ws_group / NCCL inside the library)
        L.ws_index_replicate.argtypes = [vp, C.c_int, C.POINTER(vp)]
        L.ws_group_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.POINTER(vp)]
        L.ws_group_destroy.argtypes = [vp]
        L.ws_group_destroy.restype = None
        L.ws_group_size.argtypes = [vp, C.POINTER(C.c_int)]
        L.ws_group_member.argtypes = [vp, C.c_int, C.POINTER(vp)]
        L.ws_group_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
        L.ws_group_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.ws_group_prefilter_batch.argtypes = [vp, vp, vp, u64, u32, vp, vp]
        L.ws_group_postfilter_batch.argtypes = [vp, i32, vp, vp, u64, C.POINTER(QueryParamsC), C.c_int, vp, vp]
        L.ws_group_tree_batch.argtypes = [vp, C.c_int, vp, vp, u64, C.POINTER(QueryParamsC), vp, vp]
        L.ws_nccl_unique_id.argtypes = [vp]
        L.ws_nccl_version.argtypes = [C.POINTER(C.c_int)]
        L.ws_index_comm_init.argtypes = [vp, C.c_int, C.c_int, vp]
        L.ws_index_comm_destroy.argtypes = [vp]
        L.ws_allgather_merge.argtypes = [vp, vp, vp, u64, u32, u32, vp, vp]
        _lib = L
    return _lib


class WsError(RuntimeError):
    pass


def check(status: int, what: str = "wsann") -> None:
    if status != WS_OK:
        raise WsError(f"{what}: {lib().ws_last_error().decode()} (status {status})")


def ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


class Handle:
    """Thin owner/borrower of a ws_index*; `borrow` wraps the arena of a pybind index object."""

    def __init__(self, raw: int, owned: bool):
        self.raw = C.c_void_p(raw)
        self.owned = owned

    @classmethod
    def borrow(cls, index_obj) -> "Handle":
        return cls(index_obj._arena_handle(), False)

    def __del__(self):
        if self.owned and self.raw:
            lib().ws_index_destroy(self.raw)
            self.raw = None

    # ---- plumbing
    def sync(self):
        check(lib().ws_index_sync(self.raw), "ws_index_sync")

    def dalloc(self, nbytes: int) -> C.c_void_p:
        p = C.c_void_p()
        check(lib().ws_device_alloc(self.raw, nbytes, C.byref(p)), "ws_device_alloc")
        return p

    def dfree(self, p):
        check(lib().ws_device_free(self.raw, p), "ws_device_free")

    def h2d(self, dptr, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        check(lib().ws_copy_h2d(self.raw, dptr, ptr(arr), arr.nbytes), "ws_copy_h2d")

    def d2h(self, arr: np.ndarray, dptr):
        check(lib().ws_copy_d2h(self.raw, ptr(arr), dptr, arr.nbytes), "ws_copy_d2h")

    def timer_start(self):
        check(lib().ws_timer_start(self.raw), "ws_timer_start")

    def timer_stop(self) -> float:
        ms = C.c_float()
        check(lib().ws_timer_stop(self.raw, C.byref(ms)), "ws_timer_stop")
        return ms.value

    def flush_l2(self):
        check(lib().ws_flush_l2(self.raw), "ws_flush_l2")

    def stats(self) -> dict:
        s = StatsC()
        check(lib().ws_index_get_stats(self.raw, C.byref(s)), "ws_index_get_stats")
        return {n: getattr(s, n) for n, _ in StatsC._fields_}

    def reset_stats(self):
        check(lib().ws_index_reset_stats(self.raw), "ws_index_reset_stats")

    def launches(self) -> int:
        v = C.c_uint64()
        check(lib().ws_index_launch_count(self.raw, C.byref(v)), "ws_index_launch_count")
        return v.value

    def set_option(self, name: str, value: int):
        check(lib().ws_index_set_option(self.raw, name.encode(), int(value)), "ws_index_set_option")

    KERNEL_KINDS = ("decompose", "beam_warp64", "beam_warp128", "beam_warp256", "beam_warp512", "beam_warp1024", "beam_large",
                    "scan", "merge", "gemm_sweep", "gemm_plan_pack", "gemm_rerank", "gemm_seed")

    def kernel_times(self, reset=True) -> dict:
        ms = np.zeros(len(self.KERNEL_KINDS), np.float64)
        n = np.zeros(len(self.KERNEL_KINDS), np.uint64)
        check(lib().ws_index_kernel_times(self.raw, ptr(ms), ptr(n), 1 if reset else 0), "ws_index_kernel_times")
        return {k: dict(ms=float(ms[i]), launches=int(n[i])) for i, k in enumerate(self.KERNEL_KINDS) if n[i]}

    def hbm_bytes(self) -> int:
        v = C.c_uint64()
        check(lib().ws_index_hbm_bytes(self.raw, C.byref(v)), "ws_index_hbm_bytes")
        return v.value

    # ---- batches (host or device pointers)
    def tree_batch(self, method: str, queries, windows, nq: int, qp: QueryParamsC, ids, dists, device_ptrs=False):
        f = WS_FLAG_DEVICE_PTRS if device_ptrs else 0
        a = [x if isinstance(x, C.c_void_p) else ptr(x) for x in (queries, windows, ids, dists)]
        check(lib().ws_tree_batch(self.raw, METHODS[method], a[0], a[1], nq, C.byref(qp), a[2], a[3], f), "ws_tree_batch")

    def prefilter_batch(self, queries, windows, nq: int, k: int, ids, dists, device_ptrs=False):
        f = WS_FLAG_DEVICE_PTRS if device_ptrs else 0
        a = [x if isinstance(x, C.c_void_p) else ptr(x) for x in (queries, windows, ids, dists)]
        check(lib().ws_prefilter_batch(self.raw, a[0], a[1], nq, k, a[2], a[3], f), "ws_prefilter_batch")

    def merge_partial_topk(self, ids_dev, dists_dev, parts: int, nq: int, k: int, pad_id: int, out_ids_dev, out_dists_dev):
        check(lib().ws_merge_partial_topk(self.raw, ids_dev, dists_dev, parts, nq, k, pad_id, out_ids_dev, out_dists_dev),
              "ws_merge_partial_topk")

    # ---- one process per GPU: NCCL inside the library
    def comm_init(self, nranks: int, rank: int, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, NCCL_ID_BYTES)
        check(lib().ws_index_comm_init(self.raw, nranks, rank, buf), "ws_index_comm_init")

    def comm_destroy(self):
        check(lib().ws_index_comm_destroy(self.raw), "ws_index_comm_destroy")

    def allgather_merge(self, ids_dev, dists_dev, nq: int, k: int, pad_id: int, out_ids_dev, out_dists_dev):
        check(lib().ws_allgather_merge(self.raw, ids_dev, dists_dev, nq, k, pad_id, out_ids_dev, out_dists_dev),
              "ws_allgather_merge")

    def set_decode(self, decode: np.ndarray):
        decode = np.ascontiguousarray(decode, dtype=np.uint32)
        check(lib().ws_index_set_decode(self.raw, ptr(decode)), "ws_index_set_decode")

    def replicate(self, device: int) -> "Handle":
        out = C.c_void_p()
        check(lib().ws_index_replicate(self.raw, device, C.byref(out)), "ws_index_replicate")
        return Handle(out.value, True)

    def postfilter_batch(self, node: int, queries, windows, nq: int, qp: QueryParamsC, pad: int, ids, dists,
                         device_ptrs=False):
        f = WS_FLAG_DEVICE_PTRS if device_ptrs else 0
        a = [x if isinstance(x, C.c_void_p) else ptr(x) for x in (queries, windows, ids, dists)]
        check(lib().ws_postfilter_batch(self.raw, node, a[0], a[1], nq, C.byref(qp), pad, a[2], a[3], f),
              "ws_postfilter_batch")


NCCL_ID_BYTES = 128
GROUP_REPLICATED, GROUP_LABEL_SHARDED = 0, 1


def nccl_unique_id() -> bytes:
    """ncclGetUniqueId through the library (rank 0 calls it and hands the bytes to the other ranks)."""
    buf = C.create_string_buffer(NCCL_ID_BYTES)
    check(lib().ws_nccl_unique_id(buf), "ws_nccl_unique_id")
    return buf.raw


def nccl_version() -> int:
    v = C.c_int()
    check(lib().ws_nccl_version(C.byref(v)), "ws_nccl_version")
    return v.value


class Group:
    """ws_group*: several arenas, one per device, behind one batch call (include/wsann.h)."""

    def __init__(self, raw: int, owned: bool, keep=None):
        self.raw = C.c_void_p(raw)
        self.owned = owned
        self._keep = keep  # member handles that must outlive the group

    @classmethod
    def create(cls, members: list, mode: int) -> "Group":
        arr = (C.c_void_p * len(members))(*[m.raw for m in members])
        out = C.c_void_p()
        check(lib().ws_group_create(arr, len(members), mode, C.byref(out)), "ws_group_create")
        return cls(out.value, True, list(members))

    @classmethod
    def borrow(cls, index_obj) -> "Group | None":
        raw = index_obj._group_handle()
        return cls(raw, False) if raw else None

    def __del__(self):
        if self.owned and self.raw:
            lib().ws_group_destroy(self.raw)
            self.raw = None

    def size(self) -> int:
        n = C.c_int()
        check(lib().ws_group_size(self.raw, C.byref(n)), "ws_group_size")
        return n.value

    def member(self, i: int) -> Handle:
        out = C.c_void_p()
        check(lib().ws_group_member(self.raw, i, C.byref(out)), "ws_group_member")
        return Handle(out.value, False)

    def set_option(self, name: str, value: int):
        check(lib().ws_group_set_option(self.raw, name.encode(), int(value)), "ws_group_set_option")

    def info(self) -> dict:
        peer, exch = C.c_int(), C.c_int()
        ms = (C.c_double * 3)()
        check(lib().ws_group_info(self.raw, C.byref(peer), C.byref(exch), ms), "ws_group_info")
        return {"peer_access": bool(peer.value), "exchange": "nccl_allgather" if exch.value else "peer_loads",
                "search_ms": ms[0], "exchange_merge_ms": ms[1], "total_ms": ms[2]}

    def prefilter_batch(self, queries, windows, nq: int, k: int, ids, dists):
        check(lib().ws_group_prefilter_batch(self.raw, ptr(queries), ptr(windows), nq, k, ptr(ids), ptr(dists)),
              "ws_group_prefilter_batch")

    def tree_batch(self, method: str, queries, windows, nq: int, qp: QueryParamsC, ids, dists):
        check(lib().ws_group_tree_batch(self.raw, METHODS[method], ptr(queries), ptr(windows), nq, C.byref(qp), ptr(ids), ptr(dists)),
              "ws_group_tree_batch")


def pinned_array(shape, dtype) -> np.ndarray:
    """numpy array backed by cudaMallocHost memory (kept alive for the process lifetime)."""
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    check(lib().ws_host_alloc_pinned(max(nbytes, 1), C.byref(p)), "ws_host_alloc_pinned")
    buf = (C.c_char * nbytes).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)
