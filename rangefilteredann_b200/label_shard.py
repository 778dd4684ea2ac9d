"""Label-range sharded window search, one process per GPU (BASELINE.json config 5 shape).

For data sets that do not fit one GPU, rank r owns the contiguous slice [r*N/W, (r+1)*N/W) of the label-sorted
points with its own B-WST over that slice.  Every rank answers the whole (small) query batch on its shard —
windows that miss the shard's label range come back as pads — then the per-rank [nq][k] rows are all-gathered
with NCCL over NVLink/NVSwitch and merged per query on every rank.  Both steps run inside libwsann_cuda.so
(`ws_allgather_merge`: ncclAllGather on the index stream + the K4b merge kernel, device memory only); the host
program only has to hand rank 0's 128-byte NCCL id to the other ranks (any broadcast it already has).

This equals the reference's fenwick decomposition cut at the shard boundaries: a window spanning several shards
is answered by <= W smaller sub-trees instead of one large node, so recall can only rise
(range_filter_tree.h:297-401).  Prefilter rows are identical to the single-GPU rows.

The single-process form of the same thing (all GPUs of the box behind one `batch_search` call, partial rows
gathered by peer loads) is selected with WSANN_DEVICES=… WSANN_SHARD_MODE=label (csrc/host/window_index.hpp).
"""
from __future__ import annotations

import os

import numpy as np

from . import capi
from .sharding import shard_bounds

FLT_MAX = np.float32(3.4028235e38)


def shard_of_sorted_labels(labels: np.ndarray, rank: int, world: int):
    """Original ids of the points rank `rank` owns (contiguous in label order; ties by id)."""
    order = np.argsort(labels, kind="stable")
    lo, hi = shard_bounds(len(labels), rank, world)
    return order[lo:hi]


def merge_partial_topk_numpy(ids: np.ndarray, dists: np.ndarray, k: int, pad_id: int = 0):
    """Reference implementation of the merge (CPU tests): ids/dists are [parts][nq][k]."""
    parts, nq, _ = ids.shape
    flat_i = np.transpose(ids, (1, 0, 2)).reshape(nq, parts * k)
    flat_d = np.transpose(dists, (1, 0, 2)).reshape(nq, parts * k)
    out_i = np.full((nq, k), pad_id, np.uint32)
    out_d = np.full((nq, k), FLT_MAX, np.float32)
    for q in range(nq):
        valid = flat_d[q] != FLT_MAX
        order = np.lexsort((flat_i[q][valid], flat_d[q][valid]))[:k]
        out_i[q, :len(order)] = flat_i[q][valid][order]
        out_d[q, :len(order)] = flat_d[q][valid][order]
    return out_i, out_d


class LabelShardedTree:
    """One rank's shard + the collective batch_search.  Construct on every rank; `unique_id` is
    `capi.nccl_unique_id()` of rank 0, handed to every rank by the host program (None: no communicator, only
    `local_search` works — single-process tests)."""

    def __init__(self, data: np.ndarray, labels: np.ndarray, rank: int, world: int, cache_root: str,
                 cutoff: int = 1000, split_factor: int = 2, metric: str = "Euclidian",
                 max_degree: int = 64, limit: int = 500, alpha: float = 1.0, unique_id: bytes | None = None):
        from . import load_engine
        self.rank, self.world = rank, world
        self.eng = load_engine()
        self.owned = shard_of_sorted_labels(labels, rank, world).astype(np.uint32)  # shard-local id -> data-set id
        sfx = "FloatMips" if metric == "mips" else "FloatEuclidian"
        cache = os.path.join(cache_root, f"shard{rank}of{world}") + "/"
        os.makedirs(cache, exist_ok=True)
        self.tree = getattr(self.eng, "VamanaRangeFilterTreeIndex" + sfx)(
            np.ascontiguousarray(data[self.owned]), np.ascontiguousarray(labels[self.owned]), cutoff, split_factor,
            self.eng.BuildParams(max_degree, limit, alpha, cache))
        self.h = capi.Handle.borrow(self.tree)
        # the shard's rows are already label-sorted (ties by id), so arena rank == shard-local id: result rows carry
        # data-set ids straight from the kernels
        self.h.set_decode(self.owned)
        if rank + 1 < world:  # prefiltering.h:159-184's r = n-1 rule excludes the DATA SET's last point only
            self.h.set_option("prefilter_open_tail", 1)
        self.dim = data.shape[1]
        self._cap = 0
        if unique_id is not None and world > 1:
            self.h.comm_init(world, rank, unique_id)
        self._has_comm = unique_id is not None and world > 1

    # ---- device-resident plumbing
    def _ensure(self, nq: int, k: int):
        if nq * k <= self._cap:
            return
        h = self.h
        self.dq = h.dalloc(nq * self.dim * 4)
        self.dw = h.dalloc(nq * 8)
        self.part_ids, self.part_d = h.dalloc(nq * k * 4), h.dalloc(nq * k * 4)
        self.out_ids, self.out_d = h.dalloc(nq * k * 4), h.dalloc(nq * k * 4)
        self._cap = nq * k

    def upload(self, queries: np.ndarray, windows: np.ndarray, k: int):
        nq = len(windows)
        self._ensure(nq, k)
        self.h.h2d(self.dq, np.ascontiguousarray(queries[:nq], dtype=np.float32))
        self.h.h2d(self.dw, np.ascontiguousarray(windows, dtype=np.float32))
        return nq

    def search_device(self, nq: int, method: str, qp_c: capi.QueryParamsC, k: int):
        """Enqueue: local search -> ncclAllGather -> merge, all on the index stream (no host sync)."""
        if method == "prefilter":
            self.h.prefilter_batch(self.dq, self.dw, nq, k, self.part_ids, self.part_d, device_ptrs=True)
        else:
            self.h.tree_batch(method, self.dq, self.dw, nq, qp_c, self.part_ids, self.part_d, device_ptrs=True)
        if self._has_comm:
            self.h.allgather_merge(self.part_ids, self.part_d, nq, k, 0xFFFFFFFF if method == "prefilter" else 0, self.out_ids, self.out_d)

    def fetch(self, nq: int, k: int):
        ids, d = np.empty((nq, k), np.uint32), np.empty((nq, k), np.float32)
        src = (self.out_ids, self.out_d) if self._has_comm else (self.part_ids, self.part_d)
        self.h.d2h(ids, src[0])
        self.h.d2h(d, src[1])
        return ids, d

    # ---- host-buffer calls
    def local_search(self, queries, windows, method, qp):
        """This shard's partial rows with data-set ids (pads keep dist FLT_MAX)."""
        if method == "prefilter":
            raise ValueError("local_search answers tree methods")
        ids, dists = self.tree.batch_search(queries, windows, len(windows), method, qp)
        ids = ids.copy()
        ids[dists == FLT_MAX] = 0
        return ids, dists

    def batch_search(self, queries, windows, method, qp_c: capi.QueryParamsC, k: int):
        """Collective: every rank calls it with the same batch and gets the same merged rows."""
        nq = self.upload(queries, windows, k)
        self.search_device(nq, method, qp_c, k)
        return self.fetch(nq, k)
