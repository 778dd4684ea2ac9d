"""Label-range sharded window search (BASELINE.json config 5 shape).

For datasets that do not fit one GPU, rank r owns the contiguous slice of the label-sorted
points [r*N/W, (r+1)*N/W) with its own B-WST over that slice.  Every rank answers the whole
(small) query batch on its shard — windows that miss the shard's label range come back as
pads — then the per-rank [nq][k] rows are all-gathered (NCCL over NVLink/NVSwitch on GPUs;
gloo in the CPU tests) and merged per query on every rank (`ws_merge_partial_topk`).

This equals the reference's fenwick decomposition cut at the shard boundaries: a window
spanning several shards is answered by <= W smaller sub-trees instead of one large node, so
recall can only rise (range_filter_tree.h:297-401).  torch.distributed is plumbing only.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

FLT_MAX = np.float32(3.4028235e38)


def shard_of_sorted_labels(labels: np.ndarray, rank: int, world: int):
    """Original ids of the points rank `rank` owns (contiguous in label order; ties by id)."""
    from .sharding import shard_bounds
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
    """One rank's shard + the collective batch_search.  Construct on every rank."""

    def __init__(self, data: np.ndarray, labels: np.ndarray, rank: int, world: int, cache_root: str,
                 cutoff: int = 1000, split_factor: int = 2, metric: str = "Euclidian",
                 max_degree: int = 64, limit: int = 500, alpha: float = 1.0):
        from . import load_engine
        self.rank, self.world = rank, world
        self.eng = load_engine()
        self.owned = shard_of_sorted_labels(labels, rank, world).astype(np.uint32)  # local id -> global id
        sfx = "FloatMips" if metric == "mips" else "FloatEuclidian"
        cache = os.path.join(cache_root, f"shard{rank}of{world}") + "/"
        os.makedirs(cache, exist_ok=True)
        self.tree = getattr(self.eng, "VamanaRangeFilterTreeIndex" + sfx)(
            np.ascontiguousarray(data[self.owned]), np.ascontiguousarray(labels[self.owned]), cutoff, split_factor,
            self.eng.BuildParams(max_degree, limit, alpha, cache))

    def local_search(self, queries, windows, method, qp):
        """This shard's partial rows with GLOBAL ids (pads keep dist FLT_MAX)."""
        ids, dists = self.tree.batch_search(queries, windows, len(windows), method, qp)
        gids = self.owned[np.minimum(ids, len(self.owned) - 1)]
        gids[dists == FLT_MAX] = 0
        return gids.astype(np.uint32), dists

    def batch_search(self, queries, windows, method, qp, k: int):
        """Collective: local search, all-gather of [nq][k] rows, per-query merge on device."""
        gids, dists = self.local_search(queries, windows, method, qp)
        from . import capi
        return allgather_merge(gids, dists, k, self.world, lambda: capi.Handle.borrow(self.tree))


def allgather_merge(gids: np.ndarray, dists: np.ndarray, k: int, world: int, handle_factory=None):
    """All-gather every rank's [nq][k] (global id, dist) rows and merge them per query.
    NCCL backend: tensors stay on the GPU and `ws_merge_partial_topk` merges; gloo backend
    (CPU tests): numpy merge."""
    import torch
    import torch.distributed as dist
    nq = len(gids)
    if world == 1 or not dist.is_initialized():
        return gids, dists
    on_gpu = dist.get_backend() == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
    t_ids = torch.from_numpy(np.ascontiguousarray(gids).view(np.int32)).to(dev)  # NCCL has no uint32; bits preserved
    t_d = torch.from_numpy(np.ascontiguousarray(dists)).to(dev)
    all_ids = torch.empty((world * nq, k), dtype=torch.int32, device=dev)  # rank-major concatenation
    all_d = torch.empty((world * nq, k), dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(all_ids, t_ids)
    dist.all_gather_into_tensor(all_d, t_d)
    all_ids, all_d = all_ids.view(world, nq, k), all_d.view(world, nq, k)
    if not on_gpu:
        return merge_partial_topk_numpy(all_ids.numpy().view(np.uint32), all_d.numpy(), k)
    h = handle_factory()
    out_ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
    out_d = torch.empty((nq, k), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()  # the gathered tensors were produced on torch's stream
    h.merge_partial_topk(C.c_void_p(all_ids.data_ptr()), C.c_void_p(all_d.data_ptr()), world, nq, k, 0,
                         C.c_void_p(out_ids.data_ptr()), C.c_void_p(out_d.data_ptr()))
    h.sync()
    return out_ids.cpu().numpy().view(np.uint32), out_d.cpu().numpy()
