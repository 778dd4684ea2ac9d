"""Host-side plumbing of the multi-GPU path: queries are independent units,
so the batch is sharded over ranks with the index replicated and no data-path collective.
torch.distributed is used only for the barrier and the max-over-ranks timing reduction
(NCCL on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_bounds(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of `n_items` for `rank` (sizes differ by <= 1)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_queries(queries: np.ndarray, windows: np.ndarray, rank: int, world: int):
    """Strong-scaling split of one batch: rank's rows of (queries, windows)."""
    lo, hi = shard_bounds(len(queries), rank, world)
    return np.ascontiguousarray(queries[lo:hi]), np.ascontiguousarray(windows[lo:hi])


def weak_batch(queries_all: np.ndarray, nq: int, rank: int) -> np.ndarray:
    """Weak-scaling batch of bench.py: rank r answers its own block of nq queries."""
    return np.ascontiguousarray(queries_all[rank * nq:(rank + 1) * nq])


def reduce_max(values, device=None) -> list[float]:
    """max over ranks of a small vector of timings (all ranks get the result)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.cpu()]


def gather_rows(local: np.ndarray, world: int) -> np.ndarray | None:
    """Concatenate per-rank result rows on rank 0 (host gather of nq x k ids; the data path
    itself needs no collective)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or world == 1:
        return local
    objs = [None] * world if dist.get_rank() == 0 else None
    dist.gather_object(local, objs, dst=0)
    return np.concatenate(objs, axis=0) if dist.get_rank() == 0 else None
