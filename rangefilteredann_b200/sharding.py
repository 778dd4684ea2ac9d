"""Host-side arithmetic of the multi-GPU paths: queries are independent units, so a batch is
sharded over ranks / devices with the index replicated and no data-path collective; label shards are contiguous
ranges of the label-sorted points.  Pure index arithmetic — the rank plumbing (barrier, timing reduction) belongs
to the launcher (bench.py uses torch.distributed for it)."""
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
