"""Deterministic synthetic inputs of the shapes BASELINE.json names.

Nothing here is on the query hot path: it produces the vectors / labels / windows /
ground truth that both the reference (oracle/_ref) and this engine are fed.

Recipes mirror the reference's dataset scripts:
  * windows  — generate_datasets/filter_generation_utils.py:9-74
               (`generate_random_query_filter_ranges`, follow_data_distribution=True)
  * ground truth — generate_datasets/filter_generation_utils.py:142-168 (closed interval)
  * angular sets are L2-normalised — generate_datasets/generate_ann_benchmarks_datasets.py:42-44
"""
from __future__ import annotations

import numpy as np

FILTER_POWERS = list(range(-16, 1))  # filter_generation_utils.py:5
TOP_K = 10                           # filter_generation_utils.py:6


def make_vectors(n: int, d: int, rng: np.random.Generator, centers: np.ndarray,
                 normalize: bool = False) -> np.ndarray:
    """256-centre Gaussian mixture, sigma 0.5."""
    out = np.empty((n, d), dtype=np.float32)
    step = 1 << 18
    for s in range(0, n, step):
        e = min(n, s + step)
        c = rng.integers(0, centers.shape[0], size=e - s)
        out[s:e] = centers[c] + 0.5 * rng.standard_normal((e - s, d), dtype=np.float32)
    if normalize:
        out /= np.linalg.norm(out, axis=1, keepdims=True)
    return out


def make_dataset(n: int, d: int, nq: int, seed: int = 0, angular: bool = False,
                 label_kind: str = "unique"):
    """Returns (data[n,d] f32, queries[nq,d] f32, labels[n] f32).

    label_kind:
      "unique"    rng.permutation(n)/n as fp32 — unique for n <= 2^24 (avoids the
                  equal-label sort-order hazard)
      "uniform"   rng.uniform(0,1)
      "timestamp" RedCaps-like integer seconds (generate_redcaps_data.py:77-80)
    """
    rng = np.random.default_rng(seed)
    centers = rng.standard_normal((256, d)).astype(np.float32)
    data = make_vectors(n, d, rng, centers, normalize=angular)
    queries = make_vectors(nq, d, rng, centers, normalize=angular)
    if label_kind == "unique":
        labels = (rng.permutation(n).astype(np.float64) / n).astype(np.float32)
    elif label_kind == "uniform":
        labels = rng.uniform(0, 1, size=n).astype(np.float32)
    elif label_kind == "timestamp":
        labels = (1.2e9 + rng.integers(0, int(3.5e8), size=n)).astype(np.float32)
    elif label_kind == "timestamp_unique":
        # the same recipe (integer seconds cast to float32: multiples of 128 at this magnitude), ties nudged apart
        labels = make_labels_unique((1.2e9 + rng.integers(0, int(3.5e8), size=n)).astype(np.float32))
    else:
        raise ValueError(label_kind)
    return data, queries, labels


def make_labels_unique(labels: np.ndarray) -> np.ndarray:
    """Nudges duplicated labels to the next representable float32s (in label order), so that every label is
    distinct and the label order is total.  The reference sorts labels with an UNSTABLE parallel sort
    (tree_utils.h:70-72), so with ties the two implementations lay the tied points out differently and a graph file
    written for one layout is inconsistent with the other — a benchmark in which both arms load
    ONE set of graph files needs tie-free labels.  Moves a label by a few ulps at most where ties are sparse."""
    lab = labels.astype(np.float32).copy()
    order = np.argsort(lab, kind="stable")
    s = lab[order]
    for _ in range(64):
        dup = np.nonzero(s[1:] <= s[:-1])[0] + 1
        if len(dup) == 0:
            break
        s[dup] = np.nextafter(s[dup - 1], np.float32(np.inf))
    else:
        for i in range(1, len(s)):  # dense runs of ties: one sequential pass settles them
            if s[i] <= s[i - 1]:
                s[i] = np.nextafter(s[i - 1], np.float32(np.inf))
    lab[order] = s
    return lab


def make_rank_queries(d: int, nq: int, data_seed: int, rank: int, angular: bool = False) -> np.ndarray:
    """Queries of rank `rank` (> 0) of a query-sharded run: drawn from the same mixture as
    make_dataset(..., seed=data_seed) — the centres are the first draw of that stream — but from their
    own stream, so that data and labels (and with them the graph cache) do not depend on how many
    ranks there are.  Rank 0 keeps make_dataset's own queries."""
    centers = np.random.default_rng(data_seed).standard_normal((256, d)).astype(np.float32)
    rng = np.random.default_rng([int(data_seed), 7919, int(rank)])
    return make_vectors(nq, d, rng, centers, normalize=angular)


def make_windows(labels: np.ndarray, power: int, nq: int, seed: int) -> np.ndarray:
    """Window recipe of filter_generation_utils.py:9-52 for fraction 2**power.

    Returned as float32 [nq,2]: the boundary casts windows to pair<float,float>
    (prefiltering.h:126) so both implementations see exactly these values.
    """
    rng = np.random.default_rng(seed)
    s = np.sort(labels.astype(np.float64))
    n = len(s)
    frac = 2.0 ** power
    if frac == 1:
        lo = s[0] - rng.integers(1, 100)
        hi = s[-1] + rng.integers(1, 100)
        return np.tile(np.array([[lo, hi]], dtype=np.float32), (nq, 1))
    m = int(n * frac)
    start = rng.integers(0, n - m, size=nq)
    end = start + m
    lo_v, hi_v = s[start], s[end]
    gap_l = np.where(start > 0, lo_v - s[np.maximum(start - 1, 0)], 1.0)
    gap_r = np.where(end < n - 1, s[np.minimum(end + 1, n - 1)] - hi_v, 1.0)
    lo = lo_v - rng.uniform(size=nq) * gap_l
    hi = hi_v + rng.uniform(size=nq) * gap_r
    return np.stack([lo, hi], axis=1).astype(np.float32)


def ground_truth(data: np.ndarray, queries: np.ndarray, labels: np.ndarray,
                 windows: np.ndarray, k: int = TOP_K, angular: bool = False) -> np.ndarray:
    """Brute-force closed-interval top-k (filter_generation_utils.py:142-168), numpy.

    Slices the label-sorted copy so cost is O(sum of window sizes). Pads with -1 when
    a window holds fewer than k points.
    """
    order = np.argsort(labels, kind="stable")
    sl = labels[order]
    sd = data[order]
    gt = np.full((len(queries), k), -1, dtype=np.int64)
    for i, q in enumerate(queries):
        a = np.searchsorted(sl, windows[i, 0], side="left")
        b = np.searchsorted(sl, windows[i, 1], side="right")
        if b <= a:
            continue
        x = sd[a:b]
        if angular:
            dist = -(x @ q)
        else:
            diff = x - q
            dist = np.einsum("ij,ij->i", diff, diff)
        kk = min(k, b - a)
        idx = np.argpartition(dist, kk - 1)[:kk] if kk < b - a else np.arange(b - a)
        idx = idx[np.argsort(dist[idx], kind="stable")]
        gt[i, :kk] = order[a + idx]
    return gt


def recall(results: np.ndarray, gt: np.ndarray, k: int = TOP_K) -> float:
    """compute_recall as the driver CALLS it (run_our_method.py:174-180,256: arguments
    swapped, so the denominator is the number of distinct returned ids)."""
    tot = 0.0
    for i in range(len(results)):
        res = set(int(x) for x in results[i][:k])
        g = set(int(x) for x in gt[i][:k])
        tot += len(res & g) / len(res)
    return tot / len(results)


def recall_std(results: np.ndarray, gt: np.ndarray, k: int = TOP_K) -> float:
    """|top-k ∩ GT| / |GT| averaged (the textbook definition; GT pads (-1) ignored)."""
    tot = 0.0
    for i in range(len(results)):
        g = set(int(x) for x in gt[i][:k] if x >= 0)
        if not g:
            tot += 1.0
            continue
        res = set(int(x) for x in results[i][:k])
        tot += len(res & g) / len(g)
    return tot / len(results)


def make_adversarial(n: int, d: int, seed: int = 0, clusters: int = 100, intra_var: float = 0.01,
                     inter_var: float = 1.0):
    """The reference's adversarial set (generate_datasets/generate_advserial_dataset.py:8-60): `clusters` tight
    Gaussian clusters, the label of a point is its cluster index - 0.5 + U(0,1) — so label ranges and clusters
    coincide — and every query is drawn from cluster a while its window is cluster b's label range (a != b): the
    graph neighbourhood of the query holds no in-window point, which is the worst case for postfiltering.
    Rows are L2-normalised as there.  Returns (data, queries[clusters*(clusters-1)], labels, windows)."""
    rng = np.random.default_rng(seed)
    per = n // clusters
    means = (np.sqrt(inter_var) * rng.standard_normal((clusters, d))).astype(np.float32)
    data = np.empty((per * clusters, d), np.float32)
    labels = np.empty(per * clusters, np.float32)
    for c in range(clusters):
        data[c * per:(c + 1) * per] = means[c] + np.sqrt(intra_var) * rng.standard_normal((per, d), dtype=np.float32)
        labels[c * per:(c + 1) * per] = (c - 0.5 + rng.uniform(size=per)).astype(np.float32)
    qc, gc = np.nonzero(~np.eye(clusters, dtype=bool))
    queries = means[qc] + np.sqrt(intra_var) * rng.standard_normal((len(qc), d), dtype=np.float32)
    windows = np.stack([gc - 0.5, gc + 0.5], axis=1).astype(np.float32)
    data /= np.linalg.norm(data, axis=1, keepdims=True)
    queries /= np.linalg.norm(queries, axis=1, keepdims=True)
    return data, queries.astype(np.float32), make_labels_unique(labels), windows


def make_blowup_windows(labels: np.ndarray, power: int, nq: int, seed: int) -> np.ndarray:
    """Large-blow-up windows (SURVEY.md §8d, C5-adversarial ii): windows of fraction 2**power that straddle the
    median label rank, so the smallest B-WST bucket containing one is the ROOT (blow-up = 2**-power): optimized
    postfiltering must search the whole data set with selectivity 2**power."""
    rng = np.random.default_rng(seed)
    s = np.sort(labels.astype(np.float64))
    n = len(s)
    m = max(2, int(n * 2.0 ** power))
    start = n // 2 - rng.integers(1, m, size=nq)
    start = np.clip(start, 0, n - m - 1)
    return np.stack([s[start], s[start + m]], axis=1).astype(np.float32)
