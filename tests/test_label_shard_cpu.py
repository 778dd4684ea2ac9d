"""CPU: host-side logic of the label-sharded mode and of the bench's synthetic inputs.

  * prefilter bounds on label shards: the reference's `r = n-1` rule (prefiltering.h:159-184: the data set's last
    point is never inside a window) must hold for the DATA SET's last point only.  Host-only arenas (device -1) over
    the shards' label slices, evaluated through ws_debug_decompose_host with `prefilter_open_tail` on every shard but
    the last: the union of the shards' slices must be exactly the single arena's slice.
  * synth.make_labels_unique: total order, order-preserving, minimal movement.
  * bench helpers: one workload string for both arms, routing thresholds, row comparison tolerances."""
import ctypes as C

import numpy as np
import pytest

import bench
from rangefilteredann_b200 import capi, label_shard, sharding, synth
from test_decompose_cpu import host_decompose


def host_prefilter_index(labels_sorted, open_tail=False):
    L = capi.lib()
    h = C.c_void_p()
    capi.check(L.ws_index_create(-1, 0, len(labels_sorted), 8, None, capi.ptr(np.ascontiguousarray(labels_sorted)), None, 1, C.byref(h)))
    idx = capi.Handle(h.value, True)
    capi.check(L.ws_index_finalize(idx.raw))
    if open_tail:
        idx.set_option("prefilter_open_tail", 1)
    return idx


def slices_of(idx, windows):
    """[a, b) of every window (scan tasks of one query joined; empty windows -> (0, 0))"""
    out = []
    for t in host_decompose(idx, "prefilter", windows):
        rows = [(int(r[1]), int(r[2])) for r in t if r[2] > r[1]]
        if not rows:
            out.append((0, 0))
            continue
        assert all(rows[i][1] == rows[i + 1][0] for i in range(len(rows) - 1))
        out.append((rows[0][0], rows[-1][1]))
    return out


@pytest.mark.parametrize("n,world", [(1000, 2), (1003, 4), (4096, 8)])
def test_shard_prefilter_bounds_union_equals_single_arena(n, world):
    rng = np.random.default_rng(n + world)
    labels = np.sort(rng.permutation(n).astype(np.float32) / n)
    single = host_prefilter_index(labels)
    shards, offs = [], []
    for r in range(world):
        lo, hi = sharding.shard_bounds(n, r, world)
        shards.append(host_prefilter_index(labels[lo:hi], open_tail=r + 1 < world))
        offs.append(lo)
    lo_v = rng.uniform(-0.1, 1.0, size=300).astype(np.float32)
    w = np.stack([lo_v, lo_v + rng.uniform(0, 0.6, size=300).astype(np.float32)], axis=1)
    w[0] = (-1.0, 2.0)                       # everything: the data set's last point stays excluded, no shard's does
    w[1] = (labels[n // world - 1], labels[n // world])      # one point on each side of the first boundary
    w[2] = (labels[-3], labels[-1] + 1)      # through the data set's last point
    w[3] = (0.5, 0.5)                        # empty
    ref = slices_of(single, w)
    got = [slices_of(s, w) for s in shards]
    for q in range(len(w)):
        parts = [(a + offs[r], b + offs[r]) for r, (a, b) in ((r, got[r][q]) for r in range(world)) if b > a]
        if ref[q][1] <= ref[q][0]:
            assert not parts, (q, parts)
            continue
        assert parts[0][0] == ref[q][0] and parts[-1][1] == ref[q][1], (q, ref[q], parts)
        assert all(parts[i][1] == parts[i + 1][0] for i in range(len(parts) - 1)), (q, parts)
    assert ref[0] == (0, n - 1)              # the reference's quirk itself


def test_shard_arithmetic_matches_between_python_and_labels():
    labels = np.random.default_rng(0).permutation(1001).astype(np.float32)
    owned = [label_shard.shard_of_sorted_labels(labels, r, 3) for r in range(3)]
    assert [len(o) for o in owned] == [334, 334, 333]
    assert np.array_equal(np.sort(np.concatenate(owned)), np.arange(1001))


def test_make_labels_unique():
    rng = np.random.default_rng(1)
    lab = (1.2e9 + rng.integers(0, int(3.5e8), size=50_000)).astype(np.float32)
    assert len(np.unique(lab)) < len(lab)
    u = synth.make_labels_unique(lab)
    assert u.dtype == np.float32 and len(np.unique(u)) == len(u)
    order = np.argsort(lab, kind="stable")
    assert np.all(np.diff(u[order]) > 0)                     # same order, ties broken by original id
    assert np.max(np.abs(u.astype(np.float64) - lab)) <= 128 * 8   # a few ulps (128 at this magnitude)
    dense = synth.make_labels_unique(np.array([1, 1, 1, 1, 2, 2, 3], np.float32))
    assert len(np.unique(dense)) == 7 and np.all(np.diff(dense) > 0)
    already = np.arange(10, dtype=np.float32)
    assert np.array_equal(synth.make_labels_unique(already), already)


def test_adversarial_recipe_shapes():
    data, queries, labels, windows = synth.make_adversarial(20_000, 16, seed=3, clusters=10)
    assert data.shape == (20_000, 16) and queries.shape == (90, 16) and windows.shape == (90, 2)
    assert np.allclose(np.linalg.norm(data, axis=1), 1.0, atol=1e-5) and len(np.unique(labels)) == len(labels)
    in_window = (labels[None, :] >= windows[:, 0:1]) & (labels[None, :] <= windows[:, 1:2])
    assert np.all(np.abs(in_window.sum(1) - 2000) <= 2)      # a window is one cluster's label range
    w = synth.make_blowup_windows(labels, -4, 50, seed=1)
    srt = np.sort(labels)
    assert np.all(w[:, 0] <= srt[len(srt) // 2]) and np.all(w[:, 1] >= srt[len(srt) // 2 - 1])  # straddles the median


def test_bench_helpers():
    cfg = dict(bench.CONFIGS["c2"])
    assert bench.workload_string(cfg) == bench.workload_string(dict(cfg))      # both arms print this one string
    assert "rows scaled to 5" in bench.workload_string(dict(cfg, n=5, scaled=True))
    assert bench.auto_prefilter_route(100.0) == "prefilter_direct" and bench.auto_prefilter_route(768.0) == "prefilter_tc"
    assert bench.frac_name(-3) == "2^-3" and bench.frac_name("adv") == "adv"
    ids = np.array([[1, 2, 3]], np.uint32)
    d = np.array([[1e-7, 2e-7, 0.5]], np.float32)
    d2 = np.array([[1.3e-7, 2e-7, 0.5]], np.float32)
    assert not bench.rows_equal_up_to_ties(ids, d, ids, d2)[0]                 # 30 % apart: not a tie
    assert bench.rows_equal_up_to_ties(ids, d, ids, d2, scale=1.0)[0]          # inner products of unit vectors near 0
    swapped = np.array([[2, 1, 3]], np.uint32)
    tie = np.array([[0.25, 0.25, 0.5]], np.float32)
    assert bench.rows_equal_up_to_ties(swapped, tie, ids, tie)[0]              # ids differ only inside a distance tie
    assert not bench.rows_equal_up_to_ties(np.array([[1, 3, 2]], np.uint32), d, ids, d)[0]
