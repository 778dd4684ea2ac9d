"""GPU: device merge of per-shard partial top-k lists (ws_merge_partial_topk) and the
label-range sharded search on one device (two shards built in-process, merged on the GPU)."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import GOLDEN, device_graph_build
from golden_cases import TINY
from rangefilteredann_b200 import capi, label_shard, synth

pytestmark = pytest.mark.gpu


def test_device_merge_matches_numpy(engine):
    data, queries, labels = synth.make_dataset(TINY["n"], TINY["d"], TINY["nq"], TINY["seed"])
    pre = engine.PrefilterIndexFloatEuclidian(data, labels)
    h = capi.Handle.borrow(pre)
    rng = np.random.default_rng(0)
    for parts, nq, k in ((2, 64, 10), (8, 257, 10), (3, 50, 100), (5, 33, 1)):
        ids = rng.integers(0, 1 << 20, size=(parts, nq, k)).astype(np.uint32)
        d = np.sort(rng.uniform(size=(parts, nq, k)).astype(np.float32), axis=2)
        short = rng.integers(0, k + 1, size=(parts, nq))
        for p in range(parts):
            for q in range(nq):
                d[p, q, short[p, q]:] = label_shard.FLT_MAX
        ids[d == label_shard.FLT_MAX] = 0
        di, dd = h.dalloc(ids.nbytes), h.dalloc(d.nbytes)
        oi, od = h.dalloc(nq * k * 4), h.dalloc(nq * k * 4)
        h.h2d(di, ids); h.h2d(dd, d)
        h.merge_partial_topk(di, dd, parts, nq, k, 0, oi, od)
        h.sync()
        gi = np.empty((nq, k), np.uint32); gd = np.empty((nq, k), np.float32)
        h.d2h(gi, oi); h.d2h(gd, od)
        ei, ed = label_shard.merge_partial_topk_numpy(ids, d, k)
        assert np.array_equal(gd, ed) and np.array_equal(gi, ei), (parts, nq, k)
        for p_ in (di, dd, oi, od):
            h.dfree(p_)


def test_two_label_shards_on_one_device(engine, tmp_path):
    """Shard the tiny dataset by label range into two sub-trees (graphs built on the device),
    answer on both, merge: the fenwick result must not lose recall against the single tree."""
    data, queries, labels = synth.make_dataset(TINY["n"], TINY["d"], TINY["nq"], TINY["seed"])
    single = engine.VamanaRangeFilterTreeIndexFloatEuclidian(
        data, labels, TINY["cutoff"], 2, engine.BuildParams(64, 500, 1.0, os.path.join(GOLDEN, "tiny", "wst") + "/"))
    with device_graph_build():
        shards = [label_shard.LabelShardedTree(data, labels, r, 2, str(tmp_path), cutoff=TINY["cutoff"]) for r in range(2)]
    qp = engine.QueryParams(10, 20, 1.35, 10_000_000, 10_000, 1, 10000, None, False)
    for power in (-4, -1, 0):
        w = synth.make_windows(labels, power, TINY["nq"], seed=400 + power)
        gt = synth.ground_truth(data, queries, labels, w)
        parts = [s.local_search(queries, w, "fenwick", qp) for s in shards]
        mi, md = label_shard.merge_partial_topk_numpy(np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts]), 10)
        r_shard = synth.recall_std(mi, gt)
        r_single = synth.recall_std(single.batch_search(queries, w, len(w), "fenwick", qp)[0], gt)
        assert r_shard >= r_single - 0.005, (power, r_shard, r_single)
        assert (np.diff(md, axis=1) >= 0).all()
