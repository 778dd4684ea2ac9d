"""GPU: several arenas behind one batch_search call (ws_group, include/wsann.h) — on the single-GPU box the members
share device 0, which exercises everything but the physical links: replication device to device, one host thread
per member, slicing of the batch, label shards with their own trees, the peer-load gather+merge kernel.  (The NCCL
exchange needs distinct devices: tests/multigpu_group.py, run with gpurun --gpus 2.)

  T7  query-sharded: rows identical to 1 GPU;  label-sharded: prefilter ids identical,
                         graph methods recall >= single-GPU recall - 0.005
"""
import contextlib
import os

import numpy as np
import pytest

from conftest import GOLDEN, device_graph_build
from golden_cases import TINY
from rangefilteredann_b200 import capi, synth

pytestmark = pytest.mark.gpu


@contextlib.contextmanager
def devices(spec, mode=None):
    old = {k: os.environ.get(k) for k in ("WSANN_DEVICES", "WSANN_SHARD_MODE")}
    os.environ["WSANN_DEVICES"] = spec
    if mode:
        os.environ["WSANN_SHARD_MODE"] = mode
    else:
        os.environ.pop("WSANN_SHARD_MODE", None)
    try:
        yield
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def qp_of(engine, beam, mult=1, k=10):
    return engine.QueryParams(k, beam, 1.35, 10_000_000, 10_000, mult, 10000, None, False)


@pytest.fixture(scope="module")
def tiny():
    data, queries, labels = synth.make_dataset(TINY["n"], TINY["d"], TINY["nq"], TINY["seed"])
    return data, queries, labels


def same(a, b):
    return np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))


def test_replicas_answer_slices_with_identical_rows(engine, tiny):
    data, queries, labels = tiny
    cache = os.path.join(GOLDEN, "tiny", "wst") + "/"
    bp = engine.BuildParams(64, 500, 1.0, cache)
    single = engine.VamanaRangeFilterTreeIndexFloatEuclidian(data, labels, TINY["cutoff"], 2, bp)
    pre1 = engine.PrefilterIndexFloatEuclidian(data, labels)
    with devices("0,0,0"):
        tree3 = engine.VamanaRangeFilterTreeIndexFloatEuclidian(data, labels, TINY["cutoff"], 2, bp)
        pre3 = engine.PrefilterIndexFloatEuclidian(data, labels)
    g = capi.Group.borrow(tree3)
    assert g is not None and g.size() == 3 and capi.Group.borrow(single) is None
    for power in (-8, -4, -1, 0):
        w = synth.make_windows(labels, power, len(queries), seed=500 + power)
        for method, beam, mult in (("fenwick", 10, 1), ("optimized_postfilter", 20, 2), ("three_split", 10, 1)):
            a = single.batch_search(queries, w, len(w), method, qp_of(engine, beam, mult))
            b = tree3.batch_search(queries, w, len(w), method, qp_of(engine, beam, mult))
            assert same(a, b), (power, method)
        assert same(pre1.batch_search(queries, w, len(w), qp_of(engine, 10)), pre3.batch_search(queries, w, len(w), qp_of(engine, 10)))
    # batches smaller than the group, and a ragged split
    for nq in (1, 2, 7):
        w = synth.make_windows(labels, -2, nq, seed=9)
        assert same(single.batch_search(queries[:nq], w, nq, "fenwick", qp_of(engine, 10)),
                    tree3.batch_search(queries[:nq], w, nq, "fenwick", qp_of(engine, 10)))
    # members keep their own counters: all of them worked
    launches = [g.member(i).launches() for i in range(3)]
    assert all(x > 0 for x in launches)


def test_replicated_super_tree_and_flat_index(engine, tiny, tmp_path):
    data, queries, labels = tiny
    sup_cache, flat_cache = os.path.join(GOLDEN, "tiny", "super") + "/", os.path.join(GOLDEN, "tiny", "flat") + "/"
    if not (os.path.isdir(sup_cache) and os.path.isdir(flat_cache)):
        sup_cache, flat_cache = str(tmp_path / "super") + "/", str(tmp_path / "flat") + "/"
    with device_graph_build():
        s1 = engine.SuperOptimizedPostfilterTreeIndexFloatEuclidian(data, labels, TINY["cutoff"], 2.0, 0.5, engine.BuildParams(64, 500, 1.0, sup_cache))
        f1 = engine.PostfilterVamanaIndexFloatEuclidian(data, labels, engine.BuildParams(64, 500, 1.0, flat_cache))
        with devices("0,0"):
            s2 = engine.SuperOptimizedPostfilterTreeIndexFloatEuclidian(data, labels, TINY["cutoff"], 2.0, 0.5, engine.BuildParams(64, 500, 1.0, sup_cache))
            f2 = engine.PostfilterVamanaIndexFloatEuclidian(data, labels, engine.BuildParams(64, 500, 1.0, flat_cache))
    w = synth.make_windows(labels, -3, len(queries), seed=77)
    assert same(s1.batch_search(queries, w, len(w), qp_of(engine, 20, 2)), s2.batch_search(queries, w, len(w), qp_of(engine, 20, 2)))
    assert same(f1.batch_search(queries, w, len(w), qp_of(engine, 40, 2)), f2.batch_search(queries, w, len(w), qp_of(engine, 40, 2)))


def test_label_shards_prefilter_rows_identical(engine):
    """PrefilterIndex over 4 label shards: every row identical to the single arena's, including windows that span
    shard boundaries, windows that end on the data set's last point (the reference's r = n-1 rule applies to it and
    to no shard's last point) and every routing (one-launch kernel, task path, tensor-core sweep)."""
    data, queries, labels = synth.make_dataset(40_000, 64, 600, 21)
    single = engine.PrefilterIndexFloatEuclidian(data, labels)
    with devices("0,0,0,0", "label"):
        sharded = engine.PrefilterIndexFloatEuclidian(data, labels)
    g = capi.Group.borrow(sharded)
    assert g.size() == 4 and g.info()["exchange"] == "peer_loads"
    srt = np.sort(labels)
    for power in (-12, -8, -5, -2, -1, 0):
        w = synth.make_windows(labels, power, len(queries), seed=600 + power)
        w[0] = (srt[9_990], srt[10_020])          # across the first shard boundary
        w[1] = (srt[-50], srt[-1] + 1.0)          # through the last point
        w[2] = (srt[19_999], srt[20_000])         # one point on each side of a boundary
        w[3] = (srt[5], srt[5])                   # empty
        for k in (10, 1):
            a = single.batch_search(queries, w, len(w), qp_of(engine, 10, k=k))
            b = sharded.batch_search(queries, w, len(w), qp_of(engine, 10, k=k))
            assert same(a, b), (power, k, np.nonzero((a[0] != b[0]).any(1))[0][:5])
    assert g.info()["total_ms"] > 0


def test_label_shards_tree_recall(engine, tiny, tmp_path):
    data, queries, labels = tiny
    single = engine.VamanaRangeFilterTreeIndexFloatEuclidian(data, labels, TINY["cutoff"], 2,
                                                             engine.BuildParams(64, 500, 1.0, os.path.join(GOLDEN, "tiny", "wst") + "/"))
    with device_graph_build(), devices("0,0", "label"):
        sharded = engine.VamanaRangeFilterTreeIndexFloatEuclidian(data, labels, TINY["cutoff"], 2,
                                                                  engine.BuildParams(64, 500, 1.0, str(tmp_path / "shards") + "/"))
    assert capi.Group.borrow(sharded).size() == 2
    for power in (-6, -3, -1, 0):
        w = synth.make_windows(labels, power, len(queries), seed=700 + power)
        gt = synth.ground_truth(data, queries, labels, w)
        for method, beam in (("fenwick", 20), ("optimized_postfilter", 20)):
            ids1, _ = single.batch_search(queries, w, len(w), method, qp_of(engine, beam))
            ids2, d2 = sharded.batch_search(queries, w, len(w), method, qp_of(engine, beam))
            assert synth.recall_std(ids2, gt) >= synth.recall_std(ids1, gt) - 0.005, (power, method)
            assert (np.diff(d2, axis=1) >= 0).all()
            pads = d2 == np.float32(3.4028235e38)
            assert (ids2[pads] == 0).all()


def test_group_through_the_c_abi(engine, tiny):
    """ws_index_replicate + ws_group_create called directly (what a host program embedding the library does)."""
    data, queries, labels = tiny
    pre = engine.PrefilterIndexFloatEuclidian(data, labels)
    h = capi.Handle.borrow(pre)
    rep = h.replicate(0)
    assert rep.hbm_bytes() > 0
    g = capi.Group.create([h, rep], capi.GROUP_REPLICATED)
    w = synth.make_windows(labels, -4, len(queries), seed=3)
    ids, d = np.empty((len(w), 10), np.uint32), np.empty((len(w), 10), np.float32)
    g.prefilter_batch(queries, w, len(w), 10, ids, d)
    assert same((ids, d), pre.batch_search(queries, w, len(w), qp_of(engine, 10)))
    with pytest.raises(capi.WsError):
        capi.Group.create([h, rep], 7)
    del g
