"""GPU: arena snapshot (ws_index_save / ws_index_load).  The reference persists graphs only
(postfilter_vamana.h:54-79); a snapshot is the finished arena in one file.  A loaded index must answer every method
with rows bit-identical to the index it was saved from, without the points, the labels or the graph cache."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, device_graph_build
from golden_cases import TINY
from rangefilteredann_b200 import capi, synth

pytestmark = pytest.mark.gpu


def qp_of(engine, beam, mult=1, k=10):
    return engine.QueryParams(k, beam, 1.35, 10_000_000, 10_000, mult, 10000, None, False)


def same(a, b):
    return np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))


@pytest.fixture(scope="module")
def tiny():
    return synth.make_dataset(TINY["n"], TINY["d"], TINY["nq"], TINY["seed"])


def test_tree_snapshot_round_trip(engine, tiny, tmp_path):
    data, queries, labels = tiny
    tree = engine.VamanaRangeFilterTreeIndexFloatEuclidian(data, labels, TINY["cutoff"], 2,
                                                           engine.BuildParams(64, 500, 1.0, os.path.join(GOLDEN, "tiny", "wst") + "/"))
    path = str(tmp_path / "tree.wsann")
    tree.save_snapshot(path)
    assert os.path.getsize(path) > data.nbytes
    loaded = engine.VamanaRangeFilterTreeIndexFloatEuclidian.load_snapshot(path)
    assert capi.Handle.borrow(loaded).hbm_bytes() > 0
    for power in (-8, -4, -1, 0):
        w = synth.make_windows(labels, power, len(queries), seed=800 + power)
        for method, beam, mult in (("fenwick", 10, 1), ("optimized_postfilter", 20, 2), ("three_split", 10, 1)):
            assert same(tree.batch_search(queries, w, len(w), method, qp_of(engine, beam, mult)),
                        loaded.batch_search(queries, w, len(w), method, qp_of(engine, beam, mult))), (power, method)
    # a snapshot replicated over "two GPUs" at load time
    os.environ["WSANN_DEVICES"] = "0,0"
    try:
        loaded2 = engine.VamanaRangeFilterTreeIndexFloatEuclidian.load_snapshot(path)
    finally:
        del os.environ["WSANN_DEVICES"]
    assert capi.Group.borrow(loaded2).size() == 2
    w = synth.make_windows(labels, -3, len(queries), seed=5)
    assert same(tree.batch_search(queries, w, len(w), "fenwick", qp_of(engine, 10)), loaded2.batch_search(queries, w, len(w), "fenwick", qp_of(engine, 10)))


def test_prefilter_super_flat_and_pretree_snapshots(engine, tiny, tmp_path):
    data, queries, labels = tiny
    w = synth.make_windows(labels, -3, len(queries), seed=31)
    pre = engine.PrefilterIndexFloatEuclidian(data, labels)
    pre.save_snapshot(str(tmp_path / "pre.wsann"))
    pre2 = engine.PrefilterIndexFloatEuclidian.load_snapshot(str(tmp_path / "pre.wsann"))
    assert same(pre.batch_search(queries, w, len(w), qp_of(engine, 10)), pre2.batch_search(queries, w, len(w), qp_of(engine, 10)))
    ptree = engine.RangeFilterTreeIndexFloatEuclidian(data, labels, TINY["cutoff"], 2)
    ptree.save_snapshot(str(tmp_path / "ptree.wsann"))
    ptree2 = engine.RangeFilterTreeIndexFloatEuclidian.load_snapshot(str(tmp_path / "ptree.wsann"))
    assert same(ptree.batch_search(queries, w, len(w), "fenwick", qp_of(engine, 10)), ptree2.batch_search(queries, w, len(w), "fenwick", qp_of(engine, 10)))
    with device_graph_build():
        sup = engine.SuperOptimizedPostfilterTreeIndexFloatMips(data, labels, TINY["cutoff"], 2.0, 0.5, engine.BuildParams(64, 500, 1.0, str(tmp_path / "s") + "/"))
        flat = engine.PostfilterVamanaIndexFloatEuclidian(data, labels, engine.BuildParams(64, 500, 1.0, str(tmp_path / "f") + "/"))
    sup.save_snapshot(str(tmp_path / "sup.wsann"))
    flat.save_snapshot(str(tmp_path / "flat.wsann"))
    sup2 = engine.SuperOptimizedPostfilterTreeIndexFloatMips.load_snapshot(str(tmp_path / "sup.wsann"))
    flat2 = engine.PostfilterVamanaIndexFloatEuclidian.load_snapshot(str(tmp_path / "flat.wsann"))
    assert same(sup.batch_search(queries, w, len(w), qp_of(engine, 20, 2)), sup2.batch_search(queries, w, len(w), qp_of(engine, 20, 2)))
    assert same(flat.batch_search(queries, w, len(w), qp_of(engine, 40, 2)), flat2.batch_search(queries, w, len(w), qp_of(engine, 40, 2)))


def test_snapshot_errors(engine, tiny, tmp_path):
    data, queries, labels = tiny
    with pytest.raises(RuntimeError, match="cannot open"):
        engine.PrefilterIndexFloatEuclidian.load_snapshot(str(tmp_path / "missing.wsann"))
    bad = tmp_path / "bad.wsann"
    bad.write_bytes(b"not a snapshot at all")
    with pytest.raises(RuntimeError, match="not an arena snapshot"):
        engine.PrefilterIndexFloatEuclidian.load_snapshot(str(bad))
    pre = engine.PrefilterIndexFloatEuclidian(data, labels)
    good = tmp_path / "good.wsann"
    pre.save_snapshot(str(good))
    cut = tmp_path / "cut.wsann"
    cut.write_bytes(good.read_bytes()[: os.path.getsize(good) // 2])
    with pytest.raises(RuntimeError, match="truncated"):
        engine.PrefilterIndexFloatEuclidian.load_snapshot(str(cut))
