"""GPU parity for RangeFilterTreeIndex* — the B-WST whose buckets are PrefilterIndex sub-indices
(python_bindings.cpp:119-127; range_filter_tree.h:32 default template argument).

  * golden vectors from the unmodified reference (tests/golden/tiny_pretree_ref_outputs.npz):
    every row identical up to distance ties within 1e-5 relative — this path is exact search, so
    unlike the graph methods there is no traversal that a rounding difference could steer
  * bit-exact ids + fp32 distances against the device-order oracle, tiny and a 60 000-point
    tree whose bucket scans are cut into scan_chunk pieces, L2 and MIPS
  * size-independent property at that size: fenwick over prefilter buckets returns the exact
    top-k of the window minus the last point of each cover bucket, so it can never beat, and is
    almost always equal to, the brute-force top-k
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from golden_cases import TINY, tiny_cases
from oracle_api import Oracle
from rangefilteredann_b200 import synth
from test_gpu_golden import rows_equal_up_to_ties

pytestmark = pytest.mark.gpu
METHODS = ["fenwick", "optimized_postfilter", "three_split"]
FLT_MAX = np.float32(3.4028235e38)


def _qp(engine, qkw, k=10):
    return engine.QueryParams(k, qkw["beam"], 1.35, 10_000_000, 10_000, qkw["mult"], qkw["max_beam"], qkw.get("ratio"), False)


@pytest.fixture(scope="module")
def tiny(engine):
    assert engine.device_count() > 0, "no CUDA device: the engine has no CPU fallback"
    data, queries, labels = synth.make_dataset(TINY["n"], TINY["d"], TINY["nq"], TINY["seed"])
    tree = engine.RangeFilterTreeIndexFloatEuclidian(data, labels, TINY["cutoff"], 2, engine.BuildParams(64, 500, 1.0, ""))
    orc = Oracle("pretree", data, labels, None, dist_mode=1, cutoff=TINY["cutoff"])
    return dict(tree=tree, orc=orc, queries=queries, labels=labels)


@pytest.mark.parametrize("method", METHODS)
def test_golden_prefilter_nodes(engine, tiny, method):
    gold = np.load(os.path.join(GOLDEN, "tiny_pretree_ref_outputs.npz"))
    checked = 0
    for name, windows, qkw in tiny_cases(tiny["labels"]):
        key = f"{name}/{method}/ids"
        if key not in gold:
            continue
        nq = len(windows)
        ids, d = tiny["tree"].batch_search(tiny["queries"][:nq], windows, nq, method, _qp(engine, qkw))
        assert ids.dtype == np.uint32 and d.dtype == np.float32 and ids.shape == gold[key].shape
        ok = rows_equal_up_to_ties(ids, d, gold[key], gold[f"{name}/{method}/dists"])
        assert ok.all(), f"{name}/{method}: rows {np.nonzero(~ok)[0][:8]} differ from the reference"
        checked += 1
    assert checked > 0


@pytest.mark.parametrize("method", METHODS)
def test_bit_exact_vs_oracle_tiny(engine, tiny, method):
    for name, windows, qkw in tiny_cases(tiny["labels"]):
        nq = len(windows)
        q = tiny["queries"][:nq]
        ids, d = tiny["tree"].batch_search(q, windows, nq, method, _qp(engine, qkw))
        oids, od = tiny["orc"].batch(method, q, windows, k=10, beam=qkw["beam"], mult=qkw["mult"],
                                     max_beam=qkw["max_beam"], ratio=qkw.get("ratio"), pad_id=0)
        assert np.array_equal(d.view(np.uint32), od.view(np.uint32)), f"{name}/{method}"
        assert np.array_equal(ids, oids), f"{name}/{method}"


@pytest.mark.parametrize("angular", [False, True])
def test_chunked_bucket_scans(engine, angular):
    """60 000 x 24 (rows padded to 32 floats): the upper buckets are longer than scan_chunk (8192)."""
    n, d, nq = 60_000, 24, 96
    data, queries, labels = synth.make_dataset(n, d, nq, 21, angular)
    sfx = "FloatMips" if angular else "FloatEuclidian"
    tree = getattr(engine, "RangeFilterTreeIndex" + sfx)(data, labels, 1000, 2, engine.BuildParams(64, 500, 1.0, ""))
    orc = Oracle("pretree", data, labels, None, metric=1 if angular else 0, dist_mode=1, cutoff=1000)
    for power in (-10, -4, -1, 0):
        w = synth.make_windows(labels, power, nq, seed=300 + power)
        for method in METHODS:
            for k in (1, 10, 37):
                qp = engine.QueryParams(k, 10, 1.35, 10_000_000, 10_000, 1, 10000, None, False)
                ids, dd = tree.batch_search(queries, w, nq, method, qp)
                oids, od = orc.batch(method, queries, w, k=k, pad_id=0)
                assert np.array_equal(dd.view(np.uint32), od.view(np.uint32)), (power, method, k)
                assert np.array_equal(ids, oids), (power, method, k)
        if not angular:
            # exact search minus one point per cover bucket: never better than brute force, nearly always equal
            gt = synth.ground_truth(data, queries, labels, w)
            qp = engine.QueryParams(10, 10, 1.35, 10_000_000, 10_000, 1, 10000, None, False)
            ids, dd = tree.batch_search(queries, w, nq, "fenwick", qp)
            hit = np.mean([len(set(ids[i][dd[i] < FLT_MAX].tolist()) & set(gt[i][gt[i] >= 0].tolist())) /
                           max(1, (gt[i] >= 0).sum()) for i in range(nq)])
            assert hit > 0.97, (power, hit)


def test_empty_and_padding(engine, tiny):
    hi = float(tiny["labels"].max())
    w = np.array([[hi + 1.0, hi + 2.0]] * 3, np.float32)
    qp = engine.QueryParams(10, 10, 1.35, 10_000_000, 10_000, 1, 10000, None, False)
    for method in METHODS:
        ids, d = tiny["tree"].batch_search(tiny["queries"][:3], w, 3, method, qp)
        assert (ids == 0).all() and (d == FLT_MAX).all()  # range_filter_tree.h:89-92
        ids, d = tiny["tree"].batch_search(tiny["queries"][:0], w[:0], 0, method, qp)
        assert ids.shape == (0, 10)
