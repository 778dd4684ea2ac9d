"""More GPU parity against vectors of the unmodified reference, for behaviours the experiment driver does
not exercise:

  * B-WST split factors 3 and 4 (range_filter_tree.h:129-189): RangeFilterTreeIndexFloatEuclidian (prefilter
    buckets) against tests/golden/tiny_pretree_splits_ref_outputs.npz and, bit for bit, the device-order oracle;
    VamanaRangeFilterTreeIndexFloatEuclidian with split 3 on device-built graphs against the oracle
  * super-postfilter tree with split 2.5 / shift 0.4 (float-evaluated bucket sizes) on the reference-built
    graphs of tests/golden/tiny_super/
  * PrefilterIndex on duplicate labels with windows ending on label values (tests/golden/tiny_dup_ref_outputs.npz)

The CPU halves (oracle vs the same vectors; the device decomposition evaluated on the host vs the oracle's
trace for split 3 / 4 / 7) are tests/test_oracle_golden.py and tests/test_decompose_cpu.py.

(Sorted last on purpose: written after the round's GPU budget was spent; the first GPU run of this file is the
round-end one.)"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, device_graph_build
from golden_cases import TINY, tiny_cases
from oracle_api import Oracle
from rangefilteredann_b200 import synth
from test_gpu_golden import rows_equal_up_to_ties

pytestmark = pytest.mark.gpu
METHODS = ("fenwick", "optimized_postfilter", "three_split")


def _qp(engine, qkw):
    return engine.QueryParams(10, qkw["beam"], 1.35, 10_000_000, 10_000, qkw["mult"], qkw["max_beam"], qkw.get("ratio"), False)


@pytest.mark.parametrize("split", [3, 4])
def test_prefilter_bucket_tree_other_splits(engine, split):
    assert engine.device_count() > 0, "no CUDA device: the engine has no CPU fallback"
    data, queries, labels = synth.make_dataset(TINY["n"], TINY["d"], TINY["nq"], TINY["seed"])
    gold = np.load(os.path.join(GOLDEN, "tiny_pretree_splits_ref_outputs.npz"))
    tree = engine.RangeFilterTreeIndexFloatEuclidian(data, labels, 300, split, engine.BuildParams(64, 500, 1.0, ""))
    orc = Oracle("pretree", data, labels, None, dist_mode=1, cutoff=300, split=float(split))
    for name, windows, qkw in tiny_cases(labels):
        nq = len(windows)
        q = queries[:nq]
        for m in METHODS:
            ids, d = tree.batch_search(q, windows, nq, m, _qp(engine, qkw))
            oids, od = orc.batch(m, q, windows, k=10, beam=qkw["beam"], mult=qkw["mult"], max_beam=qkw["max_beam"],
                                 ratio=qkw.get("ratio"), pad_id=0)
            assert np.array_equal(d.view(np.uint32), od.view(np.uint32)) and np.array_equal(ids, oids), f"b{split}/{name}/{m} vs oracle"
            key = f"b{split}/{name}/{m}"
            ok = rows_equal_up_to_ties(ids, d, gold[key + "/ids"], gold[key + "/dists"])
            assert ok.all(), f"{key}: rows {np.nonzero(~ok)[0][:8]} differ from the reference"


def test_vamana_bucket_tree_split_3(engine, tmp_path):
    data, queries, labels = synth.make_dataset(TINY["n"], TINY["d"], TINY["nq"], TINY["seed"])
    cache = str(tmp_path / "wst3") + "/"
    # cutoff 400: rows of 1, 3 and 9 buckets (3000 -> 1000 -> 334 / 333 points)
    with device_graph_build():
        tree = engine.VamanaRangeFilterTreeIndexFloatEuclidian(data, labels, 400, 3, engine.BuildParams(64, 500, 1.0, cache))
    assert len([f for f in os.listdir(cache) if f.endswith(".bin")]) == 1 + 3 + 9
    orc = Oracle("wst", data, labels, cache, dist_mode=1, cutoff=400, split=3.0)
    for name, windows, qkw in tiny_cases(labels):
        nq = len(windows)
        q = queries[:nq]
        for m in METHODS:
            ids, d = tree.batch_search(q, windows, nq, m, _qp(engine, qkw))
            oids, od = orc.batch(m, q, windows, k=10, beam=qkw["beam"], mult=qkw["mult"], max_beam=qkw["max_beam"],
                                 ratio=qkw.get("ratio"), pad_id=0)
            assert np.array_equal(d.view(np.uint32), od.view(np.uint32)), f"{name}/{m}: distances differ"
            assert np.array_equal(ids, oids), f"{name}/{m}: ids differ"


def test_super_tree_fractional_split(engine):
    """SuperOptimizedPostfilterTreeIndexFloatEuclidian with split 2.5 / shift 0.4 on the 20 reference-built
    graphs of tests/golden/tiny_super/: the host geometry (bucket sizes evaluated in float, as
    super_optimized_postfilter_tree.h:145-170 does) must name exactly those files — a mismatch would show up
    as a device-side build of new files — and the rows must match the oracle bit for bit and the reference's
    golden vectors."""
    from golden_cases import TINY_SUPER, tiny_super_cases
    c = TINY_SUPER
    data, queries, labels = synth.make_dataset(c["n"], c["d"], c["nq"], c["seed"])
    cache = os.path.join(GOLDEN, "tiny_super") + "/"
    before = sorted(os.listdir(cache))
    sup = engine.SuperOptimizedPostfilterTreeIndexFloatEuclidian(data, labels, c["cutoff"], c["split"], c["shift"],
                                                                 engine.BuildParams(64, 500, 1.0, cache))
    assert sorted(os.listdir(cache)) == before and len(before) == 20, "the engine's geometry named other graph files"
    gold = np.load(os.path.join(GOLDEN, "tiny_super_ref_outputs.npz"))
    orc = Oracle("super", data, labels, cache, dist_mode=1, cutoff=c["cutoff"], split=c["split"], shift=c["shift"])
    for name, windows, qkw in tiny_super_cases(labels):
        nq = len(windows)
        ids, d = sup.batch_search(queries[:nq], windows, nq, _qp(engine, qkw))
        oids, od = orc.batch("super", queries[:nq], windows, k=10, beam=qkw["beam"], mult=qkw["mult"], max_beam=qkw["max_beam"], pad_id=0)
        assert np.array_equal(d.view(np.uint32), od.view(np.uint32)) and np.array_equal(ids, oids), f"{name} vs oracle"
        ok = rows_equal_up_to_ties(ids, d, gold[f"{name}/super/ids"], gold[f"{name}/super/dists"])
        assert ok.mean() >= 0.95, f"{name}: only {ok.sum()}/{len(ok)} rows match the reference"


@pytest.mark.parametrize("k", [1, 5])
def test_prefilter_duplicate_labels(engine, k):
    """PrefilterIndex on ~12 points per label value with windows that end exactly on label values, against the
    reference's golden vectors (tests/golden/tiny_dup_ref_outputs.npz), through the task path and the one-launch
    kernel.  Rows the reference leaves undefined (fewer than k points in
    the window; windows reaching past the largest label, where its unstable sort decides which point is
    dropped) are skipped exactly as in tests/test_oracle_golden.py."""
    from golden_cases import tiny_dup_dataset, tiny_dup_windows
    from rangefilteredann_b200 import capi
    from test_oracle_golden import dup_window_sizes
    data, queries, labels = tiny_dup_dataset()
    w = tiny_dup_windows()
    gold = np.load(os.path.join(GOLDEN, "tiny_dup_ref_outputs.npz"))
    defined = (dup_window_sizes(labels, w) >= k) & (w[:, 1] <= labels.max())
    pre = engine.PrefilterIndexFloatEuclidian(data, labels)
    h = capi.Handle.borrow(pre)
    qp = engine.QueryParams(k, 10, 1.35, 10_000_000, 10_000, 1, 10000, None, False)
    rids, rd = gold[f"k{k}/ids"], gold[f"k{k}/dists"]
    h.set_option("gemm_prefilter", 0)
    for direct in (0, 1):
        h.set_option("prefilter_direct", direct)
        ids, d = pre.batch_search(queries, w, len(w), qp)
        ok = rows_equal_up_to_ties(ids[defined], d[defined], rids[defined], rd[defined])
        assert ok.all(), f"direct={direct}: rows {np.nonzero(~ok)[0][:8]} differ from the reference"
        lab = labels[ids[defined].astype(np.int64)]
        assert (lab >= w[defined, 0:1]).all() and (lab < w[defined, 1:2]).all()
        assert (ids[3] == 0xFFFFFFFF).all()  # empty window: pads
