"""BASELINE.json's other configurations at a size the oracle finishes in seconds, over ALL 17 filter
fractions 2^-16..2^0 (north_star: "graph-search recall@10 within 0.5 points of the reference at equal
beam width across all 17 filter fractions"):

  C3 shape  d = 100 angular (rows padded to 112 floats), super-postfilter tree
  C5 shape  d = 96  L2, 2-WST (fenwick + optimized postfilter)
  C2 shape  d = 128 L2, 2-WST
  C4 shape  d = 512 angular (CLIP-like), timestamp-style labels with many duplicates
            (generate_redcaps_data.py:77-80), 2-WST

65 536 points, graphs built on the device and saved in the reference's format; the oracle (pinned to
the reference bit for bit, tests/test_oracle_golden.py) loads the same files.  Per fraction: ids and
fp32 distances identical to the device-order oracle, hence recall identical; the recall difference is
asserted anyway, against an independent brute-force ground truth.
"""
import numpy as np
import pytest

from conftest import device_graph_build
from oracle_api import Oracle
from rangefilteredann_b200 import synth

pytestmark = pytest.mark.gpu
N, NQ, K = 65536, 96, 10
POWERS = list(range(-16, 1))


def recall(ids, gt):
    valid = gt >= 0
    hit = ((gt[:, :, None] == ids[:, None, :].astype(np.int64)).any(2) & valid).sum(1)
    return float(np.mean(hit / np.maximum(valid.sum(1), 1)))


@pytest.mark.parametrize("d,angular,kind,label_kind", [(100, True, "super", "unique"), (96, False, "wst", "unique"),
                                                       (128, False, "wst", "unique"), (512, True, "wst", "timestamp")])
def test_all_17_fractions(engine, tmp_path, d, angular, kind, label_kind):
    assert engine.device_count() > 0, "no CUDA device: the engine has no CPU fallback"
    data, queries, labels = synth.make_dataset(N, d, NQ, 31 + d, angular, label_kind=label_kind)
    cache = str(tmp_path / kind) + "/"
    sfx = "FloatMips" if angular else "FloatEuclidian"
    bp = engine.BuildParams(64, 500, 1.0, cache)
    metric = 1 if angular else 0
    if kind == "super":
        with device_graph_build():
            idx = getattr(engine, "SuperOptimizedPostfilterTreeIndex" + sfx)(data, labels, 1000, 2.0, 0.5, bp)
        orc = Oracle("super", data, labels, cache, metric=metric, dist_mode=1, cutoff=1000)
        methods = ["super"]
    else:
        with device_graph_build():
            idx = getattr(engine, "VamanaRangeFilterTreeIndex" + sfx)(data, labels, 1000, 2, bp)
        orc = Oracle("wst", data, labels, cache, metric=metric, dist_mode=1, cutoff=1000)
        methods = ["fenwick", "optimized_postfilter"]
    qp = engine.QueryParams(K, 20, 1.35, 10_000_000, 10_000, 2, 10000, None, False)
    worst = 0.0
    for p in POWERS:
        w = synth.make_windows(labels, p, NQ, seed=500 + p)
        gt = synth.ground_truth(data, queries, labels, w, k=K, angular=angular)
        for m in methods:
            if m == "super":
                ids, dist = idx.batch_search(queries, w, NQ, qp)
            else:
                ids, dist = idx.batch_search(queries, w, NQ, m, qp)
            oids, od = orc.batch(m, queries, w, k=K, beam=20, mult=2, pad_id=0)
            assert np.array_equal(dist.view(np.uint32), od.view(np.uint32)), f"2^{p} {m}: distances differ"
            assert np.array_equal(ids, oids), f"2^{p} {m}: ids differ"
            # pads are id 0 with FLT_MAX: mask them before counting hits
            real = dist < np.finfo(np.float32).max
            r_e = recall(np.where(real, ids, 0xFFFFFFFF), gt)
            r_o = recall(np.where(od < np.finfo(np.float32).max, oids, 0xFFFFFFFF), gt)
            worst = max(worst, abs(r_e - r_o))
            assert abs(r_e - r_o) <= 0.005, f"2^{p} {m}: recall {r_e:.4f} vs reference-order {r_o:.4f}"
            assert r_e >= 0.5, f"2^{p} {m}: recall {r_e:.4f} at beam 20 x2 is implausibly low"
    print(kind, d, "max |recall difference| over 17 fractions:", worst)
