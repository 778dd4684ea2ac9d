"""GPU parity of VamanaRangeFilterTreeIndexUInt8Euclidian against golden vectors of the unmodified
reference on the reference-built graphs under tests/golden/tiny_u8/wst/ (tests/golden/make_golden.py
--only u8).  Same graphs, same integer-valued distances, ties broken by id on both sides
(beamSearch.h:59-61): distances bit-identical, ids identical up to exact ties.  The oracle is pinned to the
same vectors on CPU (tests/test_oracle_golden.py::test_oracle_matches_reference_8bit_graph_tree).

(Own file, sorted last: it was written after the round's GPU budget was spent, so its first GPU run is the
round-end one — placed where a surprise cannot mask the rest of the suite under `-x`.)"""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from golden_cases import TINY_U8, tiny_u8_cases, tiny_u8_dataset
from test_gpu_8bit import _assert_rows

pytestmark = pytest.mark.gpu


def test_golden_8bit_graph_tree(engine):
    """Same graphs as the reference (tests/golden/tiny_u8/wst/), same integer-valued distances, ties
    broken by id on both sides (beamSearch.h:59-61)."""
    data, queries, labels = tiny_u8_dataset(False)
    gold = np.load(os.path.join(GOLDEN, "tiny_u8_ref_outputs.npz"))
    cache = os.path.join(GOLDEN, "tiny_u8", "wst") + "/"
    tree = engine.VamanaRangeFilterTreeIndexUInt8Euclidian(data, labels, TINY_U8["cutoff"], 2, engine.BuildParams(64, 500, 1.0, cache))
    for name, windows, qkw in tiny_u8_cases(labels):
        nq = len(windows)
        qp = engine.QueryParams(10, qkw["beam"], 1.35, 10_000_000, 10_000, qkw["mult"], qkw["max_beam"], None, False)
        for m in ("fenwick", "optimized_postfilter", "three_split"):
            ids, d = tree.batch_search(queries[:nq], windows, nq, m, qp)
            _assert_rows(ids, d, gold[f"UInt8Euclidian/{name}/{m}/ids"], gold[f"UInt8Euclidian/{name}/{m}/dists"], f"{name}/{m}")
