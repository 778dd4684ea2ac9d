"""GPU parity for the UInt8 / Int8 classes (python_bindings.cpp:74-86,233-236).

The reference computes their distances as int32 sums cast to float; this engine keeps one fp32 arena
and is exact on it while dim * 255^2 < 2^24 (checked at construction).  Covered here:
  * golden vectors from the unmodified reference — PrefilterIndex*, RangeFilterTreeIndex* (prefilter
    buckets): distances bit-identical, ids identical up to exact ties
    (VamanaRangeFilterTreeIndexUInt8Euclidian on reference-built graphs: tests/test_gpu_zz_8bit_reference_graphs.py)
  * every graph class of three 8-bit variants on graphs built on the device, saved in the reference's
    .bin format and re-loaded by the oracle: ids and distances bit-identical
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, device_graph_build
from golden_cases import TINY_U8, tiny_u8_cases, tiny_u8_dataset
from oracle_api import Oracle

pytestmark = pytest.mark.gpu


def _assert_rows(ids, d, rids, rd, what):
    assert ids.dtype == np.uint32 and d.dtype == np.float32
    assert np.array_equal(d.view(np.uint32), rd.view(np.uint32)), f"{what}: distances differ"
    for i, j in zip(*np.nonzero(ids != rids)):
        assert (rd[i] == rd[i, j]).sum() > 1 or rd[i, j] == rd[i, -1], f"{what}: row {i} col {j}"


@pytest.mark.parametrize("sfx,signed", [("UInt8Euclidian", False), ("Int8Mips", True)])
def test_golden_8bit(engine, sfx, signed):
    assert engine.device_count() > 0, "no CUDA device: the engine has no CPU fallback"
    data, queries, labels = tiny_u8_dataset(signed)
    gold = np.load(os.path.join(GOLDEN, "tiny_u8_ref_outputs.npz"))
    pre = getattr(engine, "PrefilterIndex" + sfx)(data, labels)
    tree = getattr(engine, "RangeFilterTreeIndex" + sfx)(data, labels, TINY_U8["cutoff"], 2, engine.BuildParams(64, 500, 1.0, ""))
    for name, windows, qkw in tiny_u8_cases(labels):
        nq = len(windows)
        qp = engine.QueryParams(10, qkw["beam"], 1.35, 10_000_000, 10_000, qkw["mult"], qkw["max_beam"], None, False)
        ids, d = pre.batch_search(queries[:nq], windows, nq, qp)
        _assert_rows(ids, d, gold[f"{sfx}/{name}/prefilter/ids"], gold[f"{sfx}/{name}/prefilter/dists"], f"{sfx}/{name}/prefilter")
        for m in ("fenwick", "optimized_postfilter", "three_split"):
            ids, d = tree.batch_search(queries[:nq], windows, nq, m, qp)
            _assert_rows(ids, d, gold[f"{sfx}/{name}/pretree_{m}/ids"], gold[f"{sfx}/{name}/pretree_{m}/dists"], f"{sfx}/{name}/pretree_{m}")


@pytest.mark.parametrize("sfx,signed,metric", [("UInt8Euclidian", False, 0), ("Int8Euclidian", True, 0), ("UInt8Mips", False, 1)])
def test_graph_classes_vs_oracle(engine, tmp_path, sfx, signed, metric):
    data, queries, labels = tiny_u8_dataset(signed)
    fdata, fq = data.astype(np.float32), queries.astype(np.float32)
    wst, sup, flat = (str(tmp_path / k) + "/" for k in ("wst", "super", "flat"))
    bp = lambda path: engine.BuildParams(64, 500, 1.0, path)
    with device_graph_build():
        tree = getattr(engine, "VamanaRangeFilterTreeIndex" + sfx)(data, labels, TINY_U8["cutoff"], 2, bp(wst))
        supt = getattr(engine, "SuperOptimizedPostfilterTreeIndex" + sfx)(data, labels, TINY_U8["cutoff"], 2.0, 0.5, bp(sup))
        flt = getattr(engine, "PostfilterVamanaIndex" + sfx)(data, labels, bp(flat))
    o_tree = Oracle("wst", fdata, labels, wst, metric=metric, dist_mode=1, cutoff=TINY_U8["cutoff"])
    o_sup = Oracle("super", fdata, labels, sup, metric=metric, dist_mode=1, cutoff=TINY_U8["cutoff"])
    o_flat = Oracle("flat", fdata, labels, flat, metric=metric, dist_mode=1)
    for name, windows, qkw in tiny_u8_cases(labels):
        nq = len(windows)
        qp = engine.QueryParams(10, qkw["beam"], 1.35, 10_000_000, 10_000, qkw["mult"], qkw["max_beam"], None, False)
        okw = dict(k=10, beam=qkw["beam"], mult=qkw["mult"], max_beam=qkw["max_beam"])
        for m in ("fenwick", "optimized_postfilter", "three_split"):
            ids, d = tree.batch_search(queries[:nq], windows, nq, m, qp)
            oids, od = o_tree.batch(m, fq[:nq], windows, pad_id=0, **okw)
            assert np.array_equal(d.view(np.uint32), od.view(np.uint32)) and np.array_equal(ids, oids), f"{sfx}/{name}/{m}"
        ids, d = supt.batch_search(queries[:nq], windows, nq, qp)
        oids, od = o_sup.batch("super", fq[:nq], windows, pad_id=0, **okw)
        assert np.array_equal(d.view(np.uint32), od.view(np.uint32)) and np.array_equal(ids, oids), f"{sfx}/{name}/super"
        ids, d = flt.batch_search(queries[:nq], windows, nq, qp)
        oids, od = o_flat.batch("flat", fq[:nq], windows, pad_id=0xFFFFFFFF, **okw)
        assert np.array_equal(d.view(np.uint32), od.view(np.uint32)) and np.array_equal(ids, oids), f"{sfx}/{name}/flat"
