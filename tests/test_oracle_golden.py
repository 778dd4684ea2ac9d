"""CPU: pins the oracle restatement (oracle/wsann_oracle.cpp) against golden vectors produced
by the unmodified reference (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from golden_cases import (TINY, TINY_MIPS, TINY_SUPER, TINY_U8, tiny_cases, tiny_dup_dataset, tiny_dup_windows,
                          tiny_mips_cases, tiny_super_cases, tiny_u8_cases, tiny_u8_dataset)
from oracle_api import Oracle
from rangefilteredann_b200 import synth

PADS = {"prefilter": 0xFFFFFFFF, "flat": 0xFFFFFFFF}


@pytest.fixture(scope="module")
def tiny():
    data, queries, labels = synth.make_dataset(TINY["n"], TINY["d"], TINY["nq"], TINY["seed"])
    gold = np.load(os.path.join(GOLDEN, "tiny_ref_outputs.npz"))
    cache = lambda kind: os.path.join(GOLDEN, "tiny", kind) + "/"
    orc = dict(
        wst=Oracle("wst", data, labels, cache("wst"), cutoff=TINY["cutoff"]),
        sup=Oracle("super", data, labels, cache("super"), cutoff=TINY["cutoff"]),
        flat=Oracle("flat", data, labels, cache("flat")),
        pre=Oracle("prefilter", data, labels, None),
    )
    return dict(queries=queries, labels=labels, gold=gold, orc=orc, cases=tiny_cases(labels))


def _oracle_for(t, method):
    return {"prefilter": t["orc"]["pre"], "super": t["orc"]["sup"], "flat": t["orc"]["flat"]}.get(method, t["orc"]["wst"])


@pytest.mark.parametrize("method", ["prefilter", "fenwick", "optimized_postfilter", "three_split", "super", "flat"])
def test_oracle_matches_reference_bit_for_bit(tiny, method):
    checked = 0
    for name, windows, qkw in tiny["cases"]:
        key = f"{name}/{method}/ids"
        if key not in tiny["gold"]:
            continue
        rids, rd = tiny["gold"][key], tiny["gold"][f"{name}/{method}/dists"]
        q = tiny["queries"][: len(windows)]
        ids, d = _oracle_for(tiny, method).batch(method, q, windows, k=10, beam=qkw["beam"], mult=qkw["mult"],
                                                 max_beam=qkw["max_beam"], ratio=qkw.get("ratio"),
                                                 pad_id=PADS.get(method, 0))
        # distances: the reference-order mode reproduces the AVX summation exactly
        assert np.array_equal(d, rd), f"{name}/{method}: distances differ (max rel {np.max(np.abs(d - rd) / np.maximum(np.abs(rd), 1e-30)):.3g})"
        # ids: identical except inside groups of exactly equal distance (unstable sort in the reference)
        diff = ids != rids
        if diff.any():
            for i, j in zip(*np.nonzero(diff)):
                assert (rd[i] == rd[i, j]).sum() > 1 or rd[i, j] == rd[i, -1], f"{name}/{method} row {i} col {j}"
        checked += 1
    assert checked > 0


@pytest.mark.parametrize("method", ["fenwick", "optimized_postfilter", "three_split"])
def test_oracle_matches_reference_prefilter_nodes(tiny, method):
    """RangeFilterTreeIndexFloatEuclidian — the tree whose buckets are PrefilterIndex sub-indices
    (python_bindings.cpp:119-127) — against tests/golden/tiny_pretree_ref_outputs.npz."""
    data, queries, labels = synth.make_dataset(TINY["n"], TINY["d"], TINY["nq"], TINY["seed"])
    gold = np.load(os.path.join(GOLDEN, "tiny_pretree_ref_outputs.npz"))
    orc = Oracle("pretree", data, labels, None, cutoff=TINY["cutoff"])
    checked = 0
    for name, windows, qkw in tiny["cases"]:
        key = f"{name}/{method}/ids"
        if key not in gold:
            continue
        rids, rd = gold[key], gold[f"{name}/{method}/dists"]
        ids, d = orc.batch(method, queries[: len(windows)], windows, k=10, beam=qkw["beam"], mult=qkw["mult"],
                           max_beam=qkw["max_beam"], ratio=qkw.get("ratio"), pad_id=0)
        assert np.array_equal(d, rd), f"{name}/{method}"
        for i, j in zip(*np.nonzero(ids != rids)):
            assert (rd[i] == rd[i, j]).sum() > 1 or rd[i, j] == rd[i, -1], f"{name}/{method} row {i} col {j}"
        checked += 1
    assert checked > 0


def test_device_order_mode_is_close(tiny):
    """dist_mode 1 (the kernels' summation order) differs from the reference order only in
    the last bits."""
    data, queries, labels = synth.make_dataset(TINY["n"], TINY["d"], TINY["nq"], TINY["seed"])
    o1 = Oracle("prefilter", data, labels, None, dist_mode=1)
    w = synth.make_windows(labels, -3, 32, seed=9)
    i0, d0 = tiny["orc"]["pre"].batch("prefilter", queries[:32], w, pad_id=0xFFFFFFFF)
    i1, d1 = o1.batch("prefilter", queries[:32], w, pad_id=0xFFFFFFFF)
    assert np.allclose(d0, d1, rtol=1e-5)
    assert (i0 == i1).mean() > 0.99


@pytest.mark.parametrize("method", ["prefilter", "fenwick", "optimized_postfilter", "three_split", "super"])
def test_oracle_matches_reference_mips(method):
    data, queries, labels = synth.make_dataset(TINY_MIPS["n"], TINY_MIPS["d"], TINY_MIPS["nq"], TINY_MIPS["seed"], angular=True)
    gold = np.load(os.path.join(GOLDEN, "tiny_mips_ref_outputs.npz"))
    kind = {"prefilter": "prefilter", "super": "super"}.get(method, "wst")
    cache = None if kind == "prefilter" else os.path.join(GOLDEN, "tiny_mips", kind) + "/"
    orc = Oracle(kind, data, labels, cache, metric=1, cutoff=TINY_MIPS["cutoff"])
    for name, windows, qkw in tiny_mips_cases(labels):
        rids, rd = gold[f"{name}/{method}/ids"], gold[f"{name}/{method}/dists"]
        ids, d = orc.batch(method, queries[: len(windows)], windows, beam=qkw["beam"], mult=qkw["mult"],
                           max_beam=qkw["max_beam"], pad_id=PADS.get(method, 0))
        assert np.array_equal(d, rd), f"{name}/{method}"
        diff = ids != rids
        for i, j in zip(*np.nonzero(diff)):
            assert (rd[i] == rd[i, j]).sum() > 1 or rd[i, j] == rd[i, -1]


@pytest.mark.parametrize("sfx,signed,metric", [("UInt8Euclidian", False, 0), ("Int8Mips", True, 1)])
@pytest.mark.parametrize("dist_mode", [0, 1])
def test_oracle_matches_reference_8bit(sfx, signed, metric, dist_mode):
    """The 8-bit classes (python_bindings.cpp:233-236) compute integer distances cast to float
    (euclidian_point.h:44-60, mips_point.h:44-58).  On the widened fp32 data every partial sum is an
    integer below 2^24, so BOTH summation orders of the oracle must reproduce the reference bit for bit."""
    data, queries, labels = tiny_u8_dataset(signed)
    gold = np.load(os.path.join(GOLDEN, "tiny_u8_ref_outputs.npz"))
    fdata, fq = data.astype(np.float32), queries.astype(np.float32)
    pre = Oracle("prefilter", fdata, labels, None, metric=metric, dist_mode=dist_mode)
    tree = Oracle("pretree", fdata, labels, None, metric=metric, dist_mode=dist_mode, cutoff=TINY_U8["cutoff"])
    for name, windows, qkw in tiny_u8_cases(labels):
        nq = len(windows)
        runs = [("prefilter", pre.batch("prefilter", fq[:nq], windows, pad_id=0xFFFFFFFF))]
        runs += [(f"pretree_{m}", tree.batch(m, fq[:nq], windows, pad_id=0)) for m in ("fenwick", "optimized_postfilter", "three_split")]
        for key, (ids, d) in runs:
            rids, rd = gold[f"{sfx}/{name}/{key}/ids"], gold[f"{sfx}/{name}/{key}/dists"]
            assert np.array_equal(d, rd), f"{sfx}/{name}/{key}"
            for i, j in zip(*np.nonzero(ids != rids)):  # integer distances tie often; the reference's sort is unstable
                assert (rd[i] == rd[i, j]).sum() > 1 or rd[i, j] == rd[i, -1], f"{sfx}/{name}/{key} row {i} col {j}"


@pytest.mark.parametrize("dist_mode", [0, 1])
def test_oracle_matches_reference_8bit_graph_tree(dist_mode):
    """VamanaRangeFilterTreeIndexUInt8Euclidian on the reference-built graphs under tests/golden/tiny_u8/wst/:
    the beam search sees integer-valued distances with many exact ties, broken by id in the reference
    (beamSearch.h:59-61) and here alike — ids and distances must match bit for bit in both summation orders."""
    data, queries, labels = tiny_u8_dataset(False)
    gold = np.load(os.path.join(GOLDEN, "tiny_u8_ref_outputs.npz"))
    fdata, fq = data.astype(np.float32), queries.astype(np.float32)
    tree = Oracle("wst", fdata, labels, os.path.join(GOLDEN, "tiny_u8", "wst") + "/", dist_mode=dist_mode,
                  cutoff=TINY_U8["cutoff"])
    for name, windows, qkw in tiny_u8_cases(labels):
        nq = len(windows)
        for m in ("fenwick", "optimized_postfilter", "three_split"):
            ids, d = tree.batch(m, fq[:nq], windows, k=10, beam=qkw["beam"], mult=qkw["mult"], max_beam=qkw["max_beam"], pad_id=0)
            rids, rd = gold[f"UInt8Euclidian/{name}/{m}/ids"], gold[f"UInt8Euclidian/{name}/{m}/dists"]
            assert np.array_equal(d, rd), f"{name}/{m}"
            for i, j in zip(*np.nonzero(ids != rids)):
                assert (rd[i] == rd[i, j]).sum() > 1 or rd[i, j] == rd[i, -1], f"{name}/{m} row {i} col {j}"


@pytest.mark.parametrize("split", [3, 4])
@pytest.mark.parametrize("method", ["fenwick", "optimized_postfilter", "three_split"])
def test_oracle_matches_reference_other_split_factors(tiny, split, method):
    """B-WST with split factor 3 / 4 (range_filter_tree.h:129-189; the driver only uses 2): the tree over
    PrefilterIndex buckets answers every bucket exactly, so result vectors pin the decomposition itself."""
    data, queries, labels = synth.make_dataset(TINY["n"], TINY["d"], TINY["nq"], TINY["seed"])
    gold = np.load(os.path.join(GOLDEN, "tiny_pretree_splits_ref_outputs.npz"))
    orc = Oracle("pretree", data, labels, None, cutoff=300, split=float(split))
    checked = 0
    for name, windows, qkw in tiny["cases"]:
        key = f"b{split}/{name}/{method}"
        if key + "/ids" not in gold:
            continue
        rids, rd = gold[key + "/ids"], gold[key + "/dists"]
        ids, d = orc.batch(method, queries[: len(windows)], windows, k=10, beam=qkw["beam"], mult=qkw["mult"],
                           max_beam=qkw["max_beam"], ratio=qkw.get("ratio"), pad_id=0)
        assert np.array_equal(d, rd), key
        for i, j in zip(*np.nonzero(ids != rids)):
            assert (rd[i] == rd[i, j]).sum() > 1 or rd[i, j] == rd[i, -1], f"{key} row {i} col {j}"
        checked += 1
    assert checked > 0


def test_oracle_matches_reference_super_fractional_split():
    """Super-postfilter tree with split 2.5 / shift 0.4: the reference evaluates bucket sizes in float
    (super_optimized_postfilter_tree.h:145-170); graphs under tests/golden/tiny_super/."""
    c = TINY_SUPER
    data, queries, labels = synth.make_dataset(c["n"], c["d"], c["nq"], c["seed"])
    gold = np.load(os.path.join(GOLDEN, "tiny_super_ref_outputs.npz"))
    orc = Oracle("super", data, labels, os.path.join(GOLDEN, "tiny_super") + "/", cutoff=c["cutoff"], split=c["split"], shift=c["shift"])
    for name, windows, qkw in tiny_super_cases(labels):
        ids, d = orc.batch("super", queries[: len(windows)], windows, k=10, beam=qkw["beam"], mult=qkw["mult"],
                           max_beam=qkw["max_beam"], pad_id=0)
        rids, rd = gold[f"{name}/super/ids"], gold[f"{name}/super/dists"]
        assert np.array_equal(d, rd), name
        for i, j in zip(*np.nonzero(ids != rids)):
            assert (rd[i] == rd[i, j]).sum() > 1 or rd[i, j] == rd[i, -1], f"{name} row {i} col {j}"


def dup_window_sizes(labels, w):
    """Points PrefilterIndex::query_knn scans: [lb(lo), lb(hi)) with both searches capped at n-1
    (prefiltering.h:159-184)."""
    sl = np.sort(labels)
    n = len(sl)
    a = np.minimum(np.searchsorted(sl, w[:, 0], side="left"), n - 1)
    b = np.minimum(np.searchsorted(sl, w[:, 1], side="left"), n - 1)
    return np.maximum(b - a, 0)


@pytest.mark.parametrize("k", [1, 5])
def test_oracle_matches_reference_duplicate_labels(k):
    """~12 points per label value, windows that end exactly on label values: lo is inclusive, hi exclusive,
    and a window reaching past the largest label loses the last sorted point.  Rows whose
    window holds fewer than k points are undefined in the reference (it reads past its frontier) and skipped."""
    data, queries, labels = tiny_dup_dataset()
    w = tiny_dup_windows()
    gold = np.load(os.path.join(GOLDEN, "tiny_dup_ref_outputs.npz"))
    assert np.array_equal(gold["windows"], w)
    sizes = dup_window_sizes(labels, w)
    defined = sizes >= k
    # a window reaching past the largest label drops the LAST sorted point; with several points on the largest
    # value, which one is last depends on the reference's unstable sort: not comparable row by row
    defined &= w[:, 1] <= labels.max()
    assert defined.sum() >= len(w) - 3 and not defined[3] and not defined[1]
    orc = Oracle("prefilter", data, labels, None)
    ids, d = orc.batch("prefilter", queries, w, k=k, pad_id=0xFFFFFFFF)
    rids, rd = gold[f"k{k}/ids"], gold[f"k{k}/dists"]
    assert np.array_equal(d[defined], rd[defined])
    for i, j in zip(*np.nonzero(ids != rids)):
        if defined[i]:
            assert (rd[i] == rd[i, j]).sum() > 1 or rd[i, j] == rd[i, -1], f"row {i} col {j}"
    # the semantics, stated independently: every returned point has lo <= label < hi
    lab = labels[ids[defined].astype(np.int64)]
    assert (lab >= w[defined, 0:1]).all() and (lab < w[defined, 1:2]).all()
