#!/usr/bin/env python
"""Generates tests/golden/* from the UNMODIFIED reference (oracle/_ref), in the build
container where /root/reference exists.  Committed outputs:

  tiny/{wst,super,flat}/*.bin   the reference builder's graph files for the `tiny` dataset
                                (3000 x 16, seed 7; rangefilteredann_b200.synth.make_dataset)
  tiny_ref_outputs.npz          reference batch_search ids/dists for every query method on
                                fixed windows (incl. the edge cases of SURVEY.md App. D)
  tiny_pretree_ref_outputs.npz  the same windows through RangeFilterTreeIndexFloatEuclidian (the tree over
                                PrefilterIndex sub-indices, python_bindings.cpp:119-127); `--only pretree`
                                regenerates just this file
  tiny_pretree_splits_ref_outputs.npz   RangeFilterTreeIndexFloatEuclidian with split factors 3 and 4, cutoff 300
                                (`--only splits`)
  tiny_dup_ref_outputs.npz      PrefilterIndex on duplicate labels, windows ending on label values (`--only dup`)
  tiny_super/*.bin, tiny_super_ref_outputs.npz   super-postfilter tree with split 2.5 / shift 0.4 (`--only super`)
  tiny_u8/wst/*.bin, tiny_u8_ref_outputs.npz   the UInt8Euclidian (prefilter, prefilter-bucket tree, Vamana-bucket
                                tree) and Int8Mips (prefilter, prefilter-bucket tree) classes on quantised data
                                (1200 x 64; `--only u8`)

Run: python tests/golden/make_golden.py      (after `make -C oracle ref`)
"""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["PARLAY_NUM_THREADS"] = "8"

from conftest import _load_ext, find_ext  # noqa: E402
from rangefilteredann_b200 import synth  # noqa: E402
from golden_cases import (tiny_cases, tiny_dup_dataset, tiny_dup_windows, tiny_mips_cases, tiny_super_cases,  # noqa: E402
                          tiny_u8_cases, tiny_u8_dataset,
                          TINY, TINY_MIPS, TINY_SUPER, TINY_U8)


def pretree(ref):
    data, queries, labels = synth.make_dataset(TINY["n"], TINY["d"], TINY["nq"], TINY["seed"])
    tree = ref.RangeFilterTreeIndexFloatEuclidian(data, labels, TINY["cutoff"], 2, ref.BuildParams(64, 500, 1.0, ""))
    out = {}
    for name, windows, qkw in tiny_cases(labels):
        nq = len(windows)
        qp = ref.QueryParams(10, qkw["beam"], 1.35, 10_000_000, 10_000, qkw["mult"], qkw["max_beam"], qkw.get("ratio"), False)
        out[f"{name}/windows"] = windows
        for method in ("fenwick", "optimized_postfilter", "three_split"):
            if method == "fenwick" and qkw.get("skip_fenwick"):
                continue
            ids, d = tree.batch_search(queries[:nq], windows, nq, method, qp)
            out[f"{name}/{method}/ids"], out[f"{name}/{method}/dists"] = ids, d
    np.savez_compressed(os.path.join(HERE, "tiny_pretree_ref_outputs.npz"), **out)
    print("wrote", len(out), "pretree arrays")


def pretree_splits(ref):
    """Split factors 3 and 4 (the driver only uses 2): the tree over PrefilterIndex buckets needs no
    graphs, so the whole window -> bucket decomposition of range_filter_tree.h (fenwick cover, optimized
    descent, three_split) is pinned for B != 2 by result vectors alone.  A case on which the reference
    throws (its out-of-range .at(), SURVEY.md §A-12, aborts the whole batch) is recorded as such."""
    data, queries, labels = synth.make_dataset(TINY["n"], TINY["d"], TINY["nq"], TINY["seed"])
    out = {}
    for split in (3, 4):
        tree = ref.RangeFilterTreeIndexFloatEuclidian(data, labels, 300, split, ref.BuildParams(64, 500, 1.0, ""))
        for name, windows, qkw in tiny_cases(labels):
            nq = len(windows)
            qp = ref.QueryParams(10, qkw["beam"], 1.35, 10_000_000, 10_000, qkw["mult"], qkw["max_beam"], qkw.get("ratio"), False)
            for method in ("fenwick", "optimized_postfilter", "three_split"):
                key = f"b{split}/{name}/{method}"
                try:
                    ids, d = tree.batch_search(queries[:nq], windows, nq, method, qp)
                except Exception as e:  # noqa: BLE001
                    print("reference throws on", key, "->", str(e)[:80])
                    out[key + "/throws"] = np.array([1])
                    continue
                out[key + "/ids"], out[key + "/dists"] = ids, d
    np.savez_compressed(os.path.join(HERE, "tiny_pretree_splits_ref_outputs.npz"), **out)
    print("wrote", len(out), "split-factor arrays")


def super_fractional(ref):
    """SuperOptimizedPostfilterTreeIndexFloatEuclidian with split 2.5 / shift 0.4 on 800 points: graphs under
    tiny_super/ and result vectors."""
    c = TINY_SUPER
    data, queries, labels = synth.make_dataset(c["n"], c["d"], c["nq"], c["seed"])
    sdir = os.path.join(HERE, "tiny_super")
    os.makedirs(sdir, exist_ok=True)
    sup = ref.SuperOptimizedPostfilterTreeIndexFloatEuclidian(data, labels, c["cutoff"], c["split"], c["shift"],
                                                              ref.BuildParams(64, 500, 1.0, sdir + "/"))
    out = {}
    for name, windows, qkw in tiny_super_cases(labels):
        nq = len(windows)
        qp = ref.QueryParams(10, qkw["beam"], 1.35, 10_000_000, 10_000, qkw["mult"], qkw["max_beam"], None, False)
        ids, d = sup.batch_search(queries[:nq], windows, nq, qp)
        out[f"{name}/super/ids"], out[f"{name}/super/dists"] = ids, d
    np.savez_compressed(os.path.join(HERE, "tiny_super_ref_outputs.npz"), **out)
    print("wrote", len(out), "super arrays,", len(os.listdir(sdir)), "graphs")


def duplicate_labels(ref):
    """PrefilterIndexFloatEuclidian on data with ~12 points per label value, windows ending on label values."""
    data, queries, labels = tiny_dup_dataset()
    w = tiny_dup_windows()
    pre = ref.PrefilterIndexFloatEuclidian(data, labels)
    out = {"windows": w}
    for k in (1, 5):
        qp = ref.QueryParams(k, 10, 1.35, 10_000_000, 10_000, 1, 10000, None, False)
        ids, d = pre.batch_search(queries, w, len(w), qp)
        out[f"k{k}/ids"], out[f"k{k}/dists"] = ids, d
    np.savez_compressed(os.path.join(HERE, "tiny_dup_ref_outputs.npz"), **out)
    print("wrote", len(out), "duplicate-label arrays")


def eight_bit(ref):
    """UInt8Euclidian: PrefilterIndex, the tree over prefilter buckets, and the tree over Vamana buckets
    (3 reference-built graphs under tiny_u8/wst/); Int8Mips: PrefilterIndex and the tree over prefilter
    buckets."""
    out = {}
    udir = os.path.join(HERE, "tiny_u8", "wst")
    os.makedirs(udir, exist_ok=True)
    for sfx, signed in (("UInt8Euclidian", False), ("Int8Mips", True)):
        data, queries, labels = tiny_u8_dataset(signed)
        pre = getattr(ref, "PrefilterIndex" + sfx)(data, labels)
        ptree = getattr(ref, "RangeFilterTreeIndex" + sfx)(data, labels, TINY_U8["cutoff"], 2, ref.BuildParams(64, 500, 1.0, ""))
        vtree = None
        if not signed:
            vtree = getattr(ref, "VamanaRangeFilterTreeIndex" + sfx)(data, labels, TINY_U8["cutoff"], 2,
                                                                     ref.BuildParams(64, 500, 1.0, udir + "/"))
        for name, windows, qkw in tiny_u8_cases(labels):
            nq = len(windows)
            qp = ref.QueryParams(10, qkw["beam"], 1.35, 10_000_000, 10_000, qkw["mult"], qkw["max_beam"], None, False)
            out[f"{sfx}/{name}/windows"] = windows
            ids, d = pre.batch_search(queries[:nq], windows, nq, qp)
            out[f"{sfx}/{name}/prefilter/ids"], out[f"{sfx}/{name}/prefilter/dists"] = ids, d
            for method in ("fenwick", "optimized_postfilter", "three_split"):
                ids, d = ptree.batch_search(queries[:nq], windows, nq, method, qp)
                out[f"{sfx}/{name}/pretree_{method}/ids"], out[f"{sfx}/{name}/pretree_{method}/dists"] = ids, d
                if vtree is not None:
                    ids, d = vtree.batch_search(queries[:nq], windows, nq, method, qp)
                    out[f"{sfx}/{name}/{method}/ids"], out[f"{sfx}/{name}/{method}/dists"] = ids, d
    np.savez_compressed(os.path.join(HERE, "tiny_u8_ref_outputs.npz"), **out)
    print("wrote", len(out), "8-bit arrays")


def main():
    ref = _load_ext(find_ext(os.path.join(ROOT, "oracle", "_ref")))
    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None
    if only in (None, "pretree"):
        pretree(ref)
    if only in (None, "u8"):
        eight_bit(ref)
    if only in (None, "splits"):
        pretree_splits(ref)
    if only in (None, "super"):
        super_fractional(ref)
    if only in (None, "dup"):
        duplicate_labels(ref)
    if only is not None:
        return
    data, queries, labels = synth.make_dataset(TINY["n"], TINY["d"], TINY["nq"], TINY["seed"])
    out = {}
    tiny_dir = os.path.join(HERE, "tiny")
    for kind in ("wst", "super", "flat"):
        os.makedirs(os.path.join(tiny_dir, kind), exist_ok=True)
    bp = lambda kind: ref.BuildParams(64, 500, 1.0, os.path.join(tiny_dir, kind) + "/")
    tree = ref.VamanaRangeFilterTreeIndexFloatEuclidian(data, labels, TINY["cutoff"], 2, bp("wst"))
    sup = ref.SuperOptimizedPostfilterTreeIndexFloatEuclidian(data, labels, TINY["cutoff"], 2.0, 0.5, bp("super"))
    flat = ref.PostfilterVamanaIndexFloatEuclidian(data, labels, bp("flat"))
    pre = ref.PrefilterIndexFloatEuclidian(data, labels)
    for name, windows, qkw in tiny_cases(labels):
        nq = len(windows)
        q = queries[:nq]
        qp = ref.QueryParams(10, qkw["beam"], 1.35, 10_000_000, 10_000, qkw["mult"], qkw["max_beam"], qkw.get("ratio"), False)
        out[f"{name}/windows"] = windows
        if qkw.get("prefilter", True):
            ids, d = pre.batch_search(q, windows, nq, qp)
            out[f"{name}/prefilter/ids"], out[f"{name}/prefilter/dists"] = ids, d
        for method in ("fenwick", "optimized_postfilter", "three_split"):
            if method == "fenwick" and qkw.get("skip_fenwick"):
                continue
            ids, d = tree.batch_search(q, windows, nq, method, qp)
            out[f"{name}/{method}/ids"], out[f"{name}/{method}/dists"] = ids, d
        ids, d = sup.batch_search(q, windows, nq, qp)
        out[f"{name}/super/ids"], out[f"{name}/super/dists"] = ids, d
        ids, d = flat.batch_search(q, windows, nq, qp)
        out[f"{name}/flat/ids"], out[f"{name}/flat/dists"] = ids, d
    np.savez_compressed(os.path.join(HERE, "tiny_ref_outputs.npz"), **out)

    # ---- MIPS variant (Mips_Point, padded rows)
    out = {}
    data, queries, labels = synth.make_dataset(TINY_MIPS["n"], TINY_MIPS["d"], TINY_MIPS["nq"], TINY_MIPS["seed"], angular=True)
    mdir = os.path.join(HERE, "tiny_mips")
    for kind in ("wst", "super"):
        os.makedirs(os.path.join(mdir, kind), exist_ok=True)
    bpm = lambda kind: ref.BuildParams(64, 500, 1.0, os.path.join(mdir, kind) + "/")
    tree = ref.VamanaRangeFilterTreeIndexFloatMips(data, labels, TINY_MIPS["cutoff"], 2, bpm("wst"))
    sup = ref.SuperOptimizedPostfilterTreeIndexFloatMips(data, labels, TINY_MIPS["cutoff"], 2.0, 0.5, bpm("super"))
    pre = ref.PrefilterIndexFloatMips(data, labels)
    for name, windows, qkw in tiny_mips_cases(labels):
        nq = len(windows)
        q = queries[:nq]
        qp = ref.QueryParams(10, qkw["beam"], 1.35, 10_000_000, 10_000, qkw["mult"], qkw["max_beam"], None, False)
        out[f"{name}/windows"] = windows
        ids, d = pre.batch_search(q, windows, nq, qp)
        out[f"{name}/prefilter/ids"], out[f"{name}/prefilter/dists"] = ids, d
        for method in ("fenwick", "optimized_postfilter", "three_split"):
            ids, d = tree.batch_search(q, windows, nq, method, qp)
            out[f"{name}/{method}/ids"], out[f"{name}/{method}/dists"] = ids, d
        ids, d = sup.batch_search(q, windows, nq, qp)
        out[f"{name}/super/ids"], out[f"{name}/super/dists"] = ids, d
    np.savez_compressed(os.path.join(HERE, "tiny_mips_ref_outputs.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
