"""Mirror of the reference's experiments/wrapper.py helpers that experiments/run_our_method.py
calls (wrapper.py:196-355): `(metric, dtype) -> class` dispatch and `build_query_params`,
bound to this engine's `window_ann` module.  Same names, arguments, defaults and errors, so
the reference driver runs unchanged with `import rangefilteredann_b200.wrapper as wp`.
Only dtype "float" is reachable in the reference as well."""
from __future__ import annotations

from . import load_engine

_eng = load_engine()
QueryParams = _eng.QueryParams
BuildParams = _eng.BuildParams


def _pick(base: str, metric: str, dtype: str):
    # wrapper.py:244-330 — same accepted spellings and the same exceptions
    if dtype != "float":
        raise Exception("Invalid data type " + dtype)
    if metric == "Euclidian":
        return getattr(_eng, base + "FloatEuclidian")
    if metric == "mips":
        return getattr(_eng, base + "FloatMips")
    raise Exception("Invalid metric " + metric)


def range_filter_tree_index_constructor(metric, dtype):  # wrapper.py:219-239 — the one helper that spells the 8-bit
    # classes the way the module registers them ("UInt8" / "Int8"), so all three dtypes are reachable through it
    names = {"float": "Float", "uint8": "UInt8", "int8": "Int8"}
    if metric not in ("Euclidian", "mips"):
        raise Exception("Invalid metric " + metric)
    if dtype not in names:
        raise Exception("Invalid data type " + dtype)
    return getattr(_eng, "RangeFilterTreeIndex" + names[dtype] + ("Euclidian" if metric == "Euclidian" else "Mips"))


def prefilter_index_constructor(metric, dtype):          # wrapper.py:242-266
    return _pick("PrefilterIndex", metric, dtype)


def postfilter_vamana_constructor(metric, dtype):        # wrapper.py:244-268
    return _pick("PostfilterVamanaIndex", metric, dtype)


def vamana_range_filter_tree_constructor(metric, dtype):  # wrapper.py:270-298
    return _pick("VamanaRangeFilterTreeIndex", metric, dtype)


def super_optimized_postfilter_tree_constructor(metric, dtype):  # wrapper.py:300-331
    return _pick("SuperOptimizedPostfilterTreeIndex", metric, dtype)


def build_query_params(k, beam_size, cut=1.35, limit=10_000_000, degree_limit=10_000, final_beam_multiply=1,
                       postfiltering_max_beam=10000, min_query_to_bucket_ratio=None, verbose=False):
    """wrapper.py:334-355 (same defaults)."""
    return QueryParams(k, beam_size, cut, limit, degree_limit, final_beam_multiply, postfiltering_max_beam,
                       min_query_to_bucket_ratio, verbose)
