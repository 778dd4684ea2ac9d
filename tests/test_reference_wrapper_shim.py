"""The reference's OWN experiments/wrapper.py (`from window_ann import *`) imported over this engine's
`window_ann.py` shim: every helper experiments/run_our_method.py calls (run_our_method.py:238-509) must resolve to
this engine's classes, build the same QueryParams / BuildParams, and raise the reference's exceptions.

Needs /root/reference (present in the build container only, never on the GPU box): a CPU test, skipped elsewhere.
The end-to-end behaviour of the classes themselves is what the -m gpu tests cover."""
import importlib.util
import os
import sys

import pytest

REF_WRAPPER = "/root/reference/experiments/wrapper.py"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "rangefilteredann_b200")

pytestmark = pytest.mark.skipif(not os.path.exists(REF_WRAPPER), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref_wrapper():
    saved_path, saved_mod = list(sys.path), sys.modules.pop("window_ann", None)
    sys.path.insert(0, PKG)  # `import window_ann` now finds rangefilteredann_b200/window_ann.py
    try:
        spec = importlib.util.spec_from_file_location("_reference_wrapper_under_test", REF_WRAPPER)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        yield mod
    finally:
        sys.path[:] = saved_path
        sys.modules.pop("window_ann", None)
        if saved_mod is not None:
            sys.modules["window_ann"] = saved_mod


def test_reference_wrapper_binds_to_this_engine(ref_wrapper, engine):
    wp = ref_wrapper
    for metric, sfx in (("Euclidian", "FloatEuclidian"), ("mips", "FloatMips")):
        assert wp.prefilter_index_constructor(metric, "float") is getattr(engine, "PrefilterIndex" + sfx)
        assert wp.postfilter_vamana_constructor(metric, "float") is getattr(engine, "PostfilterVamanaIndex" + sfx)
        assert wp.vamana_range_filter_tree_constructor(metric, "float") is getattr(engine, "VamanaRangeFilterTreeIndex" + sfx)
        assert wp.super_optimized_postfilter_tree_constructor(metric, "float") is getattr(engine, "SuperOptimizedPostfilterTreeIndex" + sfx)
        assert wp.range_filter_tree_index_constructor(metric, "float") is getattr(engine, "RangeFilterTreeIndex" + sfx)
    # the one helper that spells the 8-bit classes as the module registers them (wrapper.py:219-239)
    assert wp.range_filter_tree_index_constructor("Euclidian", "uint8") is engine.RangeFilterTreeIndexUInt8Euclidian
    assert wp.range_filter_tree_index_constructor("mips", "int8") is engine.RangeFilterTreeIndexInt8Mips
    assert wp.BuildParams is engine.BuildParams and wp.QueryParams is engine.QueryParams


def test_reference_wrapper_errors_and_params(ref_wrapper, engine):
    wp = ref_wrapper
    with pytest.raises(Exception, match="Invalid metric"):
        wp.prefilter_index_constructor("cosine", "float")
    with pytest.raises(Exception, match="Invalid data type"):
        wp.vamana_range_filter_tree_constructor("Euclidian", "double")
    # the "Uint8" spelling wrapper.py:244-330 asks for exists in neither module: same NameError
    with pytest.raises(NameError):
        wp.prefilter_index_constructor("Euclidian", "uint8")
    qp = wp.build_query_params(k=10, beam_size=40, final_beam_multiply=2, verbose=False)   # run_our_method.py:354-360
    assert isinstance(qp, engine.QueryParams)
    bp = wp.BuildParams(64, 500, 1.0, "index_cache/x/")                                   # run_our_method.py:309
    assert isinstance(bp, engine.BuildParams)


def test_this_repos_wrapper_mirror_has_the_same_helpers(ref_wrapper):
    from rangefilteredann_b200 import wrapper as mine
    for name in ("range_filter_tree_index_constructor", "prefilter_index_constructor", "postfilter_vamana_constructor",
                 "vamana_range_filter_tree_constructor", "super_optimized_postfilter_tree_constructor", "build_query_params"):
        assert callable(getattr(mine, name)) and callable(getattr(ref_wrapper, name))
    for metric in ("Euclidian", "mips"):
        assert mine.vamana_range_filter_tree_constructor(metric, "float") is ref_wrapper.vamana_range_filter_tree_constructor(metric, "float")
