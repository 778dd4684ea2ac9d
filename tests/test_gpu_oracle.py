"""GPU parity against the oracle restatement run in the kernels' own summation order
(dist_mode 1): ids AND fp32 distances must be bit-identical — the traversal, the label
predicate, the doubling loop, the decomposition, merge, decode and padding are all integer /
index work once distances agree.  Also covers the golden MIPS vectors, the larger beam
tiers, the device-pointer entry points, k sweeps and empty batches."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import DATA_CACHE, GOLDEN, REF_CACHE, ROOT
from golden_cases import TINY, TINY_MIPS, tiny_cases, tiny_mips_cases
from oracle_api import Oracle
from rangefilteredann_b200 import capi, synth
from test_gpu_golden import rows_equal_up_to_ties

pytestmark = pytest.mark.gpu
PADS = {"prefilter": 0xFFFFFFFF, "flat": 0xFFFFFFFF}


class Pair:
    """Engine indices + device-order oracles over one dataset and graph cache."""

    def __init__(self, engine, cfg, cache_root, angular=False, kinds=("wst", "super", "flat", "prefilter")):
        self.engine = engine
        self.data, self.queries, self.labels = synth.make_dataset(cfg["n"], cfg["d"], cfg["nq"], cfg["seed"], angular)
        sfx = "FloatMips" if angular else "FloatEuclidian"
        bp = lambda kind: engine.BuildParams(64, 500, 1.0, os.path.join(cache_root, kind) + "/")
        metric = 1 if angular else 0
        self.eng, self.orc = {}, {}
        if "wst" in kinds:
            self.eng["wst"] = getattr(engine, "VamanaRangeFilterTreeIndex" + sfx)(self.data, self.labels, cfg["cutoff"], 2, bp("wst"))
            self.orc["wst"] = Oracle("wst", self.data, self.labels, os.path.join(cache_root, "wst") + "/", metric=metric, dist_mode=1, cutoff=cfg["cutoff"])
        if "super" in kinds:
            self.eng["super"] = getattr(engine, "SuperOptimizedPostfilterTreeIndex" + sfx)(self.data, self.labels, cfg["cutoff"], 2.0, 0.5, bp("super"))
            self.orc["super"] = Oracle("super", self.data, self.labels, os.path.join(cache_root, "super") + "/", metric=metric, dist_mode=1, cutoff=cfg["cutoff"])
        if "flat" in kinds:
            self.eng["flat"] = getattr(engine, "PostfilterVamanaIndex" + sfx)(self.data, self.labels, bp("flat"))
            self.orc["flat"] = Oracle("flat", self.data, self.labels, os.path.join(cache_root, "flat") + "/", metric=metric, dist_mode=1)
        if "prefilter" in kinds:
            self.eng["prefilter"] = getattr(engine, "PrefilterIndex" + sfx)(self.data, self.labels)
            self.orc["prefilter"] = Oracle("prefilter", self.data, self.labels, None, metric=metric, dist_mode=1)

    def kind_of(self, method):
        return {"prefilter": "prefilter", "super": "super", "flat": "flat"}.get(method, "wst")

    def run_engine(self, method, q, w, k=10, beam=10, mult=1, max_beam=10000, ratio=None):
        qp = self.engine.QueryParams(k, beam, 1.35, 10_000_000, 10_000, mult, max_beam, ratio, False)
        idx = self.eng[self.kind_of(method)]
        if method in ("prefilter", "super", "flat"):
            return idx.batch_search(q, w, len(w), qp)
        return idx.batch_search(q, w, len(w), method, qp)

    def run_oracle(self, method, q, w, k=10, beam=10, mult=1, max_beam=10000, ratio=None):
        return self.orc[self.kind_of(method)].batch(method, q, w, k=k, beam=beam, mult=mult, max_beam=max_beam,
                                                    ratio=ratio, pad_id=PADS.get(method, 0))

    def assert_identical(self, method, q, w, **kw):
        ids, d = self.run_engine(method, q, w, **kw)
        oids, od = self.run_oracle(method, q, w, **kw)
        assert np.array_equal(d.view(np.uint32), od.view(np.uint32)), \
            f"{method} {kw}: distances not bit-identical in {np.nonzero((d != od).any(axis=1))[0][:6]}"
        assert np.array_equal(ids, oids), f"{method} {kw}: ids differ in rows {np.nonzero((ids != oids).any(axis=1))[0][:6]}"


@pytest.fixture(scope="module")
def tiny(engine):
    assert engine.device_count() > 0, "no CUDA device: the engine has no CPU fallback"
    return Pair(engine, TINY, os.path.join(GOLDEN, "tiny"))


@pytest.fixture(scope="module")
def tiny_mips(engine):
    return Pair(engine, TINY_MIPS, os.path.join(GOLDEN, "tiny_mips"), angular=True, kinds=("wst", "super", "prefilter"))


METHODS = ["prefilter", "fenwick", "optimized_postfilter", "three_split", "super", "flat"]


@pytest.mark.parametrize("method", METHODS)
def test_bit_exact_vs_device_order_oracle(tiny, method):
    for name, windows, qkw in tiny_cases(tiny.labels):
        if method == "prefilter" and not qkw.get("prefilter", True):
            continue
        q = tiny.queries[: len(windows)]
        tiny.assert_identical(method, q, windows, beam=qkw["beam"], mult=qkw["mult"], max_beam=qkw["max_beam"],
                              ratio=qkw.get("ratio") if method in ("optimized_postfilter", "three_split") else None)


@pytest.mark.parametrize("method", ["prefilter", "fenwick", "optimized_postfilter", "three_split", "super"])
def test_mips_golden_and_oracle(tiny_mips, method):
    gold = np.load(os.path.join(GOLDEN, "tiny_mips_ref_outputs.npz"))
    for name, windows, qkw in tiny_mips_cases(tiny_mips.labels):
        q = tiny_mips.queries[: len(windows)]
        kw = dict(beam=qkw["beam"], mult=qkw["mult"], max_beam=qkw["max_beam"])
        tiny_mips.assert_identical(method, q, windows, **kw)
        ids, d = tiny_mips.run_engine(method, q, windows, **kw)
        ok = rows_equal_up_to_ties(ids, d, gold[f"{name}/{method}/ids"], gold[f"{name}/{method}/dists"])
        assert ok.mean() >= (1.0 if method == "prefilter" else 0.95), f"{name}/{method}: {ok.sum()}/{len(ok)}"


@pytest.mark.parametrize("beam,mult,max_beam", [(64, 1, 10000), (65, 1, 10000), (100, 4, 10000), (300, 2, 10000),
                                                (200, 1, 10000), (120, 2, 10000), (256, 1, 10000),
                                                (1100, 1, 10000), (10, 32, 10000), (700, 8, 4000), (10, 1, 12288)])
def test_beam_tiers(tiny, beam, mult, max_beam):
    """Every tier (warp 64/128/256, CTA 1024, global-bitmap 12288), fresh and escalated."""
    w = synth.make_windows(tiny.labels, -4, 24, seed=beam)
    q = tiny.queries[:24]
    for method in ("optimized_postfilter", "flat", "super", "fenwick"):
        tiny.assert_identical(method, q, w, beam=beam, mult=mult, max_beam=max_beam)


@pytest.mark.parametrize("k", [1, 3, 10, 37, 100, 150])
def test_k_sweep(tiny, k):
    w = synth.make_windows(tiny.labels, -2, 16, seed=k)
    q = tiny.queries[:16]
    for method in ("prefilter", "fenwick", "optimized_postfilter", "super", "flat"):
        tiny.assert_identical(method, q, w, k=k, beam=max(k, 20))


def test_counters_match_oracle(tiny):
    """`visited` (beamSearch.h:117) is identical; dist_cmps can only be <= the reference's
    lossy-hash count since the engine's visited set forgets less."""
    w = synth.make_windows(tiny.labels, -3, 32, seed=77)
    q = tiny.queries[:32]
    h = capi.Handle.borrow(tiny.eng["wst"])
    h.reset_stats()
    tiny.run_engine("optimized_postfilter", q, w, beam=20, mult=2)
    st = h.stats()
    orc0 = Oracle("wst", tiny.data, tiny.labels, os.path.join(GOLDEN, "tiny", "wst") + "/", dist_mode=1, cutoff=TINY["cutoff"])
    _, _, ost = orc0.batch("optimized_postfilter", q, w, beam=20, mult=2, stats=True)
    assert st["visited"] == ost["visited"]
    assert st["graph_searches"] == ost["graph_searches"]
    assert 0 < st["dist_cmps"] <= ost["dist_cmps"]


def test_device_pointer_path_and_empty_batch(tiny):
    w = synth.make_windows(tiny.labels, -3, 32, seed=5)
    q = np.ascontiguousarray(tiny.queries[:32])
    ids_h, d_h = tiny.run_engine("fenwick", q, w, beam=20)
    h = capi.Handle.borrow(tiny.eng["wst"])
    dq, dw = h.dalloc(q.nbytes), h.dalloc(w.nbytes)
    di, dd = h.dalloc(32 * 10 * 4), h.dalloc(32 * 10 * 4)
    h.h2d(dq, q); h.h2d(dw, w)
    qp = capi.query_params(k=10, beam=20)
    h.timer_start()
    h.tree_batch("fenwick", dq, dw, 32, qp, di, dd, device_ptrs=True)
    ms = h.timer_stop()
    h.sync()
    ids = np.empty((32, 10), np.uint32); d = np.empty((32, 10), np.float32)
    h.d2h(ids, di); h.d2h(d, dd)
    assert ms > 0 and np.array_equal(ids, ids_h) and np.array_equal(d, d_h)
    for p in (dq, dw, di, dd):
        h.dfree(p)
    ids0, d0 = tiny.run_engine("fenwick", q[:0], w[:0], beam=20)
    assert ids0.shape == (0, 10) and d0.shape == (0, 10)
    assert h.launches() > 0


def test_cta_tiers_match_warp_tiers(tiny):
    """Beams <= 256 normally run on the warp-per-task kernels; with them disabled the
    CTA-per-task kernel must give the same bits."""
    h = capi.Handle.borrow(tiny.eng["wst"])
    w = synth.make_windows(tiny.labels, -3, 48, seed=31)
    q = tiny.queries[:48]
    try:
        h.set_option("warp_tiers", 0)
        for beam, mult in ((10, 1), (40, 2), (100, 4)):
            for method in ("fenwick", "optimized_postfilter", "three_split"):
                tiny.assert_identical(method, q, w, beam=beam, mult=mult)
    finally:
        h.set_option("warp_tiers", 1)
    try:
        h.set_option("hash16", 0)  # 32-bit visited table in the warp tiers
        for method in ("fenwick", "optimized_postfilter"):
            tiny.assert_identical(method, q, w, beam=40, mult=2)
    finally:
        h.set_option("hash16", 1)
    try:
        h.set_option("fuse_scan", 1)
        for method in ("fenwick", "three_split"):
            tiny.assert_identical(method, q, w, beam=10, mult=1)
    finally:
        h.set_option("fuse_scan", 0)
    try:
        h.set_option("fuse_scan", 0)  # separate scan launch instead of draining scans in the beam launch
        for method in ("fenwick", "three_split", "optimized_postfilter"):
            tiny.assert_identical(method, q, w, beam=10, mult=1)
    finally:
        h.set_option("fuse_scan", 0)
    try:
        h.set_option("warp_scan", 0)  # CTA-per-task scan kernel instead of the warp-per-task one
        h.set_option("fuse_scan", 0)
        for method in ("fenwick", "three_split"):
            tiny.assert_identical(method, q, w, beam=10, mult=1)
        hp = capi.Handle.borrow(tiny.eng["prefilter"])
        hp.set_option("warp_scan", 0)
        tiny.assert_identical("prefilter", q, w)
        hp.set_option("warp_scan", 1)
    finally:
        h.set_option("warp_scan", 1)
        h.set_option("fuse_scan", 0)
    try:
        h.set_option("warp_hash", 256)  # a saturated visited table may only cost recomputation
        tiny.assert_identical("optimized_postfilter", q, w, beam=60, mult=2)
    finally:
        h.set_option("warp_hash", 2048)


def test_rejects_unsupported(tiny, engine):
    w = synth.make_windows(tiny.labels, -2, 4, seed=1)
    with pytest.raises(RuntimeError, match="postfiltering_max_beam"):
        tiny.run_engine("optimized_postfilter", tiny.queries[:4], w, beam=10, max_beam=50000)
    with pytest.raises(RuntimeError):
        tiny.run_engine("fenwick", tiny.queries[:4], w, k=5000)
    assert os.environ.get("WSANN_GRAPH_BUILD") == "0"  # tests/conftest.py: no implicit device builds
    with pytest.raises(RuntimeError, match="graph cache miss"):
        engine.PostfilterVamanaIndexFloatEuclidian(tiny.data, tiny.labels, engine.BuildParams(64, 500, 1.0, "/nonexistent/"))


def test_small_config_bit_exact(engine, tmp_path_factory):
    """20 000 x 32, 6-row tree + super tree + flat index on REFERENCE-BUILT graphs (ref_cache/small, written by
    oracle/build_ref_cache.py with the unmodified reference builder; rebuilt on the host — under a minute — when the
    directory did not travel with the snapshot)."""
    cfg = dict(n=20000, d=32, nq=256, seed=3, cutoff=1000)
    root = os.path.join(REF_CACHE, "small")
    if not all(os.path.isdir(os.path.join(root, kind)) and os.listdir(os.path.join(root, kind)) for kind in ("wst", "super", "flat")):
        import subprocess
        import sys
        out = str(tmp_path_factory.mktemp("ref_cache_small"))
        subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "build_ref_cache.py"), "small", "--kinds", "wst,super,flat", "--out", out],
                       check=True, stdout=subprocess.DEVNULL)
        root = os.path.join(out, "small")
    p = Pair(engine, cfg, root)
    for power in (-10, -6, -3, -1, 0):
        w = synth.make_windows(p.labels, power, 128, seed=300 + power)
        q = p.queries[:128]
        for method in METHODS:
            if method == "prefilter" and int(cfg["n"] * 2.0 ** power) < 10:
                continue
            p.assert_identical(method, q, w, beam=20, mult=2)
