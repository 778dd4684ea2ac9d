"""Tensor-core prefilter (csrc/ws_gemm.cu) vs the streaming-scan prefilter: the fp16 sweep only
SELECTS candidates (with a proven error slack) and the survivors are re-ranked with the scan
kernel's fp32 arithmetic, so ids AND distances must be bit-identical to the scan path (which is
itself bit-identical to the device-order oracle, test_gpu_oracle.py) on every input: ragged
windows, windows shorter than k, empty windows, padded rows (d=100 -> 112), MIPS, multi-slice
batches, k from 1 to 16."""
import numpy as np
import pytest

from oracle_api import Oracle
from rangefilteredann_b200 import capi, synth

pytestmark = pytest.mark.gpu


def make(engine, n, d, nq, angular=False, seed=0):
    data, queries, labels = synth.make_dataset(n, d, nq, seed, angular)
    sfx = "FloatMips" if angular else "FloatEuclidian"
    idx = getattr(engine, "PrefilterIndex" + sfx)(data, labels)
    return data, queries, labels, idx, capi.Handle.borrow(idx)


def run(h, queries, windows, k, mode):
    h.set_option("gemm_prefilter", mode)
    h.set_option("prefilter_direct", 0)  # the baseline is the task path (K3 -> K1 -> K4)
    nq = len(windows)
    ids = np.empty((nq, k), np.uint32)
    dists = np.empty((nq, k), np.float32)
    h.prefilter_batch(np.ascontiguousarray(queries[:nq]), np.ascontiguousarray(windows, dtype=np.float32), nq, k, ids, dists)
    return ids, dists


def assert_same(h, queries, windows, k):
    h.set_option("profile_kernels", 1)
    h.kernel_times(reset=True)
    gi, gd = run(h, queries, windows, k, 1)
    kt = h.kernel_times(reset=True)
    assert kt.get("gemm_sweep", {}).get("launches", 0) >= 1, f"tensor-core path did not run: {kt}"
    si, sd = run(h, queries, windows, k, 0)
    kt = h.kernel_times(reset=True)
    assert "gemm_sweep" not in kt and kt.get("scan", {}).get("launches", 0) >= 1
    assert np.array_equal(gd.view(np.uint32), sd.view(np.uint32)), \
        f"distances differ in rows {np.nonzero((gd.view(np.uint32) != sd.view(np.uint32)).any(axis=1))[0][:8]}"
    assert np.array_equal(gi, si), f"ids differ in rows {np.nonzero((gi != si).any(axis=1))[0][:8]}"
    return gi, gd


def mixed_windows(labels, nq, seed):
    """all fractions mixed in one batch + degenerate windows"""
    parts = []
    powers = [0, -1, -3, -5, -7, -9, -11, -13]
    per = nq // len(powers)
    for i, p in enumerate(powers):
        parts.append(synth.make_windows(labels, p, per, seed=seed + i))
    w = np.concatenate(parts).astype(np.float32)
    rng = np.random.default_rng(seed)
    rng.shuffle(w, axis=0)
    s = np.sort(labels)
    extra = np.array([[s[10], s[10]],                 # empty (lo == hi)
                      [s[5] - 1e-6, s[8]],            # 3 points < k
                      [s[-1] + 1.0, s[-1] + 2.0],     # beyond every label
                      [s[0] - 2.0, s[0] - 1.0],       # before every label
                      [s[100], s[50]],                # inverted
                      [s[0] - 1.0, s[-1] + 1.0]], np.float32)
    return np.concatenate([w, extra])


@pytest.mark.parametrize("power", [0, -2, -5, -9])
def test_l2_128_fixed_fraction(engine, power):
    data, queries, labels, idx, h = make(engine, 60000, 128, 1500)
    w = synth.make_windows(labels, power, 1500, seed=100 + power)
    assert_same(h, queries, w, 10)


def test_mixed_windows_and_degenerate(engine):
    data, queries, labels, idx, h = make(engine, 50000, 128, 2100, seed=3)
    w = mixed_windows(labels, 2000, seed=11)
    ids, dists = assert_same(h, queries, w, 10)
    assert (ids[-6] == 0xFFFFFFFF).all() and (dists[-6] == np.finfo(np.float32).max).all()   # empty window -> pads
    assert (ids[-5][3:] == 0xFFFFFFFF).all() and (ids[-5][:3] != 0xFFFFFFFF).all()            # 3 in-window points


@pytest.mark.parametrize("k", [1, 5, 16])
def test_k_sweep(engine, k):
    data, queries, labels, idx, h = make(engine, 40000, 128, 1000, seed=5)
    w = mixed_windows(labels, 960, seed=21)
    assert_same(h, queries, w, k)


def test_k_above_tensor_path_uses_scan(engine):
    data, queries, labels, idx, h = make(engine, 20000, 128, 300, seed=5)
    w = synth.make_windows(labels, -2, 300, seed=2)
    h.set_option("profile_kernels", 1)
    h.kernel_times(reset=True)
    run(h, queries, w, 17, 1)
    kt = h.kernel_times(reset=True)
    assert "gemm_sweep" not in kt and kt["scan"]["launches"] >= 1


def test_mips_padded_rows(engine):
    """d=100 -> 112-float rows: the last 32-column block is half out of bounds (TMA zero fill)"""
    data, queries, labels, idx, h = make(engine, 40000, 100, 1200, angular=True, seed=9)
    w = mixed_windows(labels, 1160, seed=31)
    assert_same(h, queries, w, 10)


def test_l2_96(engine):
    data, queries, labels, idx, h = make(engine, 40000, 96, 1200, seed=4)
    w = mixed_windows(labels, 1160, seed=41)
    assert_same(h, queries, w, 10)


def test_multi_slice_batch(engine):
    """more queries than one plan sorts (16384): slices are planned one after the other"""
    data, queries, labels, idx, h = make(engine, 30000, 32, 20000, seed=6)
    w = np.concatenate([synth.make_windows(labels, p, 5000, seed=150 + p) for p in (-1, -4, -6, -10)]).astype(np.float32)
    assert_same(h, queries, w, 10)


def test_against_device_order_oracle(engine):
    data, queries, labels, idx, h = make(engine, 30000, 128, 600, seed=8)
    w = mixed_windows(labels, 560, seed=61)
    gi, gd = run(h, queries, w, 10, 1)
    orc = Oracle("prefilter", data, labels, None, metric=0, dist_mode=1)
    oi, od = orc.batch("prefilter", queries[:len(w)], w, k=10, beam=10, mult=1, max_beam=10000, ratio=None, pad_id=0xFFFFFFFF)
    full = (oi != 0xFFFFFFFF).all(axis=1)   # the reference reads past its frontier when the window holds < k points
    assert np.array_equal(gd[full].view(np.uint32), od[full].view(np.uint32))
    assert np.array_equal(gi[full], oi[full])


def test_auto_mode_through_the_pybind_surface(engine):
    """PrefilterIndex.batch_search (host buffers): large windows take the tensor-core path by
    themselves, small windows keep the streaming scan; rows are the same either way."""
    data, queries, labels, idx, h = make(engine, 60000, 128, 1024, seed=2)
    qp = engine.QueryParams(10, 10, 1.35, 10_000_000, 10_000, 1, 10000, None, False)
    h.set_option("profile_kernels", 1)
    for power, expect_gemm in ((-1, True), (-12, False)):
        w = synth.make_windows(labels, power, 1024, seed=170 + power)
        h.set_option("gemm_prefilter", 2)
        h.kernel_times(reset=True)
        ids, dists = idx.batch_search(queries, [tuple(x) for x in w], 1024, qp)
        kt = h.kernel_times(reset=True)
        assert ("gemm_sweep" in kt) == expect_gemm, kt
        h.set_option("gemm_prefilter", 0)
        ids0, dists0 = idx.batch_search(queries, [tuple(x) for x in w], 1024, qp)
        assert np.array_equal(ids, ids0) and np.array_equal(dists.view(np.uint32), dists0.view(np.uint32))
    h.set_option("gemm_prefilter", 2)


@pytest.mark.parametrize("d,angular", [(256, False), (192, False), (320, True), (512, True)])
def test_wide_rows(engine, d, angular):
    """rows above 128 columns: 3-8 blocks of 64 fp16 per tile (d=512: two pipeline stages per tile, the
    query operand takes 256 TMEM columns and leaves two accumulator stages; d=320: 5 blocks padded to 6)"""
    data, queries, labels, idx, h = make(engine, 12000, d, 600, angular=angular, seed=12)
    w = mixed_windows(labels, 560, seed=71)
    assert_same(h, queries, w, 10)


def test_zero_tiny_and_huge_queries(engine):
    """the error slack has an absolute term (norm-table and distance rounding), so a zero or tiny query still gets
    the exact rows; a query too large to scale into fp16 falls back to the exact scan inside the re-rank kernel"""
    data, queries, labels, idx, h = make(engine, 30000, 128, 512, seed=13)
    q = queries.copy()
    q[0] = 0.0
    q[1] *= 1e-6
    q[2] *= 1e-20
    q[3] *= 1e6
    q[4] *= 1e25      # beyond the per-query scale range: exact-scan fallback
    q[5, 7] = 3e38    # |q|^2 overflows fp32
    q[6, :] = 0.0
    q[6, 3] = 1e-30   # one subnormal-scale component
    w = synth.make_windows(labels, -2, 512, seed=81)
    assert_same(h, q, w, 10)
    w = mixed_windows(labels, 480, seed=82)
    assert_same(h, q, w, 5)


@pytest.mark.parametrize("scale", [1e-12, 3.7e9])
def test_badly_scaled_arena(engine, scale):
    """component magnitudes far from 1: the fp16 mirror is built with a power-of-two scale"""
    data, queries, labels = synth.make_dataset(20000, 64, 400, 14)
    data = (data * scale).astype(np.float32)
    queries = (queries * scale).astype(np.float32)
    idx = engine.PrefilterIndexFloatEuclidian(data, labels)
    h = capi.Handle.borrow(idx)
    w = mixed_windows(labels, 360, seed=91)
    assert_same(h, queries, w, 10)


def test_non_finite_arena_uses_scan(engine):
    """an arena with an inf component cannot be mirrored in fp16: the batch is answered by the scan kernels"""
    data, queries, labels = synth.make_dataset(5000, 32, 300, 15)
    data[17, 3] = np.inf
    idx = engine.PrefilterIndexFloatEuclidian(data, labels)
    h = capi.Handle.borrow(idx)
    w = synth.make_windows(labels, -2, 300, seed=3)
    h.set_option("profile_kernels", 1)
    h.kernel_times(reset=True)
    run(h, queries, w, 10, 1)
    kt = h.kernel_times(reset=True)
    assert "gemm_sweep" not in kt and kt["scan"]["launches"] >= 1
