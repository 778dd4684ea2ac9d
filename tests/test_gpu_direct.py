"""One-launch prefilter kernel (ws_prefilter_direct_kernel, K1d) vs the task path (K3 decomposition ->
K1 scan -> K4 merge): same bounds (prefiltering.h:159-184 with the r = n-1 rule), same fp32 arithmetic
and folding, so ids AND distances must be bit-identical on every input — small and large windows,
windows shorter than k, empty / inverted / out-of-range windows, padded rows, MIPS, k from 1 to 128,
host buffers (auto routing) and device pointers."""
import numpy as np
import pytest

from rangefilteredann_b200 import capi, synth
from test_gpu_gemm import make, mixed_windows

pytestmark = pytest.mark.gpu


def run(h, queries, windows, k, direct):
    h.set_option("gemm_prefilter", 0)
    h.set_option("prefilter_direct", direct)
    nq = len(windows)
    ids = np.empty((nq, k), np.uint32)
    dists = np.empty((nq, k), np.float32)
    h.prefilter_batch(np.ascontiguousarray(queries[:nq]), np.ascontiguousarray(windows, dtype=np.float32), nq, k, ids, dists)
    return ids, dists


def assert_same(h, queries, windows, k):
    h.set_option("profile_kernels", 1)
    h.kernel_times(reset=True)
    l0 = h.launches()
    di, dd = run(h, queries, windows, k, 1)
    assert h.launches() - l0 == 1, "the direct path is one kernel launch"
    kt = h.kernel_times(reset=True)
    assert set(kt) == {"scan"} and kt["scan"]["launches"] == 1, kt
    ti, td = run(h, queries, windows, k, 0)
    kt = h.kernel_times(reset=True)
    assert kt.get("decompose", {}).get("launches", 0) == 1, kt
    assert np.array_equal(dd.view(np.uint32), td.view(np.uint32)), \
        f"distances differ in rows {np.nonzero((dd.view(np.uint32) != td.view(np.uint32)).any(axis=1))[0][:8]}"
    assert np.array_equal(di, ti), f"ids differ in rows {np.nonzero((di != ti).any(axis=1))[0][:8]}"
    return di, dd


@pytest.mark.parametrize("k", [1, 10, 33, 128])
def test_mixed_windows_all_k(engine, k):
    data, queries, labels, idx, h = make(engine, 50000, 128, 1100, seed=3)
    w = mixed_windows(labels, 1000, seed=11)
    ids, dists = assert_same(h, queries, w, k)
    fmax = np.finfo(np.float32).max
    assert (ids[-6] == 0xFFFFFFFF).all() and (dists[-6] == fmax).all()   # empty window -> pads
    assert (ids[-4] == 0xFFFFFFFF).all() and (ids[-3] == 0xFFFFFFFF).all() and (ids[-2] == 0xFFFFFFFF).all()
    m = min(k, 3)
    assert (ids[-5][:m] != 0xFFFFFFFF).all() and (ids[-5][3:] == 0xFFFFFFFF).all()  # 3 in-window points


@pytest.mark.parametrize("d,angular", [(100, True), (96, False), (32, False), (24, True)])
def test_other_shapes(engine, d, angular):
    data, queries, labels, idx, h = make(engine, 20000, d, 700, angular=angular, seed=5)
    assert_same(h, queries, mixed_windows(labels, 640, seed=9), 10)


def test_small_batches_and_device_pointers(engine):
    data, queries, labels, idx, h = make(engine, 30000, 128, 300, seed=8)
    for nq in (1, 3, 31, 300):
        w = synth.make_windows(labels, -7, nq, seed=40 + nq)
        assert_same(h, queries, w, 10)
    # device-resident batch: the option is taken as it is
    nq = 300
    w = synth.make_windows(labels, -9, nq, seed=77)
    dq, dw = h.dalloc(queries[:nq].nbytes), h.dalloc(w.nbytes)
    di, dd = h.dalloc(nq * 10 * 4), h.dalloc(nq * 10 * 4)
    h.h2d(dq, np.ascontiguousarray(queries[:nq])); h.h2d(dw, w)
    out = {}
    for direct in (1, 0):
        h.set_option("gemm_prefilter", 0)
        h.set_option("prefilter_direct", direct)
        l0 = h.launches()
        h.prefilter_batch(dq, dw, nq, 10, di, dd, device_ptrs=True)
        h.sync()
        assert (h.launches() - l0 == 1) == (direct == 1)
        ids, dists = np.empty((nq, 10), np.uint32), np.empty((nq, 10), np.float32)
        h.d2h(ids, di); h.d2h(dists, dd)
        out[direct] = (ids, dists)
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1].view(np.uint32), out[1][1].view(np.uint32))


def test_auto_routing_with_host_buffers(engine):
    """auto (2): a batch of small windows in host memory takes the one-launch kernel, a batch of large
    windows does not (it goes to the tensor-core sweep or the task path)."""
    data, queries, labels, idx, h = make(engine, 60000, 128, 512, seed=2)
    h.set_option("prefilter_direct", 2)
    h.set_option("gemm_prefilter", 2)
    for power, expect_direct in ((-8, True), (-1, False)):
        w = synth.make_windows(labels, power, 512, seed=50 + power)
        ids, dists = np.empty((512, 10), np.uint32), np.empty((512, 10), np.float32)
        l0 = h.launches()
        h.prefilter_batch(queries, w, 512, 10, ids, dists)
        assert (h.launches() - l0 == 1) == expect_direct, (power, h.launches() - l0)
