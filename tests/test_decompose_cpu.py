"""CPU: the window -> task decomposition the device runs (ws_decompose.h, evaluated on the
host through ws_debug_decompose_host) against the oracle's trace of the reference control
flow, on the tiny tree and on BASELINE config-2 geometry (1M points, 11 rows)."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import GOLDEN
from golden_cases import TINY, tiny_cases
from oracle_api import Oracle
from rangefilteredann_b200 import capi, synth


def build_host_tree(labels_sorted, cutoff, split=2, super_params=None, prefilter_nodes=False):
    """Host-only (device -1) index with the reference's B-WST / super geometry.  prefilter_nodes: the
    B-WST buckets are PrefilterIndex sub-indices (ws_index_set_wst without node handles)."""
    n = len(labels_sorted)
    L = capi.lib()
    h = C.c_void_p()
    capi.check(L.ws_index_create(-1, 0, n, 8, None, capi.ptr(labels_sorted), None, 1, C.byref(h)))
    idx = capi.Handle(h.value, True)

    def add(start, count):
        node = C.c_int32()
        capi.check(L.ws_index_add_graph(idx.raw, start, count, 64, None, None, C.byref(node)))
        return node.value

    offs = [[0, n]]
    while offs[-1][1] > cutoff:
        last = offs[-1]
        nxt = []
        for b in range(len(last) - 1):
            ls, size = last[b], last[b + 1] - last[b]
            large = (size + split - 1) // split
            small = large - 1
            nl = size - small * split
            for i in range(split):
                nxt.append(ls + i * large if i < nl else ls + nl * large + (i - nl) * small)
        nxt.append(n)
        offs.append(nxt)
    row_nb = np.array([len(r) - 1 for r in offs], np.uint32)
    off_flat = np.array([x for r in offs for x in r], np.uint64)
    if prefilter_nodes:
        capi.check(L.ws_index_set_wst(idx.raw, len(offs), split, cutoff, capi.ptr(row_nb), capi.ptr(off_flat), None))
    else:
        nodes = np.array([add(r[b], r[b + 1] - r[b]) for r in offs for b in range(len(r) - 1)], np.int32)
        capi.check(L.ws_index_set_wst(idx.raw, len(offs), split, cutoff, capi.ptr(row_nb), capi.ptr(off_flat), capi.ptr(nodes)))
    if super_params:
        sf, sh = super_params
        sizes, shifts, nbs, snodes = [n], [0], [1], [add(0, n)]
        while sizes[-1] > cutoff:
            bsize = int(np.float32(np.float32(np.float32(sizes[-1]) + np.float32(sf)) - np.float32(1)) / np.float32(sf))
            bshift = int(np.ceil(np.float32(bsize) * np.float32(sh)))
            sizes.append(bsize); shifts.append(bshift)
            nb = ((n - bsize) + bshift - 1) // bshift + 1
            nbs.append(nb)
            for b in range(nb):
                s = b * bshift
                snodes.append(add(s, min(s + bsize, n) - s))
        a_sizes, a_shifts = np.array(sizes, np.uint64), np.array(shifts, np.uint64)
        a_nbs, a_nodes = np.array(nbs, np.uint32), np.array(snodes, np.int32)  # keep alive across the call
        capi.check(L.ws_index_set_super(idx.raw, len(sizes), cutoff, capi.ptr(a_sizes), capi.ptr(a_shifts),
                                        capi.ptr(a_nbs), capi.ptr(a_nodes)))
    capi.check(L.ws_index_finalize(idx.raw))
    return idx


def host_decompose(idx, method, windows, beam=10, mult=2, ratio=None):
    L = capi.lib()
    m = capi.METHODS.get(method, capi.MODE_PREFILTER)
    cap = C.c_uint32()
    capi.check(L.ws_index_task_capacity(idx.raw, m, C.byref(cap)))
    nq = len(windows)
    out = np.empty((nq, cap.value, 4), np.int64)
    cnt = np.empty(nq, np.uint32)
    qp = capi.query_params(beam=beam, final_multiply=mult, ratio=ratio)
    windows = np.ascontiguousarray(windows, np.float32)
    capi.check(L.ws_debug_decompose_host(idx.raw, m, capi.ptr(windows), nq, C.byref(qp), cap.value, capi.ptr(out), capi.ptr(cnt)))
    res = []
    for i in range(nq):
        t = out[i, : cnt[i]].copy()
        t[:, 0] = np.where(t[:, 0] >= 0, 0, -1)  # node handle -> kind
        t[:, 3] &= 1
        res.append(t)
    return res


def join_chunks(tasks, chunk=8192):
    """The engine cuts a long scan into scan_chunk-row tasks; the oracle records the slice whole."""
    out = []
    for t in tasks.tolist():
        if out and t[0] == -1 and out[-1][0] == -1 and out[-1][2] == t[1] and (out[-1][2] - out[-1][1]) % chunk == 0:
            out[-1][2] = t[2]
        else:
            out.append(list(t))
    return np.array(out, np.int64).reshape(-1, 4)


def check_against_oracle(idx, orc, method, windows, ratio=None, join=False):
    got = host_decompose(idx, method, windows, ratio=ratio)
    for i, w in enumerate(windows):
        exp = orc.decompose(method, float(w[0]), float(w[1]), ratio=ratio)
        if join:
            got[i] = join_chunks(got[i])
        assert np.array_equal(got[i], exp), f"{method} window {w}: engine {got[i].tolist()} vs oracle {exp.tolist()}"


@pytest.mark.parametrize("method", ["fenwick", "optimized_postfilter", "three_split", "super", "prefilter"])
def test_decompose_tiny(method):
    data, queries, labels = synth.make_dataset(TINY["n"], TINY["d"], TINY["nq"], TINY["seed"])
    sl = np.sort(labels)
    idx = build_host_tree(sl, TINY["cutoff"], super_params=(2.0, 0.5))
    kind = {"super": "super", "prefilter": "prefilter"}.get(method, "wst")
    orc = Oracle(kind, data, labels, None, cutoff=TINY["cutoff"])
    for name, windows, qkw in tiny_cases(labels):
        check_against_oracle(idx, orc, method, windows, ratio=qkw.get("ratio") if method != "super" else None)


@pytest.mark.parametrize("method", ["fenwick", "optimized_postfilter", "three_split"])
def test_decompose_prefilter_nodes(method):
    """RangeFilterTreeIndex over PrefilterIndex sub-indices (python_bindings.cpp:119-127): every bucket
    query becomes a scan of [lb(lo), lb(hi)) inside the bucket, with query_knn's r = count-1 rule."""
    data, queries, labels = synth.make_dataset(TINY["n"], TINY["d"], TINY["nq"], TINY["seed"])
    idx = build_host_tree(np.sort(labels), TINY["cutoff"], prefilter_nodes=True)
    orc = Oracle("pretree", data, labels, None, cutoff=TINY["cutoff"])
    for name, windows, qkw in tiny_cases(labels):
        check_against_oracle(idx, orc, method, windows, ratio=qkw.get("ratio"))
    # config-2 geometry: bucket scans longer than scan_chunk are cut into pieces
    n = 1_000_000
    rng = np.random.default_rng(6)
    labels = (rng.permutation(n).astype(np.float64) / n).astype(np.float32)
    idx = build_host_tree(np.sort(labels), 1000, prefilter_nodes=True)
    orc = Oracle("pretree", np.zeros((n, 1), np.float32), labels, None, cutoff=1000)
    for power in (-12, -6, -2, 0):
        check_against_oracle(idx, orc, method, synth.make_windows(labels, power, 24, seed=power + 70), join=True)


@pytest.mark.parametrize("method", ["fenwick", "optimized_postfilter", "three_split", "super"])
def test_decompose_c2_geometry(method):
    """1M unique labels, cutoff 1000, B=2 -> 11 rows / 2047 nodes."""
    n = 1_000_000
    rng = np.random.default_rng(5)
    labels = (rng.permutation(n).astype(np.float64) / n).astype(np.float32)
    sl = np.sort(labels)
    idx = build_host_tree(sl, 1000, super_params=(2.0, 0.5) if method == "super" else None)
    data = np.zeros((n, 1), np.float32)
    orc = Oracle("super" if method == "super" else "wst", data, labels, None, cutoff=1000)
    for power in (-16, -12, -9, -6, -3, -1, 0):
        w = synth.make_windows(labels, power, 40, seed=power + 50)
        check_against_oracle(idx, orc, method, w)
    # adversarial: windows around bucket boundaries and the ends of the label range
    s = sl.astype(np.float64)
    edges = []
    for a, b in [(499_990, 500_010), (0, 977), (0, 1), (999_000, 1_000_000), (999_999, 1_000_000), (250_000, 750_000),
                 (1, 999_999), (976, 1954), (500_000, 500_977), (123_456, 123_456 + 1953)]:
        lo = s[a] - 1e-9 if a > 0 else s[0] - 1.0
        hi = (0.5 * (s[b - 1] + s[b])) if b < n else s[-1] + 1.0
        edges.append((lo, hi))
    check_against_oracle(idx, orc, method, np.array(edges, np.float32))


@pytest.mark.parametrize("split", [3, 4, 7])
@pytest.mark.parametrize("prefilter_nodes", [False, True])
def test_decompose_other_split_factors(split, prefilter_nodes):
    """Split factors other than the driver's 2: the device decomposition (evaluated on the host) against the
    oracle's trace, tiny tree (cutoff 300) and a 200 000-point tree (cutoff 1000)."""
    data, queries, labels = synth.make_dataset(TINY["n"], TINY["d"], TINY["nq"], TINY["seed"])
    kind = "pretree" if prefilter_nodes else "wst"
    idx = build_host_tree(np.sort(labels), 300, split=split, prefilter_nodes=prefilter_nodes)
    orc = Oracle(kind, data, labels, None, cutoff=300, split=float(split))
    for method in ("fenwick", "optimized_postfilter", "three_split"):
        for name, windows, qkw in tiny_cases(labels):
            check_against_oracle(idx, orc, method, windows, ratio=qkw.get("ratio"))
    n = 200_000
    rng = np.random.default_rng(9)
    labels = (rng.permutation(n).astype(np.float64) / n).astype(np.float32)
    idx = build_host_tree(np.sort(labels), 1000, split=split, prefilter_nodes=prefilter_nodes)
    orc = Oracle(kind, np.zeros((n, 1), np.float32), labels, None, cutoff=1000, split=float(split))
    for method in ("fenwick", "optimized_postfilter", "three_split"):
        for power in (-14, -9, -5, -2, 0):
            check_against_oracle(idx, orc, method, synth.make_windows(labels, power, 24, seed=power + 90), join=True)
