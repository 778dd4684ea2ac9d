#!/usr/bin/env python
"""2+ GPU check of the in-process multi-GPU paths (run on a multi-GPU box: `python tests/multigpu_group.py`):

  replicas       one batch_search call cut into one slice per GPU          -> rows identical to 1 GPU
  label shards   every GPU answers the whole batch on its label range, rows gathered by peer loads over NVLink
                 (exchange 0) or by ncclAllGather inside the library (exchange 1), merged on the device
                                                                           -> prefilter rows identical to 1 GPU,
                                                                              graph methods recall >= 1 GPU - 0.005
Prints per-phase device times of the label-sharded call (ws_group_info) for both exchanges."""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from rangefilteredann_b200 import capi, load_engine, synth  # noqa: E402


def same(a, b):
    return np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))


def main():
    eng = load_engine()
    ndev = eng.device_count()
    if ndev < 2:
        print("MULTIGPU_GROUP_SKIP: needs >= 2 GPUs")
        return 0
    G = int(os.environ.get("WSANN_TEST_GPUS", ndev))
    devs = ",".join(str(i) for i in range(G))
    n, d, nq = int(os.environ.get("WSANN_TEST_N", 200_000)), 96, 4000  # Deep-shaped rows (96-d L2), scaled down
    data, queries, labels = synth.make_dataset(n, d, nq, seed=5)
    cache = os.path.join(tempfile.gettempdir(), "wsann_group_test")
    qp = lambda beam, mult=1: eng.QueryParams(10, beam, 1.35, 10_000_000, 10_000, mult, 10000, None, False)  # noqa: E731
    ok = True

    os.environ["WSANN_DEVICES"] = "0"
    t0 = time.time()
    single = eng.VamanaRangeFilterTreeIndexFloatEuclidian(data, labels, 1000, 2, eng.BuildParams(64, 500, 1.0, cache + "/single/"))
    pre1 = eng.PrefilterIndexFloatEuclidian(data, labels)
    print(f"[group] single-GPU tree ready in {time.time() - t0:.1f}s", flush=True)

    # ---- replicas
    os.environ["WSANN_DEVICES"] = devs
    t0 = time.time()
    rep = eng.VamanaRangeFilterTreeIndexFloatEuclidian(data, labels, 1000, 2, eng.BuildParams(64, 500, 1.0, cache + "/single/"))
    print(f"[group] {G} replicas ready in {time.time() - t0:.1f}s (graphs loaded once, arena cloned device to device)", flush=True)
    for power in (-8, -3, 0):
        w = synth.make_windows(labels, power, nq, seed=77 + power)
        for method, beam in (("fenwick", 20), ("optimized_postfilter", 40)):
            a = single.batch_search(queries, w, nq, method, qp(beam))
            b = rep.batch_search(queries, w, nq, method, qp(beam))
            good = same(a, b)
            ok &= good
            print(f"[group] replicas x{G} 2^{power} {method}: rows identical {good}", flush=True)
    del rep

    # ---- label shards
    os.environ["WSANN_SHARD_MODE"] = "label"
    t0 = time.time()
    sh = eng.VamanaRangeFilterTreeIndexFloatEuclidian(data, labels, 1000, 2, eng.BuildParams(64, 500, 1.0, cache + "/shards/"))
    shpre = eng.PrefilterIndexFloatEuclidian(data, labels)
    print(f"[group] {G} label shards ready in {time.time() - t0:.1f}s (one host thread and one GPU per shard)", flush=True)
    g, gpre = capi.Group.borrow(sh), capi.Group.borrow(shpre)
    print(f"[group] peer access {g.info()['peer_access']}, NCCL {capi.nccl_version()}", flush=True)
    for exchange in (0, 1):
        g.set_option("exchange", exchange)
        gpre.set_option("exchange", exchange)
        for power in (-8, -3, 0):
            w = synth.make_windows(labels, power, nq, seed=77 + power)
            gt = synth.ground_truth(data, queries[:500], labels, w[:500])
            a = pre1.batch_search(queries, w, nq, qp(10))
            b = shpre.batch_search(queries, w, nq, qp(10))
            good = same(a, b)
            ok &= good
            info = gpre.info()
            print(f"[group] label shards x{G} exchange={info['exchange']} 2^{power} prefilter: rows identical {good}; "
                  f"search {info['search_ms']:.3f} ms, exchange+merge {info['exchange_merge_ms']:.3f} ms", flush=True)
            for method, beam in (("fenwick", 20), ("optimized_postfilter", 40)):
                r1 = synth.recall_std(single.batch_search(queries, w, nq, method, qp(beam))[0][:500], gt)
                ids, dd = sh.batch_search(queries, w, nq, method, qp(beam))
                # timed repetition (buffers allocated, communicator up)
                ids, dd = sh.batch_search(queries, w, nq, method, qp(beam))
                info = g.info()
                r2 = synth.recall_std(ids[:500], gt)
                good = r2 >= r1 - 0.005 and bool((np.diff(dd, axis=1) >= 0).all())
                ok &= good
                print(f"[group] label shards x{G} exchange={info['exchange']} 2^{power} {method}: recall {r2:.4f} (1 GPU {r1:.4f}) ok {good}; "
                      f"search {info['search_ms']:.3f} ms, exchange+merge {info['exchange_merge_ms']:.3f} ms", flush=True)
    print("MULTIGPU_GROUP_OK" if ok else "MULTIGPU_GROUP_FAIL", flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
