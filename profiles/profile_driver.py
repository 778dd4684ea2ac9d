#!/usr/bin/env python
"""Small fixed workload for ncu captures: one config, one method/beam/fraction, few batches.

  python profiles/profile_driver.py --config c2 --method optimized_postfilter --beam 80 --power 0 --reps 3
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rangefilteredann_b200 import capi, load_engine, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c1")
    ap.add_argument("--method", default="fenwick")
    ap.add_argument("--beam", type=int, default=10)
    ap.add_argument("--mult", type=int, default=1)
    ap.add_argument("--power", type=int, default=-3)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--nq", type=int, default=10000)
    ap.add_argument("--opt", action="append", default=[], help="name=value engine option")
    ap.add_argument("--index", default="tree", choices=["tree", "prefilter"],
                    help="prefilter: PrefilterIndex only (no graphs to load; for captures of the scan / tensor-core kernels)")
    a = ap.parse_args()
    cfg = bench.CONFIGS[a.config]
    eng = load_engine()
    data, queries, labels = synth.make_dataset(cfg["n"], cfg["d"], a.nq, cfg["seed"])
    cdir = bench.cache_dir(a.config)
    os.makedirs(cdir, exist_ok=True)
    if a.index == "tree":
        bench.validate_cache(cdir, data, labels)
    t0 = time.time()
    if a.index == "prefilter":
        tree = eng.PrefilterIndexFloatEuclidian(data, labels)
    else:
        tree = eng.VamanaRangeFilterTreeIndexFloatEuclidian(data, labels, cfg["cutoff"], 2, eng.BuildParams(64, 500, 1.0, cdir))
    print(f"index ready in {time.time() - t0:.1f}s", flush=True)
    h = capi.Handle.borrow(tree)
    for o in a.opt:
        name, v = o.split("=")
        h.set_option(name, int(v))
    w = synth.make_windows(labels, a.power, a.nq, seed=1000 + a.power)
    dq, dw = h.dalloc(queries.nbytes), h.dalloc(w.nbytes)
    di, dd = h.dalloc(a.nq * 10 * 4), h.dalloc(a.nq * 10 * 4)
    h.h2d(dq, queries); h.h2d(dw, w)
    qp = capi.query_params(k=10, beam=a.beam, final_multiply=a.mult)
    h.reset_stats()
    for r in range(a.reps):
        h.timer_start()
        if a.method == "prefilter":
            h.prefilter_batch(dq, dw, a.nq, 10, di, dd, device_ptrs=True)
        else:
            h.tree_batch(a.method, dq, dw, a.nq, qp, di, dd, device_ptrs=True)
        ms = h.timer_stop()
        print(f"rep {r}: {ms:.3f} ms  {a.nq / ms * 1000:.0f} qps", flush=True)
    st = h.stats()
    print({k: v // a.reps for k, v in st.items()})
    try:
        print({k: round(v["ms"] / a.reps, 3) for k, v in h.kernel_times().items()})
    except Exception:
        pass


if __name__ == "__main__":
    main()
