#!/usr/bin/env python
"""Doubling-tail probe: optimized postfiltering (range_filter_tree.h:403-471 + postfilter_vamana.h:141-188) over a
range of filter fractions and start beams on one config, device-resident batches, CUDA-event timed, with the
per-tier kernel times and the device counters.  Used to size the beam tiers (profiles/r02_tail.md).

  python profiles/tail_probe.py --config c2 --powers -11,-8,-6,-5,-4,-3 --beams 10,80
"""
import argparse
import json
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
    ap.add_argument("--config", default="c2")
    ap.add_argument("--powers", default="-11,-10,-9,-8,-7,-6,-5,-4,-3,-2")
    ap.add_argument("--beams", default="10,80")
    ap.add_argument("--methods", default="optimized_postfilter")
    ap.add_argument("--nq", type=int, default=10000)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--opt", action="append", default=[], help="name=value engine option")
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    cfg = bench.CONFIGS[a.config]
    eng = load_engine()
    data, queries, labels = synth.make_dataset(cfg["n"], cfg["d"], a.nq, cfg["seed"])
    cdir = bench.cache_dir(a.config)
    os.makedirs(cdir, exist_ok=True)
    bench.validate_cache(cdir, data, labels)
    t0 = time.time()
    tree = eng.VamanaRangeFilterTreeIndexFloatEuclidian(data, labels, cfg["cutoff"], 2, eng.BuildParams(64, 500, 1.0, cdir))
    print(f"index ready in {time.time() - t0:.1f}s", flush=True)
    h = capi.Handle.borrow(tree)
    for o in a.opt:
        name, v = o.split("=")
        h.set_option(name, int(v))
    dq = h.dalloc(queries.nbytes)
    di, dd = h.dalloc(a.nq * 10 * 4), h.dalloc(a.nq * 10 * 4)
    h.h2d(dq, queries)
    out = []
    for p in [int(x) for x in a.powers.split(",")]:
        w = synth.make_windows(labels, p, a.nq, seed=1000 + p)
        dw = h.dalloc(w.nbytes)
        h.h2d(dw, w)
        gt = None
        for method in a.methods.split(","):
            for beam in [int(x) for x in a.beams.split(",")]:
                qp = capi.query_params(k=10, beam=beam, final_multiply=1)
                h.tree_batch(method, dq, dw, a.nq, qp, di, dd, device_ptrs=True)  # warm
                h.sync()
                h.set_option("profile_kernels", 1)
                h.reset_stats()
                h.kernel_times(reset=True)
                best = 1e30
                for _ in range(a.reps):
                    h.timer_start()
                    h.tree_batch(method, dq, dw, a.nq, qp, di, dd, device_ptrs=True)
                    best = min(best, h.timer_stop())
                kt = {k: round(v["ms"] / a.reps, 3) for k, v in h.kernel_times().items()}
                st = {k: v // a.reps for k, v in h.stats().items()}
                h.set_option("profile_kernels", 0)
                # un-profiled timing (events around every launch serialise nothing, but keep a clean number too)
                clean = 1e30
                for _ in range(a.reps):
                    h.timer_start()
                    h.tree_batch(method, dq, dw, a.nq, qp, di, dd, device_ptrs=True)
                    clean = min(clean, h.timer_stop())
                ids = np.empty((a.nq, 10), np.uint32)
                h.d2h(ids, di)
                if gt is None:
                    gt = synth.ground_truth(data, queries[:500], labels, w[:500])
                rec = bench.recall_at_k(ids[:500], gt)
                row = dict(power=p, method=method, beam=beam, ms=round(clean, 3), ms_profiled=round(best, 3), qps=round(a.nq / clean * 1000),
                           recall500=round(rec, 4), kernels=kt, searches=st["graph_searches"], escalated=st["escalated_tasks"],
                           visited=st["visited"], dist_cmps=st["dist_cmps"])
                out.append(row)
                print(json.dumps(row), flush=True)
        h.dfree(dw)
    if a.json:
        json.dump(out, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
