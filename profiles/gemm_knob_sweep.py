#!/usr/bin/env python
"""Sweep of the tensor-core prefilter's planning knobs (work items per plan, chunk of the label axis one
item sweeps, smallest item) on config c2, PrefilterIndex only, device-resident batches, CUDA-event timed.

  python profiles/gemm_knob_sweep.py > gpurun_out/gemm_knobs.txt
"""
import itertools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rangefilteredann_b200 import capi, load_engine, synth  # noqa: E402


def main():
    cfg = bench.CONFIGS["c2"]
    nq = cfg["nq"]
    eng = load_engine()
    data, queries, labels = synth.make_dataset(cfg["n"], cfg["d"], nq, cfg["seed"])
    idx = eng.PrefilterIndexFloatEuclidian(data, labels)
    h = capi.Handle.borrow(idx)
    dq = h.dalloc(queries.nbytes)
    di, dd = h.dalloc(nq * 10 * 4), h.dalloc(nq * 10 * 4)
    h.h2d(dq, queries)
    h.set_option("gemm_prefilter", 1)
    powers = [int(p) for p in os.environ.get("POWERS", "-10,-8,-6,-4,-2,0").split(",")]
    items = [int(x) for x in os.environ.get("ITEMS", "0,444,592,888,1480,2960").split(",")]
    chunks = [int(x) for x in os.environ.get("CHUNKS", "32,16,8,4").split(",")]
    dyns = [int(x) for x in os.environ.get("DYN", "1").split(",")]
    base_ids = {}
    for p in powers:
        w = synth.make_windows(labels, p, nq, seed=1000 + p)
        dw = h.dalloc(w.nbytes)
        h.h2d(dw, w)
        rows = []
        for it, ch, dyn in itertools.product(items, chunks, dyns):
            h.set_option("gemm_items", it)
            h.set_option("gemm_chunk_mb", ch)
            h.set_option("gemm_dynamic", dyn)
            best = 1e9
            for _ in range(4):
                h.timer_start()
                h.prefilter_batch(dq, dw, nq, 10, di, dd, device_ptrs=True)
                best = min(best, h.timer_stop())
            ids = np.empty((nq, 10), np.uint32)
            h.d2h(ids, di)
            if p not in base_ids:
                base_ids[p] = ids
            same = bool(np.array_equal(ids, base_ids[p]))
            h.set_option("profile_kernels", 1)
            h.kernel_times(reset=True)
            h.prefilter_batch(dq, dw, nq, 10, di, dd, device_ptrs=True)
            kt = h.kernel_times(reset=True)
            h.set_option("profile_kernels", 0)
            rows.append((best, it, ch, dyn, kt.get("gemm_sweep", {}).get("ms", 0.0), same))
        rows.sort()
        print(f"== 2^{p}")
        for best, it, ch, dyn, sweep, same in rows[:int(os.environ.get("TOP", "8"))]:
            print(f"   {best:.3f} ms  items={it} chunk_mb={ch} dynamic={dyn} sweep={sweep:.3f} ms rows_identical={same}")
        sys.stdout.flush()


if __name__ == "__main__":
    main()
