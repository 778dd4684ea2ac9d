#!/usr/bin/env python
"""Tensor-core prefilter vs streaming-scan prefilter on one data set: identical rows? how fast?

  python profiles/gemm_probe.py --n 1000000 --d 128 --nq 10000 --powers 0,-3,-6 --reps 3
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rangefilteredann_b200 import capi, load_engine, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--d", type=int, default=128)
    ap.add_argument("--nq", type=int, default=10000)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--powers", default="0,-3,-6")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--angular", action="store_true")
    ap.add_argument("--skip-scan-above", type=float, default=1e12, help="skip the scan arm when nq*window exceeds this")
    ap.add_argument("--opt", action="append", default=[])
    a = ap.parse_args()
    eng = load_engine()
    data, queries, labels = synth.make_dataset(a.n, a.d, a.nq, 0, a.angular)
    sfx = "FloatMips" if a.angular else "FloatEuclidian"
    t0 = time.time()
    pre = getattr(eng, "PrefilterIndex" + sfx)(data, labels)
    print(f"index ready in {time.time() - t0:.1f}s", flush=True)
    h = capi.Handle.borrow(pre)
    h.set_option("profile_kernels", 1)
    for o in a.opt:
        name, v = o.split("=")
        h.set_option(name, int(v))
    k = a.k
    dq = h.dalloc(queries.nbytes)
    h.h2d(dq, queries)
    di, dd = h.dalloc(a.nq * k * 4), h.dalloc(a.nq * k * 4)
    for power in [int(x) for x in a.powers.split(",")]:
        w = synth.make_windows(labels, power, a.nq, seed=1000 + power)
        dw = h.dalloc(w.nbytes)
        h.h2d(dw, w)
        out = {}
        window = a.n * 2.0 ** power
        for mode in (1, 0):
            if mode == 0 and a.nq * window > a.skip_scan_above:
                continue
            h.set_option("gemm_prefilter", mode)
            h.reset_stats()
            h.kernel_times(reset=True)
            best = 1e30
            for r in range(a.reps):
                h.timer_start()
                h.prefilter_batch(dq, dw, a.nq, k, di, dd, device_ptrs=True)
                best = min(best, h.timer_stop())
            ids = np.empty((a.nq, k), np.uint32)
            dists = np.empty((a.nq, k), np.float32)
            h.d2h(ids, di)
            h.d2h(dists, dd)
            kt = {kk: round(v["ms"] / a.reps, 3) for kk, v in h.kernel_times().items()}
            out[mode] = (ids, dists)
            flops = 2.0 * a.nq * window * a.d
            print(f"2^{power} {'gemm' if mode else 'scan'}: best {best:.3f} ms  {a.nq / best * 1000:.0f} qps  "
                  f"{flops / best / 1e9:.1f} TFLOP/s(useful)  kernels {kt}", flush=True)
        if 0 in out and 1 in out:
            same_i = np.array_equal(out[0][0], out[1][0])
            same_d = np.array_equal(out[0][1].view(np.uint32), out[1][1].view(np.uint32))
            bad = np.nonzero((out[0][0] != out[1][0]).any(axis=1))[0]
            print(f"2^{power}: ids identical {same_i}, dists bit-identical {same_d}, bad rows {len(bad)} {bad[:8]}", flush=True)
            if len(bad):
                r = bad[0]
                print(" scan", out[0][0][r], out[0][1][r])
                print(" gemm", out[1][0][r], out[1][1][r])
        h.dfree(dw)


if __name__ == "__main__":
    main()
