#!/usr/bin/env python
"""TEST INFRASTRUCTURE — builds Vamana graph caches with the UNMODIFIED reference builder.

Runs the reference module compiled into oracle/_ref (see oracle/Makefile) over the
deterministic synthetic datasets of rangefilteredann_b200/synth.py and leaves the
reference's own `.bin` graph files (graph.h:174-196, names postfilter_vamana.h:126-132)
under data_cache/<name>/.  Both the reference and this repo's engine then load those
very files, so they search the identical graph (BASELINE.json north_star (4)).

data_cache/ is git-ignored (large) but travels to the GPU box with the snapshot.

usage: python oracle/build_ref_cache.py <config> [--threads N]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (n, d, nq, seed, angular, tree kinds)
    "tiny": dict(n=3000, d=16, nq=64, seed=7, angular=False, cutoff=500),
    "small": dict(n=20000, d=32, nq=256, seed=3, angular=False, cutoff=1000),
    "small_mips": dict(n=20000, d=100, nq=256, seed=4, angular=True, cutoff=1000),
    "c1": dict(n=100_000, d=128, nq=10_000, seed=0, angular=False, cutoff=1000),
    "c2": dict(n=1_000_000, d=128, nq=10_000, seed=0, angular=False, cutoff=1000),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config")
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    ap.add_argument("--kinds", default="wst", help="comma list of wst,super,flat")
    ap.add_argument("--out", default=os.path.join(ROOT, "data_cache"))
    args = ap.parse_args()
    os.environ["PARLAY_NUM_THREADS"] = str(args.threads)
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    import window_ann as ref  # noqa: E402  (the reference module)
    from rangefilteredann_b200 import synth

    cfg = CONFIGS[args.config]
    data, queries, labels = synth.make_dataset(cfg["n"], cfg["d"], cfg["nq"], cfg["seed"],
                                               cfg["angular"])
    sfx = "FloatMips" if cfg["angular"] else "FloatEuclidian"
    for kind in args.kinds.split(","):
        cache = os.path.join(args.out, args.config, kind) + "/"
        os.makedirs(cache, exist_ok=True)
        bp = ref.BuildParams(64, 500, 1.0, cache)  # run_our_method.py:266-268,309
        t0 = time.time()
        if kind == "wst":
            getattr(ref, "VamanaRangeFilterTreeIndex" + sfx)(data, labels, cfg["cutoff"], 2, bp)
        elif kind == "super":
            getattr(ref, "SuperOptimizedPostfilterTreeIndex" + sfx)(data, labels, cfg["cutoff"],
                                                                    2.0, 0.5, bp)
        elif kind == "flat":  # naive postfiltering over unsorted points (run_our_method.py:263-302)
            getattr(ref, "PostfilterVamanaIndex" + sfx)(data, labels, bp)
        else:
            raise SystemExit("unknown kind " + kind)
        print(f"[build_ref_cache] {args.config}/{kind}: {time.time() - t0:.1f}s, "
              f"{len(os.listdir(cache))} graph files", flush=True)


if __name__ == "__main__":
    main()
