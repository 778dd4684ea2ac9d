"""K2 (ws_beam_warp_kernel) alone: the root-node search of config 2 (1 M x 128, one Vamana graph over all points),
10 000 queries, beam 80, window = everything — the operating point of fraction 2^0 in bench.py."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from rangefilteredann_b200 import capi, load_engine, synth
eng = load_engine()
n, d, nq = 1_000_000, 128, 10_000
beams = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "80").split(",")]
data, queries, labels = synth.make_dataset(n, d, nq, 0)
cache = os.path.join(ROOT, "data_cache", "c2", "flat") + "/"
os.makedirs(cache, exist_ok=True)
t0 = time.time()
flat = eng.PostfilterVamanaIndexFloatEuclidian(data, labels, eng.BuildParams(64, 500, 1.0, cache))
print(f"flat index ready in {time.time() - t0:.1f}s", flush=True)
h = capi.Handle.borrow(flat)
for o in sys.argv[2:]:
    name, v = o.split("=")
    h.set_option(name, int(v))
w = np.tile(np.array([[-1.0, 2.0]], np.float32), (nq, 1))
dq, dw = h.dalloc(queries.nbytes), h.dalloc(w.nbytes)
di, dd = h.dalloc(nq * 40), h.dalloc(nq * 40)
h.h2d(dq, queries); h.h2d(dw, w)
gt = None
for beam in beams:
    qp = capi.query_params(k=10, beam=beam)
    best = 1e9
    h.reset_stats()
    for r in range(6):
        h.timer_start()
        h.postfilter_batch(0, dq, dw, nq, qp, 1, di, dd, device_ptrs=True)
        best = min(best, h.timer_stop())
    st = h.stats()
    gbytes = (st["visited"] * 256 + st["dist_cmps"] * 512 + st["beam_sum"] * 4) / 6 / 1e9
    ids = np.empty((nq, 10), np.uint32); h.d2h(ids, di)
    print(f"beam {beam}: best {best:.3f} ms  {gbytes:.2f} GB algorithmic -> {gbytes / best * 1e3:.0f} GB/s  visited/search {st['visited'] / 6 / nq:.1f} cmps/search {st['dist_cmps'] / 6 / nq:.0f}  checksum {int(ids.astype(np.uint64).sum())}", flush=True)
