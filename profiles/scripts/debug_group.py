import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["WSANN_GRAPH_BUILD"] = "0"
from golden_cases import TINY
from rangefilteredann_b200 import capi, load_engine, synth
eng = load_engine()
data, queries, labels = synth.make_dataset(TINY["n"], TINY["d"], TINY["nq"], TINY["seed"])
cache = os.path.join(ROOT, "tests", "golden", "tiny", "wst") + "/"
bp = eng.BuildParams(64, 500, 1.0, cache)
single = eng.VamanaRangeFilterTreeIndexFloatEuclidian(data, labels, TINY["cutoff"], 2, bp)
os.environ["WSANN_DEVICES"] = "0,0,0"
tree3 = eng.VamanaRangeFilterTreeIndexFloatEuclidian(data, labels, TINY["cutoff"], 2, bp)
g = capi.Group.borrow(tree3)
nq = len(queries)
w = synth.make_windows(labels, -1, nq, seed=499)
qp = capi.query_params(k=10, beam=10)
def run(h, lo, hi):
    ids = np.empty((hi - lo, 10), np.uint32); d = np.empty((hi - lo, 10), np.float32)
    h.tree_batch("fenwick", np.ascontiguousarray(queries[lo:hi]), np.ascontiguousarray(w[lo:hi]), hi - lo, qp, ids, d)
    return ids, d
hs = capi.Handle.borrow(single)
full = run(hs, 0, nq)
print("single repeat identical:", np.array_equal(full[0], run(hs, 0, nq)[0]))
for i in range(3):
    r = run(g.member(i), 0, nq)
    print("member", i, "full batch identical:", np.array_equal(full[0], r[0]), np.nonzero((full[0] != r[0]).any(1))[0][:10])
third = nq // 3
for lo, hi in ((0, third), (third, 2 * third), (2 * third, nq), (0, 17)):
    r = run(hs, lo, hi)
    bad = np.nonzero((full[0][lo:hi] != r[0]).any(1))[0]
    print("single slice", lo, hi, "identical:", len(bad) == 0, bad[:10])
    if len(bad):
        b = bad[0]
        print(" full :", full[0][lo + b], full[1][lo + b]); print(" slice:", r[0][b], r[1][b]); print(" window", w[lo + b])
for t in range(3):
    ids, d = tree3.batch_search(queries, w, nq, "fenwick", eng.QueryParams(10, 10, 1.35, 10_000_000, 10_000, 1, 10000, None, False))
    bad = np.nonzero((full[0] != ids).any(1))[0]
    print("group call", t, "bad rows", bad[:10], len(bad))
