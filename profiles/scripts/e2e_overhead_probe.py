"""Where the host-buffer call's time goes: PrefilterIndex.batch_search on tiny windows (kernel ~45 us), 10 000 queries."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from rangefilteredann_b200 import capi, load_engine, synth
eng = load_engine()
n, d, nq = 1_000_000, 128, 10_000
data, queries, labels = synth.make_dataset(n, d, nq, 0)
pre = eng.PrefilterIndexFloatEuclidian(data, labels)
h = capi.Handle.borrow(pre)
w = synth.make_windows(labels, -16, nq, seed=984)
qp = eng.QueryParams(10, 10, 1.35, 10_000_000, 10_000, 1, 10000, None, False)
def t(f, reps=30):
    f(); f()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    return (time.perf_counter() - t0) / reps * 1e3
print("pybind, pageable numpy      %.3f ms" % t(lambda: pre.batch_search(queries, w, nq, qp)))
ids = np.empty((nq, 10), np.uint32); dd = np.empty((nq, 10), np.float32)
print("C ABI, pageable numpy       %.3f ms" % t(lambda: h.prefilter_batch(queries, w, nq, 10, ids, dd)))
pq = capi.pinned_array(queries.shape, np.float32); pq[:] = queries
pw = capi.pinned_array(w.shape, np.float32); pw[:] = w
pi = capi.pinned_array((nq, 10), np.uint32); pd = capi.pinned_array((nq, 10), np.float32)
print("C ABI, pinned in/out        %.3f ms" % t(lambda: h.prefilter_batch(pq, pw, nq, 10, pi, pd)))
print("C ABI, pinned in, pageable out %.3f ms" % t(lambda: h.prefilter_batch(pq, pw, nq, 10, ids, dd)))
dq, dw = h.dalloc(queries.nbytes), h.dalloc(w.nbytes); di, ddd = h.dalloc(nq * 40), h.dalloc(nq * 40)
h.h2d(dq, queries); h.h2d(dw, w)
def dev():
    h.prefilter_batch(dq, dw, nq, 10, di, ddd, device_ptrs=True); h.sync()
print("device pointers + sync      %.3f ms" % t(dev))
print("memcpy 5 MB numpy->numpy    %.3f ms" % t(lambda: np.copyto(pq, queries)))
print("host cores", os.cpu_count())
