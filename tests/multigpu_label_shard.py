#!/usr/bin/env python
"""2+ GPU check of the label-range sharded mode with ONE PROCESS PER GPU (run under torchrun on the GPU box):
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multigpu_label_shard.py
Every rank builds its shard's sub-tree (device-side builder) and answers the whole batch; the rows are all-gathered
and merged inside libwsann_cuda.so (ncclAllGather on the index stream + merge kernel — torch only launches the
ranks and hands rank 0's NCCL id around).  Rank 0 checks recall against brute-force ground truth, prefilter rows
against a single-GPU index, and that every rank ended with identical rows."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from rangefilteredann_b200 import capi, label_shard, load_engine, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ["WSANN_DEVICE"] = str(local)
    dist.init_process_group("gloo")  # plumbing only: the data path's NCCL communicator lives in the library
    eng = load_engine()
    n, d, nq = 200_000, 96, 2000  # Deep-shaped rows (96-d L2), scaled down
    data, queries, labels = synth.make_dataset(n, d, nq, seed=5)
    cache = os.path.join(tempfile.gettempdir(), "wsann_label_shard")
    uid = bench.broadcast_bytes(capi.nccl_unique_id() if rank == 0 else None)
    tree = label_shard.LabelShardedTree(data, labels, rank, world, cache, cutoff=1000, unique_id=uid)
    qp = capi.query_params(k=10, beam=20)
    pre1 = eng.PrefilterIndexFloatEuclidian(data, labels) if rank == 0 else None
    ok = True
    for power in (-8, -3, 0):
        w = synth.make_windows(labels, power, nq, seed=77 + power)
        for method in ("fenwick", "prefilter"):
            ids, dd = tree.batch_search(queries, w, method, qp, 10)
            rows = [None] * world
            dist.all_gather_object(rows, ids)
            same = all(np.array_equal(rows[0], r) for r in rows)
            if rank == 0:
                if method == "prefilter":
                    eids, ed = pre1.batch_search(queries, w, nq, eng.QueryParams(10, 10, 1.35, 10_000_000, 10_000, 1, 10000, None, False))
                    good = np.array_equal(ids, eids) and np.array_equal(dd.view(np.uint32), ed.view(np.uint32))
                    print(f"[label-shard] world {world} fraction 2^{power} prefilter: rows identical to 1 GPU {good}, "
                          f"identical on every rank {same}", flush=True)
                    ok = ok and good and same
                else:
                    gt = synth.ground_truth(data, queries, labels, w)
                    r = synth.recall_std(ids, gt)
                    print(f"[label-shard] world {world} fraction 2^{power} {method}: recall@10 {r:.4f} rows-identical {same}", flush=True)
                    ok = ok and r >= 0.95 and same and bool((np.diff(dd, axis=1) >= 0).all())
    dist.barrier()
    tree.h.comm_destroy()
    dist.destroy_process_group()
    if rank == 0:
        print("LABEL_SHARD_OK" if ok else "LABEL_SHARD_FAIL", flush=True)
        sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
