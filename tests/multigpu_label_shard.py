#!/usr/bin/env python
"""2+ GPU check of the label-range sharded mode (run under torchrun on the GPU box):
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multigpu_label_shard.py
Every rank builds its shard's sub-tree (device-side builder), answers the whole batch, the
rows are all-gathered over NCCL and merged by ws_merge_partial_topk; rank 0 checks recall
against brute-force ground truth and that every rank ended with identical rows."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from rangefilteredann_b200 import label_shard, load_engine, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ["WSANN_DEVICE"] = str(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    eng = load_engine()
    n, d, nq = 200_000, 96, 2000  # Deep-shaped rows (96-d L2), scaled down
    data, queries, labels = synth.make_dataset(n, d, nq, seed=5)
    cache = os.path.join(tempfile.gettempdir(), "wsann_label_shard")
    tree = label_shard.LabelShardedTree(data, labels, rank, world, cache, cutoff=1000)
    qp = eng.QueryParams(10, 20, 1.35, 10_000_000, 10_000, 1, 10000, None, False)
    ok = True
    for power in (-8, -3, 0):
        w = synth.make_windows(labels, power, nq, seed=77 + power)
        ids, dd = tree.batch_search(queries, w, "fenwick", qp, 10)
        # identical rows on every rank
        t = torch.from_numpy(ids.view(np.int32).copy()).cuda()
        ref = t.clone()
        dist.broadcast(ref, src=0)
        same = bool((t == ref).all().item())
        if rank == 0:
            gt = synth.ground_truth(data, queries, labels, w)
            r = synth.recall_std(ids, gt)
            print(f"[label-shard] world {world} fraction 2^{power}: recall@10 {r:.4f} rows-identical {same}", flush=True)
            ok = ok and r >= 0.95 and same and bool((np.diff(dd, axis=1) >= 0).all())
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("LABEL_SHARD_OK" if ok else "LABEL_SHARD_FAIL", flush=True)
        sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
