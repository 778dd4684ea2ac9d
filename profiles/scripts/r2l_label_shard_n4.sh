#!/bin/bash
# BASELINE.json configs[4] shape (Deep: 96-d L2), label-range sharded over N GPUs, one process per GPU, exchange by
# ncclAllGather inside libwsann_cuda.so; rows scaled to 1 M per GPU (the device builder's share of the GPU-minute budget)
N=${1:-4}
o=gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --steps 5 --warmup 3 --mode label_shard --config c5 --rows $((N * 1000000)) > $o/r2l_label_shard_n$N.json 2> $o/r2l_label_shard_n$N.err; echo label_shard_rc=$?
tail -c 900 $o/r2l_label_shard_n$N.err | tail -3 | cut -c1-400; cut -c1-600 $o/r2l_label_shard_n$N.json
