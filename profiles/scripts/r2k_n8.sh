#!/bin/bash
# 8 x B200: strong scaling of the replicated mode (one process per GPU), the in-engine group (one process), and the
# label-sharded mode on the Deep shape (one process per GPU, in-library NCCL); every command bounded
N=${1:-8}
o=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
nvidia-smi topo -m 2>&1 | head -12 > $o/r2k_topo_n$N.txt
timeout 700 $TR 29611 bench.py --gpus $N --steps 5 --warmup 3 --scaling strong --no-cpu > $o/r2k_strong_n$N.json 2> $o/r2k_strong_n$N.err; echo strong_rc=$?
timeout 500 python bench.py --mode group --group-gpus $N --steps 5 --warmup 3 > $o/r2k_group_rep_n$N.json 2> $o/r2k_group_rep_n$N.err; echo group_rep_rc=$?
timeout 600 $TR 29612 bench.py --gpus $N --steps 5 --warmup 3 --mode label_shard --config c5 --rows $((N * 1000000)) > $o/r2k_label_shard_n$N.json 2> $o/r2k_label_shard_n$N.err; echo label_shard_rc=$?
timeout 400 python bench.py --mode group --group-gpus $N --group-shard label --config c5 --rows $((N * 1000000)) --steps 5 --warmup 3 --powers=-12,-8,-4,-2,0 > $o/r2k_group_label_n$N.json 2> $o/r2k_group_label_n$N.err; echo group_label_rc=$?
for f in strong group_rep label_shard group_label; do echo "== $f"; tail -c 500 $o/r2k_${f}_n$N.err | tail -2 | cut -c1-300; cut -c1-330 $o/r2k_${f}_n$N.json; done
