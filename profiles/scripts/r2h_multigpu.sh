#!/bin/bash
# N-GPU measurements (N = $1): in-engine groups, strong scaling, label-sharded mode
N=${1:-2}
o=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
timeout 600 python tests/multigpu_group.py > $o/r2h_group_test_n$N.log 2>&1; echo group_test_rc=$?; tail -3 $o/r2h_group_test_n$N.log
timeout 900 $TR 29511 bench.py --gpus $N --steps 5 --warmup 3 --scaling strong --no-cpu > $o/r2h_strong_n$N.json 2> $o/r2h_strong_n$N.err; echo strong_rc=$?
timeout 900 python bench.py --mode group --group-gpus $N --steps 5 --warmup 3 > $o/r2h_group_rep_n$N.json 2> $o/r2h_group_rep_n$N.err; echo group_rep_rc=$?
timeout 900 python bench.py --mode group --group-gpus $N --group-shard label --steps 5 --warmup 3 > $o/r2h_group_label_n$N.json 2> $o/r2h_group_label_n$N.err; echo group_label_rc=$?
timeout 900 $TR 29512 bench.py --gpus $N --steps 5 --warmup 3 --mode label_shard --config c5 --rows $((N * 1000000)) > $o/r2h_label_shard_n$N.json 2> $o/r2h_label_shard_n$N.err; echo label_shard_rc=$?
timeout 600 $TR 29513 tests/multigpu_label_shard.py > $o/r2h_label_shard_test_n$N.log 2>&1; echo ls_test_rc=$?; tail -2 $o/r2h_label_shard_test_n$N.log
for f in strong group_rep group_label label_shard; do echo "== $f"; tail -c 400 $o/r2h_${f}_n$N.err | tail -3; cut -c1-330 $o/r2h_${f}_n$N.json; done
