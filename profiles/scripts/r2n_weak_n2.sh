#!/bin/bash
# what the driver's SCALE run does at N=2: default flags under torchrun (weak scaling, replicas)
o=gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 5 --warmup 3 > $o/r2n_weak_n2.json 2> $o/r2n_weak_n2.err; echo weak_rc=$?
tail -c 400 $o/r2n_weak_n2.err | tail -2 | cut -c1-300; cut -c1-420 $o/r2n_weak_n2.json
