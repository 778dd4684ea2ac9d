mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_r1_n2.json 2> gpurun_out/bench_r1_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 \
  bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_r1_n2_ref.json 2> gpurun_out/bench_r1_n2_ref.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 tests/multigpu_label_shard.py > gpurun_out/label_shard_n2.log 2>&1
head -c 500 gpurun_out/bench_r1_n2.json; echo; head -c 200 gpurun_out/bench_r1_n2_ref.json; echo; tail -3 gpurun_out/label_shard_n2.log; tail -3 gpurun_out/bench_r1_n2.err
