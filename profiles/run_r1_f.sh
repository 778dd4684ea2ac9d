mkdir -p gpurun_out
( timeout 150 python -m pytest tests/test_gpu_config_shapes.py -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/pytest_gpu6.log
timeout 150 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err
tail -3 gpurun_out/pytest_gpu6.log; head -c 250 gpurun_out/bench_r1f.json; echo; tail -2 gpurun_out/bench_r1f.err
