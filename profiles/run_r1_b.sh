mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/pytest_gpu2.log
timeout 600 python profiles/gemm_knob_sweep.py > gpurun_out/gemm_knobs.txt 2> gpurun_out/gemm_knobs.err
timeout 600 python bench.py --steps 5 --warmup 3 --ops-file gpurun_out/ops_r1b.json > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err
tail -5 gpurun_out/pytest_gpu2.log; cat gpurun_out/gemm_knobs.txt; head -c 400 gpurun_out/bench_r1b.json
