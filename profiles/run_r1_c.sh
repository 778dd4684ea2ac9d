mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/pytest_gpu3.log
POWERS=-10,-6,-4,-2,-1,0 ITEMS=0,592 CHUNKS=32,16,8,4 DYN=0,1 TOP=16 timeout 600 python profiles/gemm_knob_sweep.py > gpurun_out/gemm_knobs2.txt 2> gpurun_out/gemm_knobs2.err
timeout 600 python bench.py --steps 5 --warmup 3 --ops-file gpurun_out/ops_r1c.json > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err
tail -5 gpurun_out/pytest_gpu3.log; cat gpurun_out/gemm_knobs2.txt; tail -3 gpurun_out/gemm_knobs2.err; head -c 300 gpurun_out/bench_r1c.json
