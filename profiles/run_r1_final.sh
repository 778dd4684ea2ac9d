#!/bin/bash
# Round-1 closing run on one B200 (under gpurun): GPU tests, both bench arms, one ncu capture of the
# tensor-core prefilter sweep at a known fraction.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --ops-file gpurun_out/ops_r1.json > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_final_ref.json 2> gpurun_out/bench_r1_final_ref.err
for P in -2; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:ws_gemm_topk_kernel -s 1 -c 1 -f -o gpurun_out/ncu_gemm_p$P \
    python profiles/profile_driver.py --config c2 --index prefilter --method prefilter --power $P --reps 2 --opt gemm_prefilter=1 > gpurun_out/ncu_gemm_p$P.log 2>&1
  ncu -i gpurun_out/ncu_gemm_p$P.ncu-rep --page raw --csv > gpurun_out/ncu_gemm_p${P}_raw.csv 2>/dev/null
done
tail -3 gpurun_out/pytest_gpu.log; head -c 600 gpurun_out/bench_r1_final.json; echo; head -c 300 gpurun_out/bench_r1_final_ref.json; echo; tail -4 gpurun_out/ncu_gemm_p-2.log
