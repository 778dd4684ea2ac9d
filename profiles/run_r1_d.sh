mkdir -p gpurun_out
# operating points + bench line of the final build
timeout 600 python bench.py --steps 5 --warmup 3 --ops-file gpurun_out/ops_r1d.json > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err
# launch list of one step of the same command (per-launch times are cold-cache and serialised: shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ws_ -c 2000 --csv --log-file gpurun_out/launches_r1d.csv \
  python bench.py --no-cpu --steps 1 --warmup 3 --ops-file gpurun_out/ops_r1d.json > gpurun_out/launches_r1d.log 2>&1
# full capture of the sweep kernel at 2^-2 and of the one-launch prefilter kernel at 2^-14
timeout 300 ncu --set full --clock-control none -k regex:ws_gemm_topk_kernel -s 1 -c 1 -f -o gpurun_out/ncu_gemm_dyn_p-2 \
  python profiles/profile_driver.py --config c2 --index prefilter --method prefilter --power -2 --reps 2 --opt gemm_prefilter=1 > gpurun_out/ncu_gemm_dyn_p-2.log 2>&1
ncu -i gpurun_out/ncu_gemm_dyn_p-2.ncu-rep --page raw --csv > gpurun_out/ncu_gemm_dyn_p-2_raw.csv 2>/dev/null
rm -f gpurun_out/ncu_gemm_dyn_p-2.ncu-rep  # gpurun_out/ travels back only below 64 MiB
timeout 300 ncu --set full --clock-control none -k regex:ws_prefilter_direct_kernel -s 1 -c 1 -f -o gpurun_out/ncu_direct_p-14 \
  python profiles/profile_driver.py --config c2 --index prefilter --method prefilter --power -14 --reps 2 --opt gemm_prefilter=0 --opt prefilter_direct=1 > gpurun_out/ncu_direct_p-14.log 2>&1
ncu -i gpurun_out/ncu_direct_p-14.ncu-rep --page raw --csv > gpurun_out/ncu_direct_p-14_raw.csv 2>/dev/null
rm -f gpurun_out/ncu_direct_p-14.ncu-rep
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1d_ref.json 2> gpurun_out/bench_r1d_ref.err
head -c 300 gpurun_out/bench_r1d.json; echo; head -c 200 gpurun_out/bench_r1d_ref.json; echo; wc -l gpurun_out/launches_r1d.csv; tail -2 gpurun_out/ncu_gemm_dyn_p-2.log; tail -2 gpurun_out/ncu_direct_p-14.log
du -sh gpurun_out
