#!/bin/bash
# (1) where the fp16 sweep's time goes: timing experiments with the epilogue / the TMA loads switched off (results
#     invalid in those modes); (2) BASELINE.json configs[2] at full size, both arms
o=gpurun_out
for dbg in 0 1 2 3; do
  echo "== gemm_debug=$dbg (1: epilogue releases stages unread, 2: no TMA loads, 3: both)"
  python profiles/gemm_probe.py --n 1000000 --d 128 --nq 10000 --powers 0,-2,-6,-10 --reps 3 --skip-scan-above 0 --opt gemm_debug=$dbg 2>&1 | grep gemm
done > $o/r2f_gemm_diag.log 2>&1
cat $o/r2f_gemm_diag.log
python bench.py --config c3 --steps 3 --warmup 3 --cpu-budget-s 20 > $o/r2f_c3.json 2> $o/r2f_c3.err; echo c3_rc=$?
tail -c 1200 $o/r2f_c3.err
python bench.py --config c3 --impl reference --steps 1 --warmup 1 > $o/r2f_c3_ref.json 2> $o/r2f_c3_ref.err; echo c3ref_rc=$?
cut -c1-400 $o/r2f_c3.json; cut -c1-300 $o/r2f_c3_ref.json
