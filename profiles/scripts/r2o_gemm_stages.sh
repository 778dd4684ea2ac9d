#!/bin/bash
# sweep-kernel timing experiments (results invalid for gemm_debug != 0): 4 = epilogue reads TMEM and releases, nothing else;
# 8 = reads, converts, filters, keeps nothing; 1 = epilogue releases stages unread
for dbg in 0 8 4 1; do
  echo "== gemm_debug=$dbg"
  timeout 60 python profiles/gemm_probe.py --n 1000000 --d 128 --nq 10000 --powers 0,-2,-6,-10 --reps 3 --skip-scan-above 0 --opt gemm_debug=$dbg 2>&1 | grep gemm | cut -c1-120
done
