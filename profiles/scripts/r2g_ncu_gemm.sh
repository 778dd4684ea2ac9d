#!/bin/bash
# ncu --set full of the prefilter sweep kernel (fp16), fractions 2^-2 and 2^-8, second launch of each
o=gpurun_out
for p in -2 -8; do
  ncu --set full --clock-control none --import-source on -k regex:ws_gemm_topk_kernel -s 1 -c 1 -f -o $o/r2g_gemm_p$p \
    python profiles/gemm_probe.py --n 1000000 --d 128 --nq 10000 --powers=$p --reps 2 --skip-scan-above 0 > $o/r2g_ncu_p$p.log 2>&1
  echo "ncu rc=$? p=$p"
  ncu -i $o/r2g_gemm_p$p.ncu-rep --page raw --csv > $o/r2g_gemm_p${p}_raw.csv 2>/dev/null
  ncu -i $o/r2g_gemm_p$p.ncu-rep --page source --csv > $o/r2g_gemm_p${p}_source.csv 2>/dev/null
done
ls -la $o/r2g_*
