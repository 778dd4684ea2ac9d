python -m pytest tests -x -q -m gpu 2>&1 | tail -4
L=rangefilteredann_b200/libwsann_cuda.so
cp $L /tmp/lib_base.so
for v in base rows4; do
  if [ $v = base ]; then cp /tmp/lib_base.so $L; else cp rangefilteredann_b200/variants/libwsann_cuda_$v.so $L; fi
  for a in "--method optimized_postfilter --beam 80 --power 0" "--method fenwick --beam 10 --power -4" "--method fenwick --beam 10 --power -8" "--method prefilter --power -10" "--method prefilter --power -12" "--method fenwick --beam 10 --power -4 --opt warp_scan=0"; do
    echo "== $v $a"; python profiles/profile_driver.py --config c2 $a --reps 4 --opt profile_kernels=1 2>&1 | tail -3 | head -2
  done
done
cp /tmp/lib_base.so $L
