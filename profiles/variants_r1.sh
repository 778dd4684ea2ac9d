L=rangefilteredann_b200/libwsann_cuda.so
cp $L /tmp/lib_base.so
for v in base mb7 mb8; do
  if [ $v = base ]; then cp /tmp/lib_base.so $L; else cp rangefilteredann_b200/variants/libwsann_cuda_$v.so $L; fi
  for a in "--method optimized_postfilter --beam 80 --power 0" "--method optimized_postfilter --beam 80 --power 0 --opt warp_hash=4096" "--method fenwick --beam 20 --power -3" "--method fenwick --beam 10 --power -5"; do
    echo "== $v $a"; python profiles/profile_driver.py --config c2 $a --reps 5 2>&1 | grep -E "rep 4"
  done
done
cp /tmp/lib_base.so $L
