#!/bin/bash
# 1-GPU smoke of every bench configuration / mode at reduced row counts (plumbing check, not a measurement)
set -x
o=gpurun_out
true
python bench.py --config c3 --rows 100000 --steps 2 --powers=-10,-6,-3,0 --cpu-budget-s 6 > $o/r2e_c3.json 2> $o/r2e_c3.err; echo c3_rc=$?
python bench.py --config c3 --rows 100000 --impl reference --steps 1 --warmup 1 --powers=-10,-6,-3,0 > $o/r2e_c3_ref.json 2> $o/r2e_c3_ref.err; echo c3ref_rc=$?
python bench.py --config c4 --rows 60000 --steps 2 --powers=-8,-4,0 --no-cpu > $o/r2e_c4.json 2> $o/r2e_c4.err; echo c4_rc=$?
true
python bench.py --config c5 --rows 200000 --mode label_shard --steps 2 --powers=-10,-5,-2,0 > $o/r2e_ls.json 2> $o/r2e_ls.err; echo ls_rc=$?
python bench.py --config c1 --mode group --group-devices 0,0 --steps 2 --powers=-10,-5,-2,0 > $o/r2e_group.json 2> $o/r2e_group.err; echo group_rc=$?
python bench.py --config c1 --mode group --group-devices 0,0 --group-shard label --steps 2 --powers=-10,-5,-2,0 > $o/r2e_groupl.json 2> $o/r2e_groupl.err; echo groupl_rc=$?
for f in c1 c3 c3_ref c4 c5adv ls group groupl; do echo "== $f"; tail -c 600 $o/r2e_$f.err | tail -4; cut -c1-300 $o/r2e_$f.json; done
