#!/bin/bash
# BASELINE.json configs[3] (RedCaps shape, rows scaled to what the GPU-minute budget lets the device builder finish)
# and the adversarial half of configs[4], one B200, both arms' numbers in the engine line (cpu_baseline = reference)
o=gpurun_out
timeout 1100 python bench.py --config c4 --rows 1000000 --steps 3 --warmup 3 --cpu-budget-s 20 > $o/r2i_c4.json 2> $o/r2i_c4.err; echo c4_rc=$?
tail -c 700 $o/r2i_c4.err | tail -3; cut -c1-300 $o/r2i_c4.json
timeout 600 python bench.py --config c5adv --steps 3 --warmup 3 --cpu-budget-s 12 > $o/r2i_c5adv.json 2> $o/r2i_c5adv.err; echo c5adv_rc=$?
tail -c 900 $o/r2i_c5adv.err | tail -4; cut -c1-300 $o/r2i_c5adv.json
