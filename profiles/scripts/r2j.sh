#!/bin/bash
o=gpurun_out
for t in 3 0 2 4 6; do echo "== WSANN_COPY_THREADS=$t"; WSANN_COPY_THREADS=$t timeout 100 python profiles/scripts/e2e_overhead_probe.py 2>&1 | head -4; done
bash profiles/scripts/r2i_c4_c5adv.sh
