#!/bin/bash
# compute-sanitizer on the tiny configurations (memcheck everywhere; racecheck on the kernels whose
# shared-memory protocol has no intentional races — the beam kernels' visited table is racy BY DESIGN: a lost
# insertion only costs a recomputation, ws_kernels.cuh ws_seen_warp)
o=gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
timeout 200 $S --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_golden.py tests/test_gpu_direct.py tests/test_gpu_pretree.py tests/test_gpu_group.py -x -q > $o/r2_sanitizer_memcheck.log 2>&1; echo memcheck_rc=$?
tail -6 $o/r2_sanitizer_memcheck.log
timeout 100 $S --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_gemm.py -x -q -k "auto_mode or zero_tiny or non_finite" > $o/r2_sanitizer_memcheck_gemm.log 2>&1; echo memcheck_gemm_rc=$?
tail -6 $o/r2_sanitizer_memcheck_gemm.log
timeout 100 $S --tool racecheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_direct.py tests/test_gpu_pretree.py -x -q > $o/r2_sanitizer_racecheck.log 2>&1; echo racecheck_rc=$?
tail -6 $o/r2_sanitizer_racecheck.log
