#!/bin/bash
# builds variants of libwsann_cuda.so (compile-time knobs of the warp beam kernel) under build_variants/<name>/
set -e
cd "$(dirname "$0")/../.."
mk() { name=$1; shift; mkdir -p build_variants/$name; WSANN_LIB_OUT=$PWD/build_variants/$name/libwsann_cuda.so WSANN_OBJ_DIR=$PWD/build_variants/$name/obj WSANN_NVCC_EXTRA="$*" python -c "from rangefilteredann_b200 import build as b; b.build_cuda()"; echo built $name; }
for v in "$@"; do
  case $v in
    mb8) mk mb8 -DWS_WARP_MINBLOCKS=8 ;;
    rows3) mk rows3 -DWS_BEAM_ROWS=3 ;;
    rows1) mk rows1 -DWS_BEAM_ROWS=1 ;;
    pf2) mk pf2 -DWS_BEAM_PREFETCH=2 ;;
    mb8rows1) mk mb8rows1 -DWS_WARP_MINBLOCKS=8 -DWS_BEAM_ROWS=1 ;;
    mb6rows3) mk mb6rows3 -DWS_WARP_MINBLOCKS=6 -DWS_BEAM_ROWS=3 ;;
    mb5rows4) mk mb5rows4 -DWS_WARP_MINBLOCKS=5 -DWS_BEAM_ROWS=4 ;;
    *) echo unknown $v ;;
  esac
done
