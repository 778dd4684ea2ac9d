#!/bin/bash
# runs profiles/scripts/beam_probe.py once per libwsann_cuda.so variant (the box's copy of the repo is scratch)
cd "$(dirname "$0")/../.."
cp rangefilteredann_b200/libwsann_cuda.so /tmp/base.so
echo "== base"; timeout 300 python profiles/scripts/beam_probe.py 80,20 2>&1 | grep -v "^\[wsann\]"
for v in "$@"; do
  cp build_variants/$v/libwsann_cuda.so rangefilteredann_b200/libwsann_cuda.so
  echo "== $v"; timeout 120 python profiles/scripts/beam_probe.py 80,20 2>&1 | grep "^beam"
done
echo "== base, warp_hash=1024"; cp /tmp/base.so rangefilteredann_b200/libwsann_cuda.so; timeout 120 python profiles/scripts/beam_probe.py 80 warp_hash=1024 2>&1 | grep "^beam"
echo "== base, warp_hash=4096"; timeout 120 python profiles/scripts/beam_probe.py 80 warp_hash=4096 2>&1 | grep "^beam"
