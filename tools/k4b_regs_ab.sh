#!/bin/bash
# Registers per thread vs resident blocks of the block-per-reactor N2 solver at the production size of config 3 (4,096 x 100 nodes):
# 128 registers x 4 blocks per SM (default), 168 x 3, 255 x 2.  Each leg is its own process (the knobs are build-time).
mkdir -p gpurun_out
for leg in "default;;;" "r168x3;PBR_K4B_MAXTHREADS=128;3;3" "r255x2;PBR_K4B_MAXTHREADS=128;2;2" "r128x4_of128;PBR_K4B_MAXTHREADS=128;4;4"; do
  IFS=';' read -r tag defs minb persm <<< "$leg"
  echo "== $tag"
  if [ -z "$defs" ]; then
    timeout 300 python tools/k4b_march_ab.py 4096 100
  else
    PBR_EXTRA_DEFINES="$defs" PBR_K4B_MINBLOCKS=$minb PBR_K4B_PER_SM=$persm timeout 300 python tools/k4b_march_ab.py 4096 100
  fi
done
