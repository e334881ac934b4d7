#!/bin/bash
# ncu --set full capture of analysis kernels (one launch each). Usage: tools/gpu_profile_ana.sh tag kernel [kernel...]
TAG=$1; shift
mkdir -p gpurun_out
for K in "$@"; do
  SKIP=0
  if [ "$K" = "iir_filtfilt" ]; then SKIP=1; fi   # first iir launch belongs to the synthesis that makes the input
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 \
    -o gpurun_out/${TAG}_$K python tools/ana_bench.py --once --batch 128 > gpurun_out/${TAG}_ncu_$K.log 2>&1
  tail -1 gpurun_out/${TAG}_ncu_$K.log
done
