#!/bin/bash
# ncu --set full capture of named kernels during a short bench run. Usage: tools/gpu_profile.sh tag regex [regex...]
TAG=$1; shift
mkdir -p gpurun_out
for K in "$@"; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 \
    -o gpurun_out/${TAG}_$K python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline \
    > gpurun_out/${TAG}_ncu_$K.log 2>&1
  tail -2 gpurun_out/${TAG}_ncu_$K.log
done
