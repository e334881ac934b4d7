#!/bin/bash
# ncu --set full capture of named kernels during a short bench run; the summaries (tools/ncu_summary.py, tools/ncu_lines.py)
# are produced on the GPU box and only they travel back (gpurun_out is capped at 64 MiB). Usage: tools/gpu_profile2.sh tag regex...
TAG=$1; shift
mkdir -p gpurun_out /tmp/ncu
for K in "$@"; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f \
    -o /tmp/ncu/${TAG}_$K python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extras \
    > gpurun_out/${TAG}_ncu_$K.log 2>&1
  python tools/ncu_summary.py /tmp/ncu/${TAG}_$K.ncu-rep > gpurun_out/${TAG}_$K.ncu.json 2>&1
  python tools/ncu_lines.py /tmp/ncu/${TAG}_$K.ncu-rep 30 > gpurun_out/${TAG}_$K.lines.txt 2>&1
  grep -E "Kernel Name|gpu__time_duration|issue_active|warps_active" gpurun_out/${TAG}_$K.ncu.json | tr -d '\n' | cut -c1-400; echo
done
