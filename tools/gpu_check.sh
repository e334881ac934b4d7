#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list (+ optional full capture).
# Usage (from the repo root, under gpurun): tools/gpu_check.sh [tag] [full]
TAG=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu -s > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?" >> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
timeout 300 python tools/rt_bench.py > gpurun_out/${TAG}_rt_bench.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline \
  > gpurun_out/${TAG}_ncu_bench.log 2>&1
if [ "$2" = "full" ]; then
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:hm_bank -s 1 -c 1 \
    -o gpurun_out/${TAG}_bank python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline \
    > gpurun_out/${TAG}_ncu_full.log 2>&1
fi
tail -5 gpurun_out/${TAG}_pytest_gpu.log; cat gpurun_out/${TAG}_rt_bench.txt; tail -3 gpurun_out/${TAG}_smoke.log; cat gpurun_out/${TAG}_bench.json; tail -2 gpurun_out/${TAG}_bench.err
