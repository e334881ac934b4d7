#!/bin/bash
# Round-2 GPU-box visit. Usage (under gpurun): tools/gpu_r2.sh TAG [tests|notests] [ncu-kernel-regex]
TAG=${1:-r2}
MODE=${2:-tests}
KREG=$3
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
if [ "$MODE" = "tests" ]; then
  timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/${TAG}_pytest_gpu.log 2>&1
  echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
  echo "smoke exit $?" >> gpurun_out/${TAG}_smoke.log
elif [ "$MODE" != "notests" ]; then
  timeout 1200 python -m pytest $MODE -q -m gpu -x > gpurun_out/${TAG}_pytest_gpu.log 2>&1
  echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
fi
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?" >> gpurun_out/${TAG}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-extras \
  > gpurun_out/${TAG}_ncu_bench.log 2>&1
if [ -n "$KREG" ]; then
  for K in $(echo $KREG | tr ',' ' '); do
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f \
      -o gpurun_out/${TAG}_$K python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extras \
      > gpurun_out/${TAG}_ncu_$K.log 2>&1
  done
fi
tail -6 gpurun_out/${TAG}_pytest_gpu.log 2>/dev/null; tail -3 gpurun_out/${TAG}_smoke.log 2>/dev/null
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench.json"))
    print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"] and d["e2e"]["value"], "launches", d["gpu_launches"])
    for k, v in d["kernels"].items(): print("  %-22s %8.3f ms  frac %.4f" % (k, v["ms"], v["frac"] or 0))
    print("legs", d.get("legs")); print("c1", json.dumps(d.get("c1"))); print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/${TAG}_bench.err").read()[-3000:])
PY
