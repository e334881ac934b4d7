#!/bin/bash
# Quick GPU visit: analysis/synthesis parity tests, then the per-kernel bench table once per environment setting.
# Usage (under gpurun): tools/gpu_kern.sh TAG [tests-or-none] ["ENV=V ENV2=W" ...]
TAG=${1:-k}; TESTS=${2:-none}; shift; shift
mkdir -p gpurun_out
if [ "$TESTS" != "none" ]; then
  timeout 900 python -m pytest $TESTS -q -m gpu -x > gpurun_out/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log; tail -4 gpurun_out/${TAG}_pytest.log
fi
[ $# -eq 0 ] && set -- "LLSM_NOP=1"
N=0
for E in "$@"; do
  N=$((N+1))
  env $E timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/${TAG}_bench$N.json 2> gpurun_out/${TAG}_bench$N.err
  echo "== $E"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench$N.json"))
    print("value %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]))
    print(" ".join("%s=%.3f" % (k, v["ms"]) for k, v in d["kernels"].items()))
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/${TAG}_bench$N.err").read()[-2000:])
PY
done
