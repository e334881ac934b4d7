#!/bin/bash
# e2e A/B: bench.py (headline + e2e, no side legs) once per environment setting.
# Usage (under gpurun): tools/gpu_e2e.sh TAG [tests-or-none] ["ENV=V ENV2=W" ...]
TAG=${1:-e}; TESTS=${2:-none}; shift; shift
mkdir -p gpurun_out
if [ "$TESTS" != "none" ]; then
  timeout 900 python -m pytest $TESTS -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log; grep -E "^E   .*Assertion|passed|failed|^FAILED|pytest exit" gpurun_out/${TAG}_pytest.log
fi
N=0
for E in "$@"; do
  N=$((N+1))
  env $E timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_e2e$N.json 2> gpurun_out/${TAG}_e2e$N.err
  echo "== $E"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_e2e$N.json"))
    print("value %.0f ms/step %.3f  e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/${TAG}_e2e$N.err").read()[-2000:])
PY
done
