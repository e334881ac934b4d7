#!/bin/bash
# Final visit of a session: everything the driver will run (tests, smoke, bench, reference arm), the streaming timing,
# the ncu launch list of the bench command, and ncu --set full captures (summarised on the box) of the kernels named.
# Usage (under gpurun): tools/gpu_final.sh TAG [kernel ...]
TAG=${1:-fin}; shift
tools/gpu_check.sh $TAG > gpurun_out/${TAG}_check.log 2>&1
python tools/parity_diag.py > gpurun_out/${TAG}_parity_diag.txt 2>&1
if [ $# -gt 0 ]; then
  tools/gpu_profile_ana.sh $TAG "$@" > gpurun_out/${TAG}_profile.log 2>&1
  for K in "$@"; do
    if [ -f gpurun_out/${TAG}_$K.ncu-rep ]; then
      python tools/ncu_summary.py gpurun_out/${TAG}_$K.ncu-rep > gpurun_out/${TAG}_$K.ncu.json 2>/dev/null
      python tools/ncu_lines.py gpurun_out/${TAG}_$K.ncu-rep > gpurun_out/${TAG}_$K.lines.txt 2>/dev/null
    fi
  done
fi
tail -3 gpurun_out/${TAG}_pytest_gpu.log; tail -2 gpurun_out/${TAG}_smoke.log; cat gpurun_out/${TAG}_parity_diag.txt | cut -c1-400
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench.json"))
print("value %.0f ms/step %.3f e2e %.0f analysis %.3f synthesis %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["legs"]["analysis"]["ms"], d["legs"]["synthesis"]["ms"]))
print(" ".join("%s=%.3f" % (k, v["ms"]) for k, v in d["kernels"].items()))
r = json.load(open("gpurun_out/${TAG}_bench_ref.json")); print("reference arm", r["value"], r["cpu_baseline"]["cores"])
PY
