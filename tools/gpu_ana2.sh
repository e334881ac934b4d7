#!/bin/bash
# analysis-only GPU visit: parity tests of the analysis path, timing, launch list
TAG=${1:-r1}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_analysis.py tests/test_gpu_compat.py -x -q -m gpu > gpurun_out/${TAG}_pytest_ana.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_ana.log
timeout 600 python tools/ana_bench.py > gpurun_out/${TAG}_ana_bench.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/${TAG}_ana_launches.csv python tools/ana_bench.py --once --batch 128 > gpurun_out/${TAG}_ncu_ana.log 2>&1
tail -3 gpurun_out/${TAG}_pytest_ana.log; cat gpurun_out/${TAG}_ana_bench.txt
python - <<PY
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/${TAG}_ana_launches.csv") if l.startswith('"')))
h = rows[0]; ki = h.index("Kernel Name"); vi = h.index("Metric Value"); ui = h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    if r[ui] in ("ns", "nsecond"): v /= 1e6
    elif r[ui] in ("us", "usecond"): v /= 1e3
    elif r[ui] in ("s", "second"): v *= 1e3
    k = r[ki].split("(")[0]
    a = agg.setdefault(k, [0, 0.0, []]); a[0] += 1; a[1] += v; a[2].append(round(v, 3))
for k, (n, ms, l) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-50s n=%3d  %.3f ms  %s" % (k[:50], n, ms, l[:6]))
PY
