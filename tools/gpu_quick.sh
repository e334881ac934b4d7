#!/bin/bash
# Quick synthesis timing: parity tests of the synthesis path, bench without the side legs, launch list.
TAG=${1:-r1}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_synth.py -x -q -m gpu 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-analysis > gpurun_out/${TAG}_bench_quick.json 2>/dev/null
python -c "import json; d=json.load(open('gpurun_out/${TAG}_bench_quick.json')); print('ms_per_step', d['ms_per_step'], 'value', d['value'], 'bank_ms', d['roofline']['ms_per_launch'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv \
  --log-file gpurun_out/${TAG}_launches_quick.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-analysis > /dev/null 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/${TAG}_launches_quick.csv") if l.startswith('"')))
h = rows[0]; ki = h.index("Kernel Name"); vi = h.index("Metric Value"); ui = h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    try: v = float(r[vi].replace(",", ""))
    except ValueError: continue
    if r[ui] in ("ns", "nsecond"): v /= 1e6
    elif r[ui] in ("us", "usecond"): v /= 1e3
    agg.setdefault(r[ki].split("(")[0], []).append(round(v, 3))
for k, l in agg.items(): print("%-50s n=%2d %s" % (k[:50], len(l), l[:4]))
PY
