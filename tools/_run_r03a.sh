mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r03b_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r03b_pytest_gpu.log
timeout 200 python tools/stretch_bench.py > gpurun_out/r03b_stretch.json 2> gpurun_out/r03b_stretch.err
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r03b_bench.json 2> gpurun_out/r03b_bench.err
tail -4 gpurun_out/r03b_pytest_gpu.log; cat gpurun_out/r03b_stretch.json; tail -2 gpurun_out/r03b_stretch.err; python -c "
import json;d=json.load(open('gpurun_out/r03b_bench.json'));print(d['value'],d['ms_per_step'],{k:round(v['ms'],3) for k,v in d['kernels'].items()},d['e2e']['value'])"
