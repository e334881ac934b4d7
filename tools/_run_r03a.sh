mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_synth.py -x -q > gpurun_out/r03d_pytest.log 2>&1; tail -2 gpurun_out/r03d_pytest.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r03d_bench.json 2> gpurun_out/r03d_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r03d_bench.json'));print(d['value'],d['ms_per_step'],{k:round(v['ms'],3) for k,v in d['kernels'].items()})"
