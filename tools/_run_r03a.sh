mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_analysis.py tests/test_gpu_kat.py tests/test_gpu_compat.py -x -q > gpurun_out/r03h_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r03h_pytest_gpu.log
tail -3 gpurun_out/r03h_pytest_gpu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r03h_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r03h_ncu.log 2>&1
grep -c harmonic_dft gpurun_out/r03h_launches.csv
