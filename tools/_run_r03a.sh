mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r03a_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r03a_pytest_gpu.log
timeout 200 python tools/stretch_bench.py > gpurun_out/r03a_stretch.json 2> gpurun_out/r03a_stretch.err
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r03a_bench.json 2> gpurun_out/r03a_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:frames_stretch -s 2 -c 1 -o gpurun_out/r03a_stretch python tools/stretch_bench.py --no-cpu > gpurun_out/r03a_ncu.log 2>&1
tail -4 gpurun_out/r03a_pytest_gpu.log; cat gpurun_out/r03a_stretch.json; tail -2 gpurun_out/r03a_stretch.err; cat gpurun_out/r03a_bench.json
