mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r03e_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r03e_pytest_gpu.log
tail -5 gpurun_out/r03e_pytest_gpu.log
