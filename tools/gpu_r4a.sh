#!/bin/bash
# Round-2 late visit: full GPU suite on the default build, analysis parity with the tensor-core residual, bench A/B.
TAG=${1:-r4a}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/${TAG}_pytest_all.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_all.log; tail -4 gpurun_out/${TAG}_pytest_all.log
LLSM_RESIDUAL_TC=1 timeout 600 python -m pytest tests/test_gpu_analysis.py tests/test_gpu_speech.py tests/test_gpu_kat.py tests/test_gpu_compat.py tests/test_gpu_variants.py -q -m gpu > gpurun_out/${TAG}_pytest_restc.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_restc.log; tail -15 gpurun_out/${TAG}_pytest_restc.log
tools/gpu_kern.sh ${TAG} none "LLSM_NOP=1" "LLSM_NS_HANN_TABLE=0" "LLSM_RESIDUAL_TC=1"
