#!/bin/sh
# Build libllsm2_b200.so in-tree for B200 (sm_100a). nvcc cross-compiles without a GPU.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
"$NVCC" -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a \
  -Xcompiler -fPIC,-ffp-contract=off,-O2 -Xptxas -v \
  -shared -o "$HERE/libllsm2_b200.so" \
  "$HERE/csrc/api.cu" "$HERE/csrc/plan.cpp" "$HERE/csrc/compat.c" -lcudart -ldl 2> "$HERE/ptxas.log" || { cat "$HERE/ptxas.log"; exit 1; }
grep -E "error|warning" "$HERE/ptxas.log" | grep -v "ptxas info" || true
