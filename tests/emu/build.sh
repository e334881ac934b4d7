#!/bin/sh
# Build the CPU thread emulation of the kernels (test infrastructure only).
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
g++ -std=c++17 -O2 -g -ffp-contract=off -fPIC -shared -DLLSM_EMU -x c++ \
  -I"$HERE" -I"$ROOT/libllsm2_b200/csrc" \
  "$HERE/cuda_emu.cpp" "$HERE/emu_api.cpp" "$ROOT/libllsm2_b200/csrc/plan.cpp" \
  -o "$HERE/libllsm2_emu.so" -lpthread
