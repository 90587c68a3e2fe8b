#!/bin/bash
# Development tool: build libpulse_b200 with -DPULSE_GEMM_VARIANT=$1 into pulse_b200/build/libpulse_v$1.so (A/B on one GPU box:
# PULSE_ALT_LIB=pulse_b200/build/libpulse_v1.so python tools/bench_update_gemms.py).
set -e
cd "$(dirname "$0")/.."
V=$1
python -m pulse_b200.build > /dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Iinclude -Ipulse_b200/csrc -DPULSE_GEMM_VARIANT=$V \
  -c pulse_b200/csrc/gemm_tcgen05.cu -o pulse_b200/build/gemm_v$V.o
OBJS=$(ls pulse_b200/build/*.o | grep -v "gemm_tcgen05.o\|gemm_v")
nvcc -gencode arch=compute_100a,code=sm_100a --shared -o pulse_b200/build/libpulse_v$V.so $OBJS pulse_b200/build/gemm_v$V.o -lcudart
echo pulse_b200/build/libpulse_v$V.so
