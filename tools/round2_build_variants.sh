#!/bin/bash
# Build (in the container, BEFORE gpurun: nvcc cross-compiles) the A/B libraries tools/round2_ab.sh measures.
set -e
cd "$(dirname "$0")/.."
tools/build_variant.sh t2 im_step.cu -DPULSE_STEP_TEAMS=2
tools/build_variant.sh t2s4 im_step.cu -DPULSE_STEP_TEAMS=2 -DPULSE_STEP_STAGES=4
tools/build_variant.sh t3s4 im_step.cu -DPULSE_STEP_STAGES=4
tools/build_variant.sh trace gemm_tcgen05.cu -DPULSE_GEMM_VARIANT=3
ls -la pulse_b200/build/libpulse_*.so
