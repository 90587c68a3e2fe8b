#!/bin/bash
# ncu source-level captures of two GEMM instances (one launch each)
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 8 -c 1 -f -o gpurun_out/gemm_fwd1_v7 python tools/bench_update_gemms.py --only actor.fwd1 > gpurun_out/ncu_g1.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 8 -c 1 -f -o gpurun_out/gemm_silu_dgrad_v7 python tools/bench_update_gemms.py --vae --only enc.dgrad1 > gpurun_out/ncu_g2.log 2>&1
tail -2 gpurun_out/ncu_g1.log gpurun_out/ncu_g2.log
