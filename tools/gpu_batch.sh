#!/bin/bash
# One gpurun call: GPU tests, secondary workloads, an ncu capture of the step kernel, the headline bench.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
python tools/bench_pulse.py --workload vae --json gpurun_out/bench_vae.json > gpurun_out/bench_vae.log 2>&1
python tools/bench_pulse.py --workload reach --json gpurun_out/bench_reach.json > gpurun_out/bench_reach.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:im_step -s 5 -c 1 -f -o gpurun_out/prof_im_step_v5 python tools/microbench.py --iters 3 > gpurun_out/ncu_im_step.log 2>&1
python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/pytest_gpu.log; tail -c 1500 gpurun_out/bench_vae.log; echo; tail -c 1200 gpurun_out/bench_reach.log; echo; tail -2 gpurun_out/ncu_im_step.log; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k: d[k] for k in ('value','ms_per_step','e2e','gpu_launches','roofline','roofline_update')})"
