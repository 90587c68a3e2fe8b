#!/bin/bash
python -m pytest tests/test_gpu_gemm.py tests/test_gpu_ppo.py tests/test_gpu_vae.py -q -x 2>&1 | tail -3
echo "=== pdl on"; python tools/bench_update_gemms.py --json gpurun_out/gemms_pdl.json | grep -v "gp\."
echo "=== pdl off"; PULSE_GEMM_PDL=0 python tools/bench_update_gemms.py --json gpurun_out/gemms_nopdl.json | grep -v "gp\."
echo "=== vae pdl on"; python tools/bench_update_gemms.py --vae --json gpurun_out/gemms_vae_pdl.json
python tools/bench_pulse.py --workload vae --json gpurun_out/bench_vae.json > gpurun_out/bench_vae.log 2>&1; python -c "
import json; d=json.load(open('gpurun_out/bench_vae.json')); print('VAE', d['value'], d['ms_per_iteration'], d['update_ms'], d['roofline_update']['frac'])"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print('PPO', d['value'], d['ms_per_step'], d['roofline_update']['update_ms'], d['roofline_update']['frac'], d['roofline']['frac'])"
PULSE_GEMM_PDL=0 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nopdl.json 2> gpurun_out/bench_nopdl.err; python -c "
import json; d=json.load(open('gpurun_out/bench_nopdl.json')); print('PPO nopdl', d['value'], d['ms_per_step'], d['roofline_update']['update_ms'], d['roofline_update']['frac'])"
