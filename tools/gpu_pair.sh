#!/bin/bash
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -6
python tools/bench_update_gemms.py --json gpurun_out/gemms_v10.json | grep "fwd\|sum"
python tools/bench_update_gemms.py --vae --json gpurun_out/gemms_vae_v10.json | grep "fwd\|sum"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print('PPO', d['value'], d['ms_per_step'], d['roofline_update']['update_ms'], d['roofline_update']['frac'], d['roofline']['frac'])"
python bench.py --envs 2048 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_e2048.json 2> gpurun_out/bench_e2048.err; python -c "
import json; d=json.load(open('gpurun_out/bench_e2048.json')); print('PPO 2048 envs (per-rank work of N=8)', d['value'], d['ms_per_step'], d['roofline_update']['update_ms'], d['roofline']['avg_launch_ms'])"
