#!/bin/bash
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -8
python tools/bench_update_gemms.py --json gpurun_out/gemms_v11.json | grep "dgrad\|sum"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print('PPO', d['value'], d['ms_per_step'], d['roofline_update']['update_ms'], d['roofline_update']['frac'], d['roofline']['frac'])"
