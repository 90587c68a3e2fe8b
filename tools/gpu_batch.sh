#!/bin/bash
# One gpurun call: GPU tests, smoke, headline bench, secondary workloads, ncu launch list and full captures of the GEMM.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
python tools/bench_pulse.py --workload vae --json gpurun_out/bench_vae.json > gpurun_out/bench_vae.log 2>&1
python tools/bench_pulse.py --workload reach --json gpurun_out/bench_reach.json > gpurun_out/bench_reach.log 2>&1
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_v6.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 8 -c 1 -f -o gpurun_out/gemm_fwd1_v10 python tools/bench_update_gemms.py --only actor.fwd1 > gpurun_out/ncu_g1.log 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 8 -c 1 -f -o gpurun_out/gemm_wgrad1_v10 python tools/bench_update_gemms.py --only actor.wgrad1 > gpurun_out/ncu_g2.log 2>&1
cat gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; python -c "
import json
d=json.load(open('gpurun_out/bench.json')); print('PPO', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'upd', d['roofline_update']['frac'], 'step', d['roofline']['frac'], d['roofline']['traffic'], 'cpu', d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline',{}).get('cores'), d['clocks'])
for w in ('vae','reach'):
    d=json.load(open('gpurun_out/bench_%s.json'%w)); print(w, d['value'], d['ms_per_iteration'], d['update_ms'], d['roofline_update']['frac'])
"; tail -2 gpurun_out/ncu_bench.log | cut -c1-200; ls -la gpurun_out/*.ncu-rep gpurun_out/launches_v6.csv
