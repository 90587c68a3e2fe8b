python -m pytest tests/test_gpu_rollout.py -q 2>&1 | tail -8
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; tail -3 gpurun_out/bench_r2a.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2a.json'))
print('value', round(d['value']), 'ms', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value']), 'phases', {k:(round(v,2) if isinstance(v,float) else v) for k,v in d['phases_ms'].items() if k!='note'}, 'step frac', round(d['roofline']['frac'],3), 'k_ms', round(d['roofline']['avg_launch_ms'],4), 'upd frac', round(d['roofline_update']['frac'],3), 'resets', round(d['resets_per_env_step'],4), 'launches', d['gpu_launches'])
PY
PULSE_ROLLOUT_GRAPH=0 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('segment graphs:', round(d['value']), d['phases_ms']['rollout_32_steps'])"
python bench.py --envs 2048 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('2048 envs:', round(d['ms_per_step'],2), 'ms', {k:(round(v,2) if isinstance(v,float) else v) for k,v in d['phases_ms'].items() if k!='note'}, 'step frac', round(d['roofline']['frac'],3))"
