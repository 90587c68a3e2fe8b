#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ppo.py tests/test_gpu_boundary.py tests/test_gpu_rollout.py -x -q 2>&1 | tail -8
for m in none start loss reduce; do timeout 120 python tools/profile_update.py --time --prefetch $m 2>&1 | tail -1; done
for pf in 0 1; do
  PULSE_PREFETCH=$pf timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pf$pf.json 2> gpurun_out/bench_pf$pf.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_pf$pf.json").read().strip().splitlines()[-1])
print("prefetch $pf", round(d["value"]), round(d["ms_per_step"], 2), d["phases_ms"]["update"], d["e2e"]["value"])
PY
done
PULSE_PREFETCH_AT=start timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('prefetch at start', round(d['value']), d['phases_ms']['update'])"
