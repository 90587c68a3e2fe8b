#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rollout.py -x -q 2>&1 | tail -15
for envs in 2048 16384; do for ov in 1 0; do
  PULSE_ROLLOUT_OVERLAP=$ov timeout 300 python bench.py --envs $envs --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ov${ov}_$envs.json 2> gpurun_out/bench_ov${ov}_$envs.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_ov${ov}_$envs.json").read().strip().splitlines()[-1])
    print("envs $envs overlap $ov", round(d["value"]), round(d["ms_per_step"], 2), d["phases_ms"]["rollout_32_steps"], d["phases_ms"]["update"], "e2e", round(d["e2e"]["value"]), "step frac", round(d["roofline"]["frac"], 3))
except Exception as e:
    print("envs $envs overlap $ov FAILED", e); print(open("gpurun_out/bench_ov${ov}_$envs.err").read()[-1500:])
PY
done; done
