#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_boundary.py tests/test_gpu_reset.py -x -q 2>&1 | tail -8
for gr in 0 1; do
  PULSE_GROUPED=$gr timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_gr$gr.json 2> gpurun_out/bench_gr$gr.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_gr$gr.json").read().strip().splitlines()[-1])
print("grouped $gr", round(d["value"]), round(d["ms_per_step"], 2), d["phases_ms"]["rollout_32_steps"], d["phases_ms"]["update"], d["e2e"]["value"])
PY
done
