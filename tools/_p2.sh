#!/bin/bash
# 1 GPU: emulated-rank tests of the peer optimizer step, rollout tests with the stream forks, rollout fork A/B at 16384 and 2048 envs.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_peer_adam.py tests/test_gpu_rollout.py tests/test_gpu_boundary.py -x -q 2>&1 | tail -15
for envs in 16384 2048; do for f in 1 0; do
  PULSE_ROLLOUT_FORK=$f timeout 300 python bench.py --envs $envs --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fork${f}_$envs.json 2> gpurun_out/bench_fork${f}_$envs.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_fork${f}_$envs.json").read().strip().splitlines()[-1])
    print("envs $envs fork $f", round(d["value"]), round(d["ms_per_step"], 2), d["phases_ms"]["rollout_32_steps"], d["phases_ms"]["update"], "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("envs $envs fork $f FAILED", e); print(open("gpurun_out/bench_fork${f}_$envs.err").read()[-1500:])
PY
done; done
