#!/bin/bash
# Round-end check on one B200 (tight GPU budget): full GPU suite, headline bench, ncu launch list of the bench.
mkdir -p gpurun_out
( time timeout 170 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -14 ) > gpurun_out/pytest_gpu.log 2>&1
cat gpurun_out/pytest_gpu.log
timeout 90 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
    print("PPO", round(d["value"]), round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]), d["phases_ms"]["rollout_32_steps"], d["phases_ms"]["post_rollout"], d["phases_ms"]["update"],
          "step", round(d["roofline"]["frac"], 3), d["roofline"]["avg_launch_ms"], "upd", round(d["roofline_update"]["frac"], 3), d["clocks"])
except Exception as e:
    print("bench FAILED", e, open("gpurun_out/bench.err").read()[-1500:])
PY
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r02_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-200; ls -la gpurun_out/launches_r02_final.csv
