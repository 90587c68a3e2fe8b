#!/bin/bash
# N GPUs: bench.py with all defaults (peer-memory optimizer step, overlapped rollout schedule)
N=${1:-4}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n${N}_final.json 2> gpurun_out/bench_n${N}_final.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_n${N}_final.json").read().strip().splitlines()[-1])
    print("N=$N", round(d["value"]), round(d["ms_per_step"], 2), d["phases_ms"]["rollout_32_steps"], d["phases_ms"]["post_rollout"], d["phases_ms"]["update"], "e2e", round(d["e2e"]["value"]), d["optimizer_step"][:40])
except Exception as e:
    print("N=$N FAILED", e); print(open("gpurun_out/bench_n${N}_final.err").read()[-2500:])
PY
