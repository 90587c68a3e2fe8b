#!/bin/bash
mkdir -p gpurun_out
run() {
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/n2_$1.err | tail -1 > gpurun_out/n2_$1.json
  python - <<PY
import json
d = json.loads(open("gpurun_out/n2_$1.json").read().strip().splitlines()[-1])
print("$1", round(d["value"]), round(d["ms_per_step"], 2), d["phases_ms"])
PY
}
PULSE_PREFETCH=0 run off
PULSE_PREFETCH_AT=reduce run reduce
PULSE_PREFETCH_AT=start run start
PULSE_PREFETCH_AT=loss run loss
