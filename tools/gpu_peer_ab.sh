#!/bin/bash
# N GPUs (gpurun --gpus N): real peer-memory optimizer step vs NCCL (tools/probe_peer.py), then bench.py with the peer step on / off.
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
NCCL_DEBUG=INFO timeout 300 $TR --master-port 29511 tools/probe_peer.py > gpurun_out/probe_n$N.log 2>&1
grep -c "NVLS" gpurun_out/probe_n$N.log | sed 's/^/NVLS lines: /'
grep "PROBE\|Error\|error\|Traceback\|unavailable" gpurun_out/probe_n$N.log | head -20
for pa in 1 0; do
  PULSE_PEER_ADAM=$pa timeout 400 $TR --master-port 2952$pa bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n${N}_peer$pa.json 2> gpurun_out/bench_n${N}_peer$pa.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_n${N}_peer$pa.json").read().strip().splitlines()[-1])
    print("N=$N peer=$pa", round(d["value"]), round(d["ms_per_step"], 2), d["phases_ms"]["rollout_32_steps"], d["phases_ms"]["update"], "e2e", round(d["e2e"]["value"]), d["optimizer_step"][:40])
except Exception as e:
    print("N=$N peer=$pa FAILED", e); print(open("gpurun_out/bench_n${N}_peer$pa.err").read()[-2500:])
PY
done
