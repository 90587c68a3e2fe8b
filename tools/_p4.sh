#!/bin/bash
# 8 GPUs: probe of the peer-memory optimizer step (p2p and multicast) vs NCCL, then bench.py: NCCL baseline, peer (prefetch at start / reduce), peer + multicast.
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tools/probe_peer.py > gpurun_out/probe_n$N.log 2>&1
grep "PROBE\|Error\|Traceback" gpurun_out/probe_n$N.log | head -5
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 300 $TR --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n${N}_$tag.json 2> gpurun_out/bench_n${N}_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_n${N}_$tag.json").read().strip().splitlines()[-1])
    print("N=$N $tag", round(d["value"]), round(d["ms_per_step"], 2), d["phases_ms"]["rollout_32_steps"], d["phases_ms"]["post_rollout"], d["phases_ms"]["update"], "e2e", round(d["e2e"]["value"]), d["optimizer_step"][:30])
except Exception as e:
    print("N=$N $tag FAILED", e); print(open("gpurun_out/bench_n${N}_$tag.err").read()[-2000:])
PY
}
run nccl PULSE_PEER_ADAM=0
run peer_start PULSE_PEER_ADAM=1
run peer_reduce PULSE_PEER_ADAM=1 PULSE_PREFETCH_AT=reduce
run peer_mc PULSE_PEER_ADAM=1 PULSE_PEER_MC=1
