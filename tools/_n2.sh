for mode in single chain; do
for ch in "" 4; do
  export PULSE_GRAD_REDUCE=$mode
  if [ -n "$ch" ]; then export NCCL_MAX_NCHANNELS=$ch; else unset NCCL_MAX_NCHANNELS; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$mode', 'nch=$ch', round(d['value']), round(d['ms_per_step'],1), {k:(round(v,2) if isinstance(v,float) else v) for k,v in d['phases_ms'].items() if k!='note'})"
done; done
