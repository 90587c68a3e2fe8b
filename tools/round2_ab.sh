#!/bin/bash
# One gpurun call that validates and measures everything round 1 left prepared but unvalidated (run
# tools/round2_build_variants.sh first):
#   1. opt-in GPU tests: grouped GEMM launches + lock-step PPO update, MotionDatasetB200.load_motions, reset_ref_state
#   2. step-kernel A/B: consumer teams / stages (isolated microbench, L2 flushed)
#   3. headline bench with grouped launches off / on
mkdir -p gpurun_out
echo "== opt-in tests"
PULSE_GROUPED_TEST=1 PULSE_EXPERIMENTAL_DATASET=1 PULSE_EXPERIMENTAL_RESET=1 timeout 300 python -m pytest tests/test_gpu_grouped.py tests/test_gpu_loader.py tests/test_gpu_reset.py -q 2>&1 | tail -15
echo "== step kernel variants (im_step_ms / GB/s at 16384 envs)"
for v in product t2 t2s4 t3s4; do
  if [ $v = product ]; then unset PULSE_ALT_LIB; else export PULSE_ALT_LIB=$PWD/pulse_b200/build/libpulse_$v.so; fi
  [ $v != product ] && [ ! -f "$PULSE_ALT_LIB" ] && { echo "$v: not built"; continue; }
  python - <<PY
import json, os, subprocess, sys
from pulse_b200 import _lib
if os.environ.get("PULSE_ALT_LIB"): _lib.LIB_PATH = os.environ["PULSE_ALT_LIB"]
sys.argv = ["microbench", "--iters", "20"]
import io, contextlib
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    import runpy; runpy.run_path("tools/microbench.py", run_name="__main__")
d = json.loads(buf.getvalue().strip().splitlines()[-1])
print("$v", round(d["im_step_ms"] * 1e3, 1), "us", round(d["im_step_GBs_algorithmic_9396"]), "GB/s  best", round(d["im_step_ms_best"] * 1e3, 1), "us")
PY
done
unset PULSE_ALT_LIB
echo "== headline bench, grouped launches off / on"
for g in 0 1; do
  PULSE_GROUPED=$g python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_grouped$g.json 2> gpurun_out/bench_grouped$g.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_grouped$g.json')); print('PULSE_GROUPED=$g', round(d['value']), 'env-steps/s', round(d['ms_per_step'],1), 'ms  update', round(d['roofline_update']['update_ms'],1), 'ms', d['gemm_switches'])" || tail -3 gpurun_out/bench_grouped$g.err
done
PULSE_GROUPED=1 python bench.py --envs 2048 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('2048 envs (per-rank work at N=8), grouped:', round(d['ms_per_step'],2), 'ms  (ungrouped round-1: 23.9 ms)')"
