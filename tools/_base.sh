#!/bin/bash
# Baseline check of the restored tree: GPU tests, smoke, headline bench.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -25 ) > gpurun_out/pytest_gpu.log 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench.json | cut -c1-3000
