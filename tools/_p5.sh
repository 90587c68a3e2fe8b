#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_eval.py tests/test_gpu_task_obs.py tests/test_gpu_boundary.py -q 2>&1 | tail -40
