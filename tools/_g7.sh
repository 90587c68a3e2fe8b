python -m pytest tests/test_gpu_step.py tests/test_gpu_reset.py tests/test_gpu_rollout.py -q 2>&1 | tail -3
python tools/microbench.py --iters 30 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:(round(v,4) if isinstance(v,float) else v) for k,v in d.items() if 'im_step' in k})"
python tools/microbench.py --iters 30 --envs 2048 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('2048:', {k:(round(v,4) if isinstance(v,float) else v) for k,v in d.items() if 'im_step' in k})"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:im_step_kernel -s 3 -c 1 -f -o gpurun_out/r02_im_step_v6 python tools/microbench.py --iters 3 > gpurun_out/ncu_step.log 2>&1; ls -la gpurun_out/r02_im_step_v6.ncu-rep
