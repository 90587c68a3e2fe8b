"""Kernel micro-benchmarks (CUDA events, L2 flushed between iterations). Development tool; bench.py is the contract."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pulse_b200.humanoid_im import HumanoidImCompute  # noqa: E402
from pulse_b200.motion_lib import MotionLibB200  # noqa: E402
from pulse_b200.rollout import discount_values  # noqa: E402
from tools.synth import device_step_inputs, device_tables  # noqa: E402


def timeit(fn, iters=20, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.fill_(1.0)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=16384)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--median-frames", type=int, default=150)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    tabs = device_tables(a.envs, dev, median_frames=a.median_frames)
    ml = MotionLibB200.from_tables(tabs)
    F = tabs["gts"].shape[0]
    z = device_step_inputs(ml, a.envs)
    comp = HumanoidImCompute(ml)
    n = a.envs
    out = dict(obs_buf=torch.zeros(n, 934, device=dev), rew_buf=torch.zeros(n, device=dev), reward_raw=torch.zeros(n, 5, device=dev),
               reset_buf=torch.zeros(n, dtype=torch.long, device=dev), terminate_buf=torch.zeros(n, dtype=torch.long, device=dev))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    kw = dict(body_state=z["body_state"], dof_vel=z["dof_vel"], dof_force=z["dof_force"], progress_buf=z["progress_buf"],
              motion_ids=z["motion_ids"], motion_start_times=z["motion_start_times"], motion_start_offset=z["motion_start_offset"],
              global_offset=z["global_offset"], cycle_counter=z["cycle_counter"])
    res = {"envs": n, "frames": F, "table_GB": F * 312 * 4 / 1e9}
    med, best = timeit(lambda: comp.step(**kw, **out), a.iters, flush=flush)
    res["im_step_ms"] = med
    res["im_step_ms_best"] = best
    res["im_step_GBs_algorithmic_9396"] = 9396 * n / (med * 1e-3) / 1e9
    med_w, _ = timeit(lambda: comp.step(**kw, **out), a.iters, flush=None)
    res["im_step_ms_warmL2"] = med_w
    amp = torch.zeros(n, 10, 196, device=dev)
    med, _ = timeit(lambda: comp.amp_obs(body_state=z["body_state"], dof_pos=z["dof_pos"], dof_vel=z["dof_vel"], amp_obs_buf=amp), a.iters, flush=flush)
    res["amp_obs_ms"] = med
    res["amp_obs_GBs"] = n * (2 * 9 * 196 * 4 + 196 * 4 + 69 * 8 + 7 * 52) / (med * 1e-3) / 1e9
    T = 32
    r, v, nv = (torch.randn(T, n, 1, device=dev) for _ in range(3))
    d = (torch.rand(T, n, device=dev) < 0.05).float()
    med, _ = timeit(lambda: discount_values(d, v, r, nv, normalize_advantage=True), a.iters, flush=flush)
    res["gae_ms"] = med
    res["gae_GBs"] = T * n * 24 / (med * 1e-3) / 1e9
    print(json.dumps(res))


if __name__ == "__main__":
    main()
