"""One env step of the rollout (PlayStepsB200 segment + fused step kernel) for profiling at a given env count.

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/profile_rollout.py --envs 2048
"""
import argparse
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pulse_b200.humanoid_im import HumanoidImCompute  # noqa: E402
from pulse_b200.motion_lib import MotionLibB200  # noqa: E402
from pulse_b200.ppo import PPOPolicy  # noqa: E402
from pulse_b200.rollout import PlayStepsB200  # noqa: E402
from tools.synth import device_step_inputs, device_tables  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=2048)
    ap.add_argument("--time", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    n = a.envs
    ml = MotionLibB200.from_tables(device_tables(n, dev, seed=100))
    z = device_step_inputs(ml, n, seed=200)
    comp = HumanoidImCompute(ml)
    pol = PPOPolicy(device=dev, seed=0, with_disc=True)
    root = torch.zeros(n, 1, 13, device=dev)
    sim = dict(body_state=z["body_state"], root_states=root[:, 0], dof_pos=z["dof_pos"], dof_vel=z["dof_vel"], dof_force=z["dof_force"],
               progress_buf=z["progress_buf"], motion_ids=z["motion_ids"], motion_start_times=z["motion_start_times"],
               motion_start_offset=z["motion_start_offset"], global_offset=z["global_offset"], cycle_counter=z["cycle_counter"])
    ps = PlayStepsB200(comp, pol, sim, horizon=32, use_graphs=a.time, single_graph=True)
    ps.first_observation()
    for _ in range(3):
        ps.play_steps()
    torch.cuda.synchronize()
    if a.time:
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            ps.play_steps()
        e.record()
        torch.cuda.synchronize()
        print(f"{n} envs: {s.elapsed_time(e) / 10:.3f} ms per 32-step horizon ({s.elapsed_time(e) / 320 * 1e3:.1f} us per step)")
        return
    rt = ctypes.CDLL("libcudart.so")
    rt.cudaProfilerStart()
    ps._segment(5)
    ps._env_step(5)
    torch.cuda.synchronize()
    rt.cudaProfilerStop()


if __name__ == "__main__":
    main()
