"""Synthetic AMASS-shaped MotionLib tables and simulator state generated directly ON THE DEVICE
(one loaded clip per env, lognormal clip lengths at 30 fps)."""
import math

import torch


def device_tables(num_motions, device, seed=0, median_frames=150, min_frames=5, max_frames=1800, sigma=0.7):
    g = torch.Generator(device=device).manual_seed(seed)
    nf = torch.exp(torch.randn(num_motions, generator=g, device=device) * sigma + math.log(median_frames)).long().clamp(min_frames, max_frames)
    F = int(nf.sum().item())
    starts = torch.cumsum(nf, 0) - nf
    unit = lambda x: torch.nn.functional.normalize(x, dim=-1)
    rep = lambda x: x.repeat_interleave(nf, dim=0)
    grs = unit(rep(unit(torch.randn(num_motions, 24, 4, generator=g, device=device))) + 0.15 * torch.randn(F, 24, 4, generator=g, device=device))
    lrs = unit(rep(unit(torch.randn(num_motions, 24, 4, generator=g, device=device))) + 0.15 * torch.randn(F, 24, 4, generator=g, device=device))
    gts = rep(torch.randn(num_motions, 24, 3, generator=g, device=device) * 0.4 + torch.tensor([0.0, 0.0, 0.9], device=device)) \
        + 0.05 * torch.randn(F, 24, 3, generator=g, device=device)
    fps = 30.0
    return {
        "gts": gts, "grs": grs, "lrs": lrs, "gvs": torch.randn(F, 24, 3, generator=g, device=device),
        "gavs": torch.randn(F, 24, 3, generator=g, device=device), "dvs": torch.randn(F, 23, 3, generator=g, device=device),
        "motion_aa": torch.zeros(F, 72, device=device),
        "lengths": ((nf - 1).double() * (1.0 / fps)).float(), "num_frames": nf, "dt": torch.full((num_motions,), 1.0 / fps, device=device),
        "length_starts": starts, "fps": torch.full((num_motions,), fps, device=device),
    }


def device_step_inputs(ml, n_envs, seed=1, bodies_per_env=24, dofs_per_env=69):
    """Task buffers + simulator state near the reference pose; Isaac-Gym shaped views."""
    dev = ml._device
    g = torch.Generator(device=dev).manual_seed(seed)
    motion_ids = torch.arange(n_envs, device=dev) % ml.num_motions()
    progress = torch.randint(0, 40, (n_envs,), generator=g, device=dev)
    start = ml.sample_time_interval(motion_ids, phase=torch.rand(n_envs, generator=g, device=dev))
    start_off = torch.zeros(n_envs, device=dev)
    goff = torch.zeros(n_envs, 3, device=dev)
    dt = float(torch.tensor(1.0 / 60.0) * 2)
    t = progress * dt + start + start_off
    pose = ml.get_motion_state(motion_ids, t, goff)
    body = torch.zeros(n_envs, bodies_per_env, 13, device=dev)
    body[:, :24, 0:3] = pose["rg_pos"] + 0.03 * torch.randn(n_envs, 24, 3, generator=g, device=dev)
    body[:, :24, 3:7] = torch.nn.functional.normalize(pose["rb_rot"] + 0.05 * torch.randn(n_envs, 24, 4, generator=g, device=dev), dim=-1)
    body[:, :24, 7:10] = pose["body_vel"] + 0.5 * torch.randn(n_envs, 24, 3, generator=g, device=dev)
    body[:, :24, 10:13] = pose["body_ang_vel"] + 0.5 * torch.randn(n_envs, 24, 3, generator=g, device=dev)
    dof_state = torch.zeros(n_envs, dofs_per_env, 2, device=dev)
    dof_state[:, :69, 0] = pose["dof_pos"] + 0.05 * torch.randn(n_envs, 69, generator=g, device=dev)
    dof_state[:, :69, 1] = pose["dof_vel"] + 0.5 * torch.randn(n_envs, 69, generator=g, device=dev)
    return {
        "motion_ids": motion_ids, "progress_buf": progress, "motion_start_times": start, "motion_start_offset": start_off,
        "global_offset": goff, "cycle_counter": torch.zeros(n_envs, dtype=torch.int32, device=dev), "body_state": body,
        "dof_state": dof_state, "dof_pos": dof_state[:, :69, 0], "dof_vel": dof_state[:, :69, 1],
        "dof_force": 30 * torch.randn(n_envs, 69, generator=g, device=dev),
    }
