"""4096-env golden of one HumanoidIm post-physics step (BASELINE config C2 size: 4096 envs, 100 clips, env -> clip =
arange % 100) produced by the UNMODIFIED reference functions (build container only):

  python tests/golden/make_golden_step4096.py      -> tests/golden/step_n4096.npz

Inputs are regenerated bit for bit on every host by tests.helpers.exact_tables / exact_step_inputs (checksum stored); the
MotionLib tables are handed to an un-initialised `MotionLibSMPL` instance, so `get_motion_state` / `_calc_frame_blend`
(phc/utils/motion_lib_base.py:434-517, :546-556) are the reference's own code, followed by the same call sequence as
make_golden.py: `compute_imitation_reward` + power term (humanoid_im.py:853-919, :1543-1574), `compute_humanoid_im_reset`
(+ the recovery override, :1119-1192, :1600-1628), `compute_humanoid_observations_smpl_max` (humanoid.py:1675-1731),
`compute_imitation_observations_v6` (humanoid_im.py:1328-1378).  Stored: all per-env scalars, float64 row sums of the 934
observation columns for every env, the full observation of every 32nd env.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))
from load_reference import load_reference  # noqa: E402
from tests.helpers import exact_step_inputs, exact_tables  # noqa: E402

RESET_BODY_IDS = [j for j in range(24) if j not in (3, 4, 7, 8)]
N, CLIPS = 4096, 100


def reference_step(ref, tb, z, dt):
    from phc.utils.motion_lib_smpl import MotionLibSMPL
    him, hum = ref.humanoid_im, ref.humanoid
    lib = MotionLibSMPL.__new__(MotionLibSMPL)
    for k in ("gts", "grs", "lrs", "gvs", "gavs", "dvs"):
        setattr(lib, k, getattr(tb, k))
    lib._motion_aa, lib._motion_lengths, lib._motion_num_frames, lib._motion_dt = tb.motion_aa, tb.lengths, tb.num_frames, tb.dt
    lib.length_starts, lib._motion_bodies, lib._motion_limb_weights = tb.length_starts, tb.motion_bodies, tb.motion_limb_weights
    lib.num_bodies = 24                          # set by load_motions (motion_lib_base.py:318)
    n = z["motion_ids"].shape[0]
    ids, progress, goff = z["motion_ids"], z["progress_buf"], z["global_offset"]
    bs = z["body_state"]
    body_pos, body_rot, body_vel, body_ang = bs[..., 0:3], bs[..., 3:7], bs[..., 7:10], bs[..., 10:13]
    t_rew = progress * dt + z["start_times"] + z["start_offset"]
    pose = lib.get_motion_state(ids, t_rew, offset=goff)
    specs = {"k_pos": 100.0, "k_rot": 10.0, "k_vel": 0.1, "k_ang_vel": 0.1, "w_pos": 0.5, "w_rot": 0.3, "w_vel": 0.1, "w_ang_vel": 0.1}
    rew, raw = him.compute_imitation_reward(body_pos[:, 0], body_rot[:, 0], body_pos, body_rot, body_vel, body_ang,
                                            pose["rg_pos"], pose["rb_rot"], pose["body_vel"], pose["body_ang_vel"], specs)
    power_reward = -0.0005 * torch.abs(torch.multiply(z["dof_force"], z["dof_vel"])).sum(dim=-1)
    power_reward[progress <= 3] = 0
    rew = rew + power_reward
    raw = torch.cat([raw, power_reward[:, None]], dim=-1)
    pass_time = t_rew >= tb.lengths[ids]
    rb = torch.tensor(RESET_BODY_IDS)
    term = torch.full((1, 24), 0.25)[..., rb]
    reset, terminated = him.compute_humanoid_im_reset(z["reset_buf_in"], progress, torch.zeros(n, 24, 3), torch.zeros(4, dtype=torch.long),
                                                      body_pos[..., rb, :].clone(), pose["rg_pos"][..., rb, :].clone(), pass_time, True, term,
                                                      False, False)
    rec = torch.logical_and(~pass_time, z["cycle_counter"] > 0)
    reset, terminated = reset.clone(), terminated.clone()
    reset[rec] = 0
    terminated[rec] = 0
    t_obs = (progress + 1) * dt + z["start_times"] + z["start_offset"]
    nxt = lib.get_motion_state(ids, t_obs, offset=goff)
    empty = torch.zeros(n, 0)
    self_obs = hum.compute_humanoid_observations_smpl_max(body_pos, body_rot, body_vel, body_ang, empty, empty, True, True, True, False, False)
    task_obs = him.compute_imitation_observations_v6(body_pos[:, 0], body_rot[:, 0], body_pos, body_rot, body_vel, body_ang,
                                                     nxt["rg_pos"], nxt["rb_rot"], nxt["body_vel"], nxt["body_ang_vel"], 1, True)
    L, nfr, mdt = tb.lengths[ids], tb.num_frames[ids], tb.dt[ids]
    i0r, i1r, _ = lib._calc_frame_blend(t_rew, L, nfr, mdt)
    i0o, i1o, _ = lib._calc_frame_blend(t_obs, L, nfr, mdt)
    return {"rew_buf": rew, "reward_raw": raw, "reset_buf": reset, "terminate_buf": terminated, "obs_buf": torch.cat([self_obs, task_obs], -1),
            "frame_idx_rew": torch.stack([i0r, i1r], -1), "frame_idx_obs": torch.stack([i0o, i1o], -1)}


def main():
    cwd = os.getcwd()
    os.chdir("/tmp")
    ref = load_reference()
    tb = exact_tables(CLIPS)
    z, chk = exact_step_inputs(tb, N)
    dt = float(torch.tensor(1.0 / 60.0) * 2)
    out = reference_step(ref, tb, z, dt)
    obs = out.pop("obs_buf")
    d = {k: v.numpy() for k, v in out.items()}
    d.update(checksum=np.float64(chk), obs_row_sum=obs.double().sum(1).numpy(), obs_col_sum=obs.double().sum(0).numpy(),
             obs_rows=obs[::32].numpy(), dims=np.array([N, CLIPS]))
    path = os.path.join(HERE, "step_n4096.npz")
    np.savez_compressed(path, **d)
    os.chdir(cwd)
    print("step_n4096.npz", os.path.getsize(path) // 1024, "KiB; checksum", chk, "terminated", int(out["terminate_buf"].sum()), "reset",
          int(out["reset_buf"].sum()), "mean reward", float(out["rew_buf"].mean()))


if __name__ == "__main__":
    main()
