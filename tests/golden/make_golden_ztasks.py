"""Fixture for the downstream Z tasks HumanoidSpeedZ / HumanoidStrikeZ: outputs of the UNMODIFIED reference's
compute_speed_observations / compute_speed_reward (humanoid_speed.py:310-343), compute_strike_observations / compute_strike_reward and
the strike compute_humanoid_reset (humanoid_strike.py:270-375), compute_humanoid_reset and compute_humanoid_observations_smpl_max
(humanoid.py:1573-1608, :1675-1731) on seeded inputs.

  python tests/golden/make_golden_ztasks.py     (needs /root/reference; writes tests/golden/ztasks.npz)
"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

CONTACT_IDS = [7, 3, 8, 4]        # R_Ankle, L_Ankle, R_Toe, L_Toe
STRIKE_IDS = [22, 23]             # R_Wrist, R_Hand (humanoid_strike.py strikeBodyNames)
DT = 1.0 / 30.0
MAX_LEN = 300


def inputs(N, seed=0):
    g = torch.Generator().manual_seed(seed)
    unit = lambda q: q / q.norm(dim=-1, keepdim=True)
    bs = torch.zeros(N, 24, 13)
    bs[..., 0:3] = torch.randn(N, 24, 3, generator=g) * 0.4 + torch.tensor([0.0, 0.0, 0.9])
    bs[::5, 9:, 2] = 0.05                                   # some envs with low bodies (fall by height)
    bs[..., 3:7] = unit(torch.randn(N, 24, 4, generator=g))
    bs[..., 7:13] = torch.randn(N, 24, 6, generator=g)
    prev_root = bs[:, 0, 0:3] - torch.randn(N, 3, generator=g) * 0.04
    tar_speed = torch.rand(N, generator=g) * 4 + 0.5
    target = torch.zeros(N, 13)
    target[:, 0:3] = bs[:, 0, 0:3] + torch.randn(N, 3, generator=g) * 2
    target[:, 2] = 0.9
    target[:, 3:7] = unit(torch.randn(N, 4, generator=g) * torch.tensor([0.3, 0.3, 1.0, 1.0]))
    target[::3, 3:7] = unit(torch.tensor([[0.7, 0.0, 0.0, 0.7]]))    # toppled targets: success branch
    target[:, 7:13] = torch.randn(N, 6, generator=g)
    contact = torch.zeros(N, 24, 3)
    contact[::2] = torch.randn((N + 1) // 2, 24, 3, generator=g) * (torch.rand((N + 1) // 2, 24, 1, generator=g) < 0.15) * 60
    tar_contact = torch.randn(N, 3, generator=g) * 60
    progress = torch.randint(0, 310, (N,), generator=g)
    progress[::11] = 1
    term_h = torch.full((24,), 0.15)
    dof_force, dof_vel = torch.randn(N, 69, generator=g) * 30, torch.randn(N, 69, generator=g)
    return dict(body_state=bs, prev_root_pos=prev_root, tar_speed=tar_speed, target_states=target, contact_forces=contact,
                tar_contact_forces=tar_contact, progress_buf=progress, termination_heights=term_h, dof_force=dof_force, dof_vel=dof_vel)


def main():
    from oracle.refshim.load_reference import load_reference
    ref = load_reference()
    speed = importlib.import_module("env.tasks.humanoid_speed")
    strike = importlib.import_module("env.tasks.humanoid_strike")
    N = 203
    z = inputs(N)
    bs, root = z["body_state"], z["body_state"][:, 0]
    empty = torch.zeros(N, 0)
    out = {"num_envs": np.int64(N)}
    out["self_obs"] = ref.humanoid.compute_humanoid_observations_smpl_max(bs[..., 0:3], bs[..., 3:7], bs[..., 7:10], bs[..., 10:13], empty, empty,
                                                                          True, True, True, False, False)
    out["speed_obs"] = speed.compute_speed_observations(root, z["tar_speed"])
    out["speed_reward"] = speed.compute_speed_reward(root[:, 0:3], z["prev_root_pos"], root[:, 3:7], z["tar_speed"], DT)
    rs, tm = ref.humanoid.compute_humanoid_reset(torch.zeros(N, dtype=torch.long), z["progress_buf"], z["contact_forces"], torch.tensor(CONTACT_IDS),
                                                 bs[..., 0:3], MAX_LEN, True, z["termination_heights"])
    out["speed_reset"], out["speed_terminate"] = rs, tm
    out["strike_obs"] = strike.compute_strike_observations(root, z["target_states"])
    out["strike_reward"] = strike.compute_strike_reward(z["target_states"][:, 0:3], z["target_states"][:, 3:7], root, z["prev_root_pos"],
                                                        bs[:, STRIKE_IDS[0], 7:10], DT, 1.5)
    rs, tm = strike.compute_humanoid_reset(torch.zeros(N, dtype=torch.long), z["progress_buf"], z["contact_forces"], torch.tensor(CONTACT_IDS),
                                           bs[..., 0:3], z["tar_contact_forces"], torch.tensor(STRIKE_IDS), MAX_LEN, True, z["termination_heights"])
    out["strike_reset"], out["strike_terminate"] = rs, tm
    np.savez_compressed(os.path.join(HERE, "ztasks.npz"), **{k: (v.numpy() if torch.is_tensor(v) else v) for k, v in out.items()})
    print({k: getattr(v, "shape", None) for k, v in out.items()}, "terminated speed/strike:", int(out["speed_terminate"].sum()), int(out["strike_terminate"].sum()),
          "strike successes:", int((out["strike_reward"] == 1).sum()))


if __name__ == "__main__":
    main()
