"""SURVEY 8f-4, downstream Z tasks: the oracle's restatements of the speed / strike observation, reward and reset functions against
the fixture written by the UNMODIFIED reference (tests/golden/make_golden_ztasks.py)."""
import importlib.util
import os

import numpy as np
import torch

from oracle import pulse_oracle as po

HERE = os.path.dirname(os.path.abspath(__file__))


def gen():
    spec = importlib.util.spec_from_file_location("make_golden_ztasks", os.path.join(HERE, "golden", "make_golden_ztasks.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_oracle_matches_reference_fixture():
    m = gen()
    g = np.load(os.path.join(HERE, "golden", "ztasks.npz"))
    z = m.inputs(int(g["num_envs"]))
    bs, root = z["body_state"], z["body_state"][:, 0]
    close = lambda a, name: torch.testing.assert_close(a, torch.from_numpy(g[name]), atol=2e-6, rtol=2e-6, msg=lambda s: f"{name}: {s}")
    close(po.self_obs_smpl_max(bs[..., 0:3], bs[..., 3:7], bs[..., 7:10], bs[..., 10:13]), "self_obs")
    close(po.speed_obs(root, z["tar_speed"]), "speed_obs")
    close(po.speed_reward(root[:, 0:3], z["prev_root_pos"], z["tar_speed"], m.DT), "speed_reward")
    rs, tm = po.humanoid_reset(z["progress_buf"], z["contact_forces"], torch.tensor(m.CONTACT_IDS), bs[..., 0:3], m.MAX_LEN, True, z["termination_heights"])
    assert torch.equal(rs, torch.from_numpy(g["speed_reset"])) and torch.equal(tm, torch.from_numpy(g["speed_terminate"]))
    close(po.strike_obs(root, z["target_states"]), "strike_obs")
    close(po.strike_reward(z["target_states"][:, 0:3], z["target_states"][:, 3:7], root[:, 0:3], z["prev_root_pos"], m.DT), "strike_reward")
    rs, tm = po.strike_reset(z["progress_buf"], z["contact_forces"], torch.tensor(m.CONTACT_IDS), bs[..., 0:3], z["tar_contact_forces"],
                             torch.tensor(m.STRIKE_IDS), m.MAX_LEN, True, z["termination_heights"])
    assert torch.equal(rs, torch.from_numpy(g["strike_reset"])) and torch.equal(tm, torch.from_numpy(g["strike_terminate"]))
