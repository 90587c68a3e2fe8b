"""SURVEY 8f-4, downstream Z tasks on the device: pulse_ztask_step (SpeedTaskB200 / StrikeTaskB200) against the fixture written by the
unmodified reference: observations / rewards within 1e-4 (speed reward 2e-4: exp of a squared finite-difference velocity), reset and
terminate masks bit-exact; the speed task's power term against the oracle."""
import os

import numpy as np
import pytest
import torch

from tests.test_ztasks_cpu import gen

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
DEV = "cuda:0"


def _sim(z):
    """Isaac-Gym-shaped views: 26 bodies per env (24 + the target's body for strike), 2 actors per env."""
    n = z["body_state"].shape[0]
    rb = torch.full((n, 26, 13), 5.0, device=DEV)
    rb[:, :24] = z["body_state"].to(DEV)
    cf = torch.zeros(n, 26, 3, device=DEV)
    cf[:, :24] = z["contact_forces"].to(DEV)
    cf[:, 24] = z["tar_contact_forces"].to(DEV)
    roots = torch.zeros(n, 2, 13, device=DEV)
    roots[:, 0], roots[:, 1] = z["body_state"][:, 0].to(DEV), z["target_states"].to(DEV)
    return rb, cf, roots


@pytest.mark.parametrize("power", [False, True])
def test_speed_task_step(power):
    from oracle import pulse_oracle as po
    from pulse_b200.ztasks import SpeedTaskB200
    m = gen()
    g = np.load(os.path.join(HERE, "golden", "ztasks.npz"))
    z = m.inputs(int(g["num_envs"]))
    n = z["body_state"].shape[0]
    rb, cf, _ = _sim(z)
    task = SpeedTaskB200(n, DEV, max_episode_length=m.MAX_LEN, dt=m.DT, power_reward=power)
    task._prev_root_pos.copy_(z["prev_root_pos"].to(DEV))
    task._tar_speed.copy_(z["tar_speed"].to(DEV))
    dof_state = torch.zeros(n, 69, 2, device=DEV)
    dof_state[:, :, 1] = z["dof_vel"].to(DEV)
    task.post_physics_step(rb, z["progress_buf"].to(DEV), cf, dof_force=z["dof_force"].to(DEV), dof_vel=dof_state[:, :, 1])
    torch.cuda.synchronize()
    T = lambda k: torch.from_numpy(g[k])
    torch.testing.assert_close(task.obs_buf[:, :358].cpu(), T("self_obs"), atol=1e-4, rtol=0)
    torch.testing.assert_close(task.obs_buf[:, 358:].cpu(), T("speed_obs"), atol=1e-4, rtol=0)
    want = T("speed_reward")
    torch.testing.assert_close(task.reward_raw[:, 0].cpu(), want, atol=2e-4, rtol=0)
    if power:
        pw = po.power_reward(z["dof_force"], z["dof_vel"], z["progress_buf"], 0.0005)
        torch.testing.assert_close(task.reward_raw[:, 1].cpu(), pw, atol=1e-5, rtol=1e-5)
        want = want + pw
    torch.testing.assert_close(task.rew_buf.cpu(), want, atol=2e-4, rtol=0)
    assert torch.equal(task.reset_buf.cpu(), T("speed_reset")) and torch.equal(task._terminate_buf.cpu(), T("speed_terminate"))


def test_strike_task_step():
    from pulse_b200.ztasks import StrikeTaskB200
    m = gen()
    g = np.load(os.path.join(HERE, "golden", "ztasks.npz"))
    z = m.inputs(int(g["num_envs"]))
    n = z["body_state"].shape[0]
    rb, cf, roots = _sim(z)
    task = StrikeTaskB200(n, DEV, max_episode_length=m.MAX_LEN, dt=m.DT)
    task.pre_physics_step(z["prev_root_pos"].to(DEV))
    task.post_physics_step(rb, z["progress_buf"].to(DEV), roots[:, 1], cf[:, 24], contact_forces=cf)
    torch.cuda.synchronize()
    T = lambda k: torch.from_numpy(g[k])
    torch.testing.assert_close(task.obs_buf[:, :358].cpu(), T("self_obs"), atol=1e-4, rtol=0)
    torch.testing.assert_close(task.obs_buf[:, 358:].cpu(), T("strike_obs"), atol=1e-4, rtol=0)
    torch.testing.assert_close(task.rew_buf.cpu(), T("strike_reward"), atol=2e-4, rtol=0)
    assert torch.equal(task.reset_buf.cpu(), T("strike_reset")) and torch.equal(task._terminate_buf.cpu(), T("strike_terminate"))
    # target placement (humanoid_strike.py:124-145) with injected draws
    ids = torch.tensor([0, 5, 9], device=DEV)
    r = torch.tensor([[0.1, 0.5, 0.25, 0.0], [0.9, 1.0, 0.0, 0.5], [0.4, 0.0, 0.5, 0.25]], device=DEV)
    task.reset_target(ids, roots[:, 0], roots[:, 1], rand=r)
    d = (roots[ids, 1, 0:2] - roots[ids, 0, 0:2]).norm(dim=-1).cpu()
    torch.testing.assert_close(d, torch.tensor([1.0, 10.0, 0.5]), atol=1e-5, rtol=0)
    assert torch.allclose(roots[ids, 1, 2].cpu(), torch.full((3,), 0.9)) and float(roots[ids, 1, 7:].abs().max()) == 0.0
    torch.testing.assert_close(roots[ids, 1, 3:7].norm(dim=-1).cpu(), torch.ones(3), atol=1e-6, rtol=0)
