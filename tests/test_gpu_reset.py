"""EXPERIMENTAL reset path: `HumanoidImCompute.reset_ref_state` -- reset poses of a subset of envs
scattered into the simulator's tensors and the AMP history back-filled, built from the validated MotionLib-query and AMP-obs
kernels -- against the oracle's composition of the same reference steps.  Opt-in (PULSE_EXPERIMENTAL_RESET=1): written after
round 1's GPU budget was spent."""
import pytest
import torch

from tests.helpers import oracle_tables

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_reset_ref_state_matches_oracle():
    from oracle import pulse_oracle as po
    from pulse_b200.humanoid_im import HumanoidImCompute
    from pulse_b200.motion_lib import MotionLibB200
    tb = oracle_tables()
    ml = MotionLibB200.from_tables({k: getattr(tb, k) for k in ("gts", "grs", "lrs", "gvs", "gavs", "dvs", "motion_aa", "lengths", "num_frames", "dt",
                                                                 "length_starts")}, device=DEV)
    comp = HumanoidImCompute(ml)
    g = torch.Generator().manual_seed(11)
    N, n = 64, 17
    env_ids = torch.randperm(N, generator=g)[:n]
    motion_ids = torch.randint(0, tb.num_motions, (n,), generator=g)
    times = po.sample_time_interval(tb, motion_ids, torch.rand(n, generator=g)) + 0.4     # leave room for 9 earlier frames
    goff = torch.randn(n, 3, generator=g) * torch.tensor([1.0, 1.0, 0.0])
    root = torch.full((N, 13), 7.0, device=DEV)
    dof_state = torch.full((N, 69, 2), 7.0, device=DEV)
    body = torch.full((N, 26, 13), 7.0, device=DEV)
    amp = torch.full((N, 10, 196), 7.0, device=DEV)
    comp.reset_ref_state(env_ids.to(DEV), motion_ids.to(DEV), times.to(DEV), goff.to(DEV), root_states=root, dof_pos=dof_state[..., 0],
                         dof_vel=dof_state[..., 1], rigid_body_state=body, amp_obs_buf=amp)
    ms = po.motion_state(tb, motion_ids, times, goff)
    torch.testing.assert_close(root[env_ids].cpu(), torch.cat([ms["root_pos"], ms["root_rot"], ms["root_vel"], ms["root_ang_vel"]], -1), atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(dof_state[env_ids, :, 0].cpu(), ms["dof_pos"], atol=1e-4, rtol=1e-4)
    torch.testing.assert_close(dof_state[env_ids, :, 1].cpu(), ms["dof_vel"], atol=1e-4, rtol=1e-4)
    torch.testing.assert_close(body[env_ids, :24].cpu(), torch.cat([ms["rg_pos"], ms["rb_rot"], ms["body_vel"], ms["body_ang_vel"]], -1), atol=1e-5, rtol=1e-5)
    untouched = torch.ones(N, dtype=torch.bool)
    untouched[env_ids] = False
    assert float((root[untouched] - 7.0).abs().max()) == 0.0 and float((body[:, 24:] - 7.0).abs().max()) == 0.0
    assert float((amp[:, 0] - 7.0).abs().max()) == 0.0 and float((amp[untouched] - 7.0).abs().max()) == 0.0
    dt = po.STEP_DT
    for k in range(9):                                                                     # _init_amp_obs_ref (humanoid_amp.py:535-563)
        t_k = times + (-dt) * (k + 1)
        h = po.motion_state(tb, motion_ids, t_k)
        ref = po.amp_obs_smpl(h["root_pos"], h["root_rot"], h["root_vel"], h["root_ang_vel"], h["dof_pos"], h["dof_vel"],
                              h["rg_pos"][:, list(po.KEY_BODY_IDS)], po.amp_dof_subset())
        torch.testing.assert_close(amp[env_ids, k + 1].cpu(), ref, atol=1e-4, rtol=1e-4)
