"""Rollout glue kernels (pulse_policy_post / pulse_value_post / pulse_amp_obs_row) and the PlayStepsB200 driver
(AMPAgent.play_steps, phc/learning/amp_agent.py:341-439) on the GPU."""
import math

import pytest
import torch

from tests.helpers import exact_step_inputs, exact_tables

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _policy(**kw):
    from pulse_b200.ppo import PPOPolicy
    return PPOPolicy(device=DEV, seed=3, **kw)


def test_policy_post_injected_noise_matches_formula():
    """a = mu + exp(logstd) eps; neglogp = 0.5 sum(((a-mu)/sigma)^2) + 0.5 A log(2 pi) + sum(logstd) [rl_games ModelA2CContinuousLogStd];
    value de-normalisation (running_mean_std.py:84-87); PD targets (humanoid.py:1392-1394) -- all through strided experience slices."""
    pol = _policy()
    pol.value_rms.running_mean.fill_(0.7)
    pol.value_rms.running_var.fill_(2.3)
    M, A, T = 300, 69, 5
    g = torch.Generator(device=DEV).manual_seed(1)
    obs = torch.randn(M, 934, device=DEV, generator=g)
    eps = torch.randn(M, A, device=DEV, generator=g)
    actions, mus, nlp = torch.zeros(M, T, A, device=DEV), torch.zeros(M, T, A, device=DEV), torch.zeros(M, T, device=DEV)
    values = torch.zeros(T, M, 1, device=DEV)
    off, sc, pd = torch.randn(A, device=DEV, generator=g), torch.rand(A, device=DEV, generator=g) * 3, torch.zeros(M, A, device=DEV)
    t = 3
    pol.act_into(obs, actions=actions[:, t], neglogp=nlp[:, t], mus=mus[:, t], values=values[t], pd=(off, sc, pd), eps=eps)
    ref = pol.act(obs, eps=eps)          # the validated round-1 path (test_gpu_ppo)
    torch.cuda.synchronize()
    torch.testing.assert_close(mus[:, t], ref["mus"], atol=0, rtol=0)
    torch.testing.assert_close(actions[:, t], ref["actions"], atol=1e-6, rtol=1e-6)
    torch.testing.assert_close(nlp[:, t], ref["neglogpacs"], atol=2e-4, rtol=1e-5)
    torch.testing.assert_close(values[t], ref["values"], atol=1e-6, rtol=1e-6)
    torch.testing.assert_close(pd, off + sc * actions[:, t], atol=1e-6, rtol=1e-6)
    for buf in (actions, mus, nlp.unsqueeze(-1)):        # nothing outside slice t was touched
        keep = [i for i in range(T) if i != t]
        assert float(buf[:, keep].abs().max()) == 0.0


def test_policy_post_philox_noise_statistics_and_reproducibility():
    pol = _policy()
    M, A = 4096, 69
    obs = torch.randn(M, 934, device=DEV)
    outs = []
    for step, bump in ((0, 0), (0, 0), (1, 0), (0, 32)):
        if bump:
            pol.advance_rng(bump)
        actions, mus, nlp = torch.zeros(M, A, device=DEV), torch.zeros(M, A, device=DEV), torch.zeros(M, device=DEV)
        pol.act_into(obs, actions=actions, neglogp=nlp, mus=mus, rng_step=step)
        outs.append((actions.clone(), mus.clone(), nlp.clone()))
    torch.cuda.synchronize()
    assert torch.equal(outs[0][0], outs[1][0])                                     # same (seed, offset, step) -> same draws
    assert not torch.equal(outs[0][0], outs[2][0]) and not torch.equal(outs[0][0], outs[3][0])
    sigma = math.exp(-2.9)
    e = (outs[0][0] - outs[0][1]) / sigma                                          # the standard-normal draws
    assert abs(float(e.mean())) < 0.01 and abs(float(e.var()) - 1.0) < 0.02
    assert abs(float((e ** 4).mean()) - 3.0) < 0.15                                # kurtosis of a Gaussian
    c = torch.corrcoef(e[:, :8].t())
    assert float((c - torch.eye(8, device=DEV)).abs().max()) < 0.06               # neighbouring columns (one Box-Muller pair) uncorrelated
    ref_nlp = 0.5 * (e ** 2).sum(-1) + 0.5 * A * math.log(2 * math.pi) + A * (-2.9)
    torch.testing.assert_close(outs[0][2], ref_nlp, atol=2e-3, rtol=1e-5)


def test_value_post_matches_reference_expression():
    pol = _policy()
    pol.value_rms.running_mean.fill_(-0.4)
    pol.value_rms.running_var.fill_(0.6)
    M = 777
    obs = torch.randn(M, 934, device=DEV)
    term = (torch.rand(M, device=DEV) < 0.3).long()
    out = torch.zeros(4, M, 1, device=DEV)
    pol.critic_values_into(obs, out[2].view(-1), terminate=term)
    ref = pol.critic_values(obs) * (1.0 - term.unsqueeze(1).float())
    torch.cuda.synchronize()
    torch.testing.assert_close(out[2], ref, atol=1e-6, rtol=1e-6)
    assert float(out[[0, 1, 3]].abs().max()) == 0.0


def _sim(n, clips=50, seed=2):
    from pulse_b200.humanoid_im import HumanoidImCompute
    from pulse_b200.motion_lib import MotionLibB200
    tb = exact_tables(clips, seed=seed)
    z, _ = exact_step_inputs(tb, n, seed=seed + 1)
    ml = MotionLibB200.from_tables({k: getattr(tb, k) for k in ("gts", "grs", "lrs", "gvs", "gavs", "dvs", "motion_aa", "lengths", "num_frames", "dt",
                                                                 "length_starts")}, device=DEV)
    comp = HumanoidImCompute(ml)
    body = torch.zeros(n, 26, 13, device=DEV)
    body[:, :24] = z["body_state"].to(DEV)
    dof_state = torch.zeros(n, 72, 2, device=DEV)
    dof_state[:, :69, 0], dof_state[:, :69, 1] = z["dof_pos"].to(DEV), z["dof_vel"].to(DEV)
    root = torch.zeros(n, 2, 13, device=DEV)
    root[:, 0] = body[:, 0]
    sim = dict(body_state=body, root_states=root[:, 0], dof_pos=dof_state[:, :69, 0], dof_vel=dof_state[:, :69, 1], dof_force=z["dof_force"].to(DEV),
               progress_buf=z["progress_buf"].to(DEV), motion_ids=z["motion_ids"].to(DEV), motion_start_times=z["start_times"].to(DEV),
               motion_start_offset=z["start_offset"].to(DEV), global_offset=z["global_offset"].to(DEV), cycle_counter=z["cycle_counter"].to(DEV),
               contact_forces=torch.zeros(n, 26, 3, device=DEV), actor_ids=torch.arange(n, dtype=torch.int32, device=DEV) * 2)
    return tb, comp, sim


def test_amp_obs_row_equals_in_place_shift_and_takes_fresh_rows():
    """Row mode [cur | prev[:1764]] == the validated in-place shift kernel (test_gpu_step), step after step; flagged envs take their history
    from the back-filled rows and the flag is cleared."""
    n, T = 97, 4
    tb, comp, sim = _sim(n)
    g = torch.Generator(device=DEV).manual_seed(5)
    ref_buf = torch.randn(n, 10, 196, device=DEV, generator=g)
    exp = torch.zeros(n, T, 1960, device=DEV)
    exp[:, T - 1] = ref_buf.view(n, 1960)                   # "previous iteration's last row"
    fresh_rows = torch.randn(n, 10, 196, device=DEV, generator=g)
    fresh = torch.zeros(n, dtype=torch.int32, device=DEV)
    for t in range(T):
        sim["body_state"][:, :24, :3] += 0.01 * (t + 1)     # the state changes between steps
        if t == 2:
            fresh[::7] = 1
            ref_buf[::7] = fresh_rows[::7]                  # what _init_amp_obs leaves in the task-side buffer
        comp.amp_obs(body_state=sim["body_state"], dof_pos=sim["dof_pos"], dof_vel=sim["dof_vel"], amp_obs_buf=ref_buf)
        prev = exp[:, t - 1] if t > 0 else exp[:, T - 1]
        comp.amp_obs_row(body_state=sim["body_state"], dof_pos=sim["dof_pos"], dof_vel=sim["dof_vel"], prev=prev, out=exp[:, t], fresh=fresh,
                         fresh_rows=fresh_rows)
        torch.cuda.synchronize()
        assert torch.equal(exp[:, t], ref_buf.view(n, 1960)), t
        assert int(fresh.sum()) == 0


@pytest.mark.parametrize("single_graph", [False, True])
def test_play_steps_graph_replay_equals_eager_and_last_step_matches_oracle(single_graph):
    """The whole horizon through CUDA graphs (segment graphs / ONE graph with the step kernels inside) reproduces the eager run bit for bit
    (Philox draws are a function of (seed, offset, step)), twice in a row; the final observation / reward / reset of every env equal the
    oracle's fused step on the final simulator state; the bookkeeping identities of play_steps hold."""
    from oracle import pulse_oracle as po
    from pulse_b200.rollout import PlayStepsB200
    n, T = 192, 6
    runs = []
    for graphs in (False, True):
        tb, comp, sim = _sim(n)
        pol = _policy(with_disc=True)
        ps = PlayStepsB200(comp, pol, sim, horizon=T, use_graphs=graphs, single_graph=single_graph, reset_seed=9)
        ps.first_observation()
        for it in range(3):
            ps.play_steps()
            if it == 2:   # next_values = unnorm(critic(next obs)) * (1 - terminated), with the value statistics of the rollout (before finish() merges)
                nv = (ps.policy.critic_values(ps.obs_carry) * (1.0 - ps.terminate_buf.unsqueeze(1).float())).clone()
            ps.finish()
        torch.cuda.synchronize()
        runs.append((ps, sim, tb, nv))
    a, b = runs[0][0], runs[1][0]
    for name in ("obses", "actions", "mus", "neglogp", "amp_obs", "values", "next_values", "rewards", "dones", "obs_carry", "adv", "ret"):
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    for k in ("progress_buf", "motion_start_times", "body_state"):
        assert torch.equal(runs[0][1][k], runs[1][1][k]), k
    ps, sim, tb, nv = runs[1]
    assert 0.0 < float(ps.dones.mean()) < 0.9                      # some envs reset inside the horizon, not all
    assert len(ps.step_kernel_ms()) == T and all(0.0 < ms < 5.0 for ms in ps.step_kernel_ms())
    # last env step vs the oracle on the final state (progress_buf already advanced by the kernel)
    cpu = lambda t: t.detach().cpu()
    ref = po.humanoid_im_step(tb, po.ImStepConfig(), cpu(sim["body_state"][:, :24]), cpu(sim["dof_vel"]), cpu(sim["dof_force"]), cpu(sim["progress_buf"]),
                              cpu(sim["motion_ids"]), cpu(sim["motion_start_times"]), cpu(sim["motion_start_offset"]), cpu(sim["global_offset"]),
                              cpu(sim["cycle_counter"]), torch.zeros(n, dtype=torch.long))
    torch.testing.assert_close(cpu(ps.obs_carry), ref["obs_buf"], atol=1e-4, rtol=0)
    torch.testing.assert_close(cpu(ps.rewards[T - 1]), ref["rew_buf"], atol=1e-4, rtol=0)
    assert torch.equal(cpu(ps.dones[T - 1]), ref["reset_buf"].float()) and torch.equal(cpu(ps.reset_buf), ref["reset_buf"])
    torch.testing.assert_close(ps.next_values[T - 1], nv, atol=1e-5, rtol=1e-5)     # on the PRE-reset observation
    # envs reset before step t start it with progress 1 after the step's own increment
    assert bool((sim["progress_buf"] >= 1).all())


def test_overlapped_schedule_equals_the_sequential_one():
    """_whole_overlapped (next values of step t beside reset / actor of step t+1, AMP row beside the step kernel; three streams inside one
    graph) produces bit for bit the experience of the sequential schedule, eagerly and replayed, over several iterations."""
    from pulse_b200.rollout import PlayStepsB200
    n, T = 1024, 8
    runs = {}
    for overlap, graphs in ((False, True), (True, False), (True, True), ("amp_after_step", True)):
        tb, comp, sim = _sim(n, clips=64)
        pol = _policy(with_disc=True)
        ps = PlayStepsB200(comp, pol, sim, horizon=T, use_graphs=graphs, single_graph=True, reset_seed=9)
        ps.overlap = bool(overlap)
        if overlap == "amp_after_step":
            ps.amp_with_step = False        # the ordering used above 4096 envs
        ps.first_observation()
        for it in range(4):
            ps.play_steps()
            ps.finish()
        torch.cuda.synchronize()
        runs[(overlap, graphs)] = (ps, sim)
    ref, ref_sim = runs[(False, True)]
    for key in ((True, False), (True, True), ("amp_after_step", True)):
        ps, sim = runs[key]
        for name in ("obses", "actions", "mus", "neglogp", "amp_obs", "values", "next_values", "rewards", "dones", "obs_carry", "adv", "ret"):
            assert torch.equal(getattr(ref, name), getattr(ps, name)), (key, name)
        for k in ("progress_buf", "motion_start_times", "body_state"):
            assert torch.equal(ref_sim[k], sim[k]), (key, k)
