"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Every array written here is an input to, or an output of, one of the reference's own functions
(imported through oracle/refshim/load_reference.py).  The reference ships no tests or golden
vectors for this path, so these fixtures are the pin: the oracle
(oracle/pulse_oracle.py) and the CUDA path are both checked against them.

Method-level glue that needs a live Isaac Gym task object (HumanoidIm._compute_reward etc.) is
replayed here as the same sequence of reference calls, cited inline.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))
from load_reference import load_learning, load_reference  # noqa: E402

PARENTS = [-1, 0, 1, 2, 3, 0, 5, 6, 7, 0, 9, 10, 11, 12, 11, 14, 15, 16, 17, 11, 19, 20, 21, 22]
CLIP_FRAMES = [2, 5, 17, 31, 64, 90, 45, 150]
KEY_BODY_IDS = [7, 3, 22, 17]
RESET_BODY_IDS = [j for j in range(24) if j not in (3, 4, 7, 8)]


def synth_clip(rng, n_frames, fps=30.0):
    """One AMASS-shaped clip in the on-disk schema convert_amass_isaac.py:127-136 writes."""
    from scipy.spatial.transform import Rotation as sRot
    aa = np.cumsum(rng.normal(0, 0.3 * 0.08, size=(n_frames, 24, 3)), axis=0) + rng.normal(0, 0.3, size=(1, 24, 3))
    aa[:, 0] *= 0.3
    local = sRot.from_rotvec(aa.reshape(-1, 3)).as_quat().reshape(n_frames, 24, 4)  # xyzw
    glob = np.zeros_like(local)
    for j, p in enumerate(PARENTS):
        rj = sRot.from_quat(local[:, j])
        glob[:, j] = rj.as_quat() if p < 0 else (sRot.from_quat(glob[:, p]) * rj).as_quat()
    trans = np.zeros((n_frames, 3))
    trans[:, :2] = np.cumsum(rng.normal(0, 0.02, size=(n_frames, 2)), axis=0) + rng.normal(0, 1.0, size=(1, 2))
    trans[:, 2] = 0.9 + 0.05 * np.sin(np.arange(n_frames) * 0.2)
    return {
        "pose_quat_global": glob, "pose_aa": aa.reshape(n_frames, 72), "root_trans_offset": torch.from_numpy(trans),
        "fps": fps, "beta": np.zeros(16), "gender": "neutral",
    }


def load_reference_motionlib(ref, pkl_path, n_clips):
    from phc.utils.motion_lib_smpl import MotionLibSMPL
    from phc.utils.motion_lib_base import FixHeightMode
    cfg = ref.EasyDict({"motion_file": pkl_path, "device": torch.device("cpu"), "fix_height": FixHeightMode.full_fix,
                        "min_length": -1, "max_length": -1, "im_eval": False, "multi_thread": False,
                        "smpl_type": "smpl", "randomrize_heading": True})
    lib = MotionLibSMPL(cfg)
    sk = ref.skeleton3d.SkeletonTree.from_mjcf(os.path.join(os.environ.get("PULSE_REFERENCE_ROOT", "/root/reference"),
                                                            "phc/data/assets/mjcf/smpl_humanoid.xml"))
    assert list(sk.parent_indices.numpy()) == PARENTS
    np.random.seed(7)  # heading randomisation uses np.random (motion_lib_smpl.py:106,134-139)
    torch.manual_seed(7)
    lib.load_motions(skeleton_trees=[sk] * n_clips, gender_betas=torch.zeros(n_clips, 17),
                     limb_weights=torch.zeros(n_clips, 10), random_sample=False)
    return lib


def np_(x):
    return x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x)


def main():
    import joblib
    cwd = os.getcwd()
    os.chdir("/tmp")  # keep data/smpl lookups (mesh parsers) off: motion_lib_smpl.py:53-68
    ref = load_reference()
    lrn = load_learning()
    rng = np.random.default_rng(0)
    gen = torch.Generator().manual_seed(0)

    # ------------------------------------------------------------------ MotionLib tables
    clips = {f"clip_{i:02d}": synth_clip(rng, nf) for i, nf in enumerate(CLIP_FRAMES)}
    pkl = "/tmp/pulse_golden_clips.pkl"
    joblib.dump(clips, pkl)
    lib = load_reference_motionlib(ref, pkl, len(CLIP_FRAMES))
    tables = {
        "gts": lib.gts, "grs": lib.grs, "lrs": lib.lrs, "gvs": lib.gvs, "gavs": lib.gavs, "dvs": lib.dvs,
        "motion_aa": lib._motion_aa, "lengths": lib._motion_lengths, "num_frames": lib._motion_num_frames,
        "dt": lib._motion_dt, "fps": lib._motion_fps, "length_starts": lib.length_starts,
        "motion_bodies": lib._motion_bodies, "motion_limb_weights": lib._motion_limb_weights,
    }
    np.savez_compressed(os.path.join(HERE, "motionlib.npz"), **{k: np_(v) for k, v in tables.items()})
    M = len(CLIP_FRAMES)
    lens = lib._motion_lengths

    # ------------------------------------------------------------------ get_motion_state queries
    ids, times = [], []
    for m in range(M):
        L = float(lens[m])
        dtm = float(lib._motion_dt[m])
        nf = CLIP_FRAMES[m]
        cand = [-0.5, 0.0, L, L + 0.3, 0.5 * L, dtm, dtm * (nf - 1), np.nextafter(np.float32(L), np.float32(0)),
                L * 0.999, 1e-6] + [dtm * k for k in range(0, nf, max(1, nf // 5))] + list(rng.uniform(0, L, size=12))
        ids += [m] * len(cand)
        times += cand
    ids = torch.tensor(ids, dtype=torch.long)
    times = torch.tensor(np.array(times, dtype=np.float32))
    offs = torch.from_numpy(rng.normal(0, 1, size=(len(ids), 3)).astype(np.float32))
    offs[::3] = 0
    res = lib.get_motion_state(ids, times, offset=offs)
    i0, i1, blend = lib._calc_frame_blend(times, lib._motion_lengths[ids], lib._motion_num_frames[ids], lib._motion_dt[ids])
    q = {"ids": ids, "times": times, "offset": offs, "frame_idx0": i0, "frame_idx1": i1, "blend": blend}
    for k in ("root_pos", "root_rot", "dof_pos", "root_vel", "root_ang_vel", "dof_vel", "motion_aa", "rg_pos", "rb_rot",
              "body_vel", "body_ang_vel"):
        q[k] = res[k]
    q["root_pos_smpl"] = lib.get_root_pos_smpl(ids, times)["root_pos"]
    phase = torch.rand(len(ids), generator=gen)
    # sample_time_interval (motion_lib_base.py:411-420) with the uniform draw injected
    q["phase"] = phase
    q["sampled_time"] = ((phase * lib._motion_lengths[ids]) / (1 / 30)).long() * (1 / 30)
    np.savez_compressed(os.path.join(HERE, "motion_state.npz"), **{k: np_(v) for k, v in q.items()})

    # ------------------------------------------------------------------ slerp / quaternion edge cases
    tu = ref.torch_utils
    qa = torch.nn.functional.normalize(torch.randn(64, 4, generator=gen), dim=-1)
    qb = torch.nn.functional.normalize(torch.randn(64, 4, generator=gen), dim=-1)
    qb[0:8] = qa[0:8]                       # identical -> |cos| >= 1 branch
    qb[8:16] = -qa[8:16]                    # antipodal sign flip
    tiny = torch.nn.functional.normalize(qa[16:32] + 1e-4 * torch.randn(16, 4, generator=gen), dim=-1)
    qb[16:32] = tiny                        # |sin| < 1e-3 lerp branch
    qb[32:40] = torch.nn.functional.normalize(qa[32:40] + 3e-3 * torch.randn(8, 4, generator=gen), dim=-1)
    tt = torch.rand(64, 1, generator=gen)
    tt[40:44] = 0.0
    tt[44:48] = 1.0
    qe = qa.clone()
    qe[0] = torch.tensor([0.0, 0.0, 0.0, 1.0])      # w == 1 -> default axis
    qe[1] = torch.tensor([0.0, 0.0, 0.0, -1.0])
    qe[2] = torch.nn.functional.normalize(torch.tensor([1e-6, 0.0, 0.0, 1.0]), dim=-1)
    em = torch.randn(64, 3, generator=gen)
    em[0] = 0.0
    em[1] = torch.tensor([1e-7, 0.0, 0.0])
    em[2] = torch.tensor([0.0, 3.5, 0.0])   # angle > pi wraps
    qd = {
        "qa": qa, "qb": qb, "t": tt, "slerp": tu.slerp(qa, qb, tt), "quat_mul": tu.quat_mul(qa, qb),
        "rotate": tu.my_quat_rotate(qa, qb[:, :3].contiguous()), "qe": qe, "exp_map": tu.quat_to_exp_map(qe),
        "angle": tu.quat_to_angle_axis(qe)[0], "axis": tu.quat_to_angle_axis(qe)[1], "tan_norm": tu.quat_to_tan_norm(qe),
        "heading": tu.calc_heading(qe), "heading_quat": tu.calc_heading_quat(qe), "heading_quat_inv": tu.calc_heading_quat_inv(qe),
        "em": em, "exp_map_to_quat": tu.exp_map_to_quat(em),
    }
    np.savez_compressed(os.path.join(HERE, "quat.npz"), **{k: np_(v) for k, v in qd.items()})

    # ------------------------------------------------------------------ one HumanoidIm post-physics step
    him, hum, hamp = ref.humanoid_im, ref.humanoid, ref.humanoid_amp
    for tag, N in (("n2", 2), ("n257", 257)):
        dt = float(torch.tensor(1.0 / 60.0) * 2)  # 2 * C-float sim dt (humanoid.py:122, config.py:47)
        motion_ids = torch.zeros(N, dtype=torch.long) if N == 2 else torch.arange(N) % M
        if N == 2:
            motion_ids[:] = 5  # config C1: both envs on the same clip
        L = lens[motion_ids]
        progress = torch.randint(0, 40, (N,), generator=gen)
        progress[: min(N, 6)] = torch.tensor([0, 1, 2, 3, 4, 5])[: min(N, 6)]
        start = ((torch.rand(N, generator=gen) * L) / (1 / 30)).long() * (1 / 30)
        start_off = torch.zeros(N)
        start_off[N // 2:] = -progress[N // 2:] * dt * (torch.rand(N - N // 2, generator=gen) < 0.3)
        goff = torch.zeros(N, 3)
        goff[::4, :2] = torch.randn((N + 3) // 4, 2, generator=gen)
        cycle = torch.zeros(N, dtype=torch.int32)
        cycle[::7] = 5
        t_rew = progress * dt + start + start_off
        pose = lib.get_motion_state(motion_ids, t_rew, offset=goff)
        noise = torch.full((N, 1, 1), 0.03)
        noise[::5] = 0.12  # some envs far enough to terminate
        body_pos = pose["rg_pos"] + noise * torch.randn(N, 24, 3, generator=gen)
        dq = torch.nn.functional.normalize(torch.cat([0.1 * torch.randn(N, 24, 3, generator=gen), torch.ones(N, 24, 1)], -1), dim=-1)
        body_rot = torch.nn.functional.normalize(tu.quat_mul(pose["rb_rot"], dq), dim=-1)
        body_vel = pose["body_vel"] + 0.5 * torch.randn(N, 24, 3, generator=gen)
        body_ang = pose["body_ang_vel"] + 0.5 * torch.randn(N, 24, 3, generator=gen)
        body_state = torch.cat([body_pos, body_rot, body_vel, body_ang], dim=-1)
        dof_pos = pose["dof_pos"] + 0.05 * torch.randn(N, 69, generator=gen)
        dof_vel = pose["dof_vel"] + 0.5 * torch.randn(N, 69, generator=gen)
        dof_force = 30 * torch.randn(N, 69, generator=gen)
        reset_buf = torch.zeros(N, dtype=torch.long)

        # --- _compute_reward (humanoid_im.py:853-919), full-body reward + power term
        specs = {"k_pos": 100.0, "k_rot": 10.0, "k_vel": 0.1, "k_ang_vel": 0.1, "w_pos": 0.5, "w_rot": 0.3, "w_vel": 0.1, "w_ang_vel": 0.1}
        rew, raw = him.compute_imitation_reward(body_pos[:, 0], body_rot[:, 0], body_pos, body_rot, body_vel, body_ang,
                                                pose["rg_pos"], pose["rb_rot"], pose["body_vel"], pose["body_ang_vel"], specs)
        power = torch.abs(torch.multiply(dof_force, dof_vel)).sum(dim=-1)
        power_reward = -0.0005 * power
        power_reward[progress <= 3] = 0
        rew = rew + power_reward
        raw = torch.cat([raw, power_reward[:, None]], dim=-1)

        # --- _compute_reset (humanoid_im.py:1119-1192), cycle_motion False
        pass_time = t_rew >= lens[motion_ids]
        rb = torch.tensor(RESET_BODY_IDS)
        term = torch.full((1, 24), 0.25)[..., rb]
        reset, terminated = him.compute_humanoid_im_reset(reset_buf, progress, torch.zeros(N, 24, 3), torch.zeros(4, dtype=torch.long),
                                                          body_pos[..., rb, :].clone(), pose["rg_pos"][..., rb, :].clone(),
                                                          pass_time, True, term, False, False)
        is_recovery = torch.logical_and(~pass_time, cycle > 0)
        reset = reset.clone()
        terminated = terminated.clone()
        reset[is_recovery] = 0
        terminated[is_recovery] = 0
        # eval-style mean criterion (flags.im_eval and not strict_eval), termination distance 0.5 (im_amp.py:174)
        reset_mean, terminated_mean = him.compute_humanoid_im_reset(reset_buf, progress, torch.zeros(N, 24, 3), torch.zeros(4, dtype=torch.long),
                                                                    body_pos[..., rb, :].clone(), pose["rg_pos"][..., rb, :].clone(),
                                                                    pass_time, True, torch.full((1, 24), 0.08)[..., rb], False, True)

        # --- _compute_observations (humanoid_im.py:677-706, :708-851), obs_v 6
        t_obs = (progress + 1) * dt + start + start_off
        nxt = lib.get_motion_state(motion_ids, t_obs, offset=goff)
        empty = torch.zeros(N, 0)
        self_obs = hum.compute_humanoid_observations_smpl_max(body_pos, body_rot, body_vel, body_ang, empty, empty, True, True, True, False, False)
        task_obs = him.compute_imitation_observations_v6(body_pos[:, 0], body_rot[:, 0], body_pos, body_rot, body_vel, body_ang,
                                                         nxt["rg_pos"], nxt["rb_rot"], nxt["body_vel"], nxt["body_ang_vel"], 1, True)
        track3 = torch.tensor([13, 18, 23])  # Head, L_Hand, R_Hand (env_pulse_im.yaml trackBodies, obs_v 7)
        task_obs_v7 = him.compute_imitation_observations_v7(body_pos[:, 0], body_rot[:, 0], body_pos[:, track3], body_vel[:, track3],
                                                            nxt["rg_pos"][:, track3], nxt["body_vel"][:, track3], 1, True)
        i0r, i1r, _ = lib._calc_frame_blend(t_rew, L, lib._motion_num_frames[motion_ids], lib._motion_dt[motion_ids])
        i0o, i1o, _ = lib._calc_frame_blend(t_obs, L, lib._motion_num_frames[motion_ids], lib._motion_dt[motion_ids])

        # --- AMP observation (humanoid_amp.py:632-667, :924-969) + history shift (:622-630)
        subset = torch.tensor([k for k in range(69) if (k // 3) not in (3, 7, 17, 22)])
        amp_cur = hamp.build_amp_observations_smpl(body_pos[:, 0], body_rot[:, 0], body_vel[:, 0], body_ang[:, 0], dof_pos, dof_vel,
                                                   body_pos[:, KEY_BODY_IDS], empty, empty, subset, True, True, True, False, False, True)
        NH = min(N, 48)  # history fixture on the first NH envs only (keeps the file small)
        amp_hist_in = torch.randn(NH, 10, 196, generator=gen)
        amp_hist_out = torch.cat([amp_cur[:NH, None], amp_hist_in[:, :9]], dim=1)

        d = {
            "dt": np.float32(dt), "motion_ids": motion_ids, "progress_buf": progress, "start_times": start, "start_offset": start_off,
            "global_offset": goff, "cycle_counter": cycle, "body_state": body_state, "dof_pos": dof_pos, "dof_vel": dof_vel,
            "dof_force": dof_force, "reset_buf_in": reset_buf,
            "rew_buf": rew, "reward_raw": raw, "reset_buf": reset, "terminate_buf": terminated,
            "reset_buf_mean": reset_mean, "terminate_buf_mean": terminated_mean,
            "self_obs": self_obs, "task_obs": task_obs, "task_obs_v7": task_obs_v7, "obs_buf": torch.cat([self_obs, task_obs], dim=-1),
            "ref_body_pos": nxt["rg_pos"], "ref_body_rot": nxt["rb_rot"], "ref_body_vel": nxt["body_vel"], "ref_dof_pos": nxt["dof_pos"],
            "frame_idx_rew": torch.stack([i0r, i1r], -1), "frame_idx_obs": torch.stack([i0o, i1o], -1),
            "amp_cur": amp_cur, "amp_hist_in": amp_hist_in, "amp_hist_out": amp_hist_out,
        }
        np.savez_compressed(os.path.join(HERE, f"step_{tag}.npz"), **{k: np_(v) for k, v in d.items()})

    # ------------------------------------------------------------------ agent-side arithmetic
    T, N = 32, 96
    fake = types.SimpleNamespace(horizon_length=T, gamma=0.99, tau=0.95, normalize_advantage=True, bounds_loss_coef=10)
    CA = lrn.common_agent.CommonAgent
    rewards = torch.randn(T, N, 1, generator=gen)
    values = torch.randn(T, N, 1, generator=gen)
    next_values = torch.randn(T, N, 1, generator=gen)
    fdones = (torch.rand(T, N, generator=gen) < 0.05).float()
    advs = CA.discount_values(fake, fdones, values, rewards, next_values)
    returns = advs + values
    flat = lambda x: x.transpose(0, 1).reshape(T * N, *x.shape[2:])  # swap_and_flatten01 [3P-memory]
    adv_norm = CA._calc_advs(fake, {"returns": flat(returns), "values": flat(values)})
    B, A = 512, 69
    old_nlp = torch.randn(B, generator=gen) * 3 + 60
    new_nlp = old_nlp + 0.3 * torch.randn(B, generator=gen)
    adv_b = torch.randn(B, generator=gen)
    a_info = CA._actor_loss(fake, old_nlp, new_nlp, adv_b, 0.2)
    c_info = CA._critic_loss(fake, values[:, 0, :], values[:, 1, :], 0.2, returns[:, 0, :], False)
    mu = torch.randn(B, A, generator=gen) * 0.8
    b_loss = CA.bound_loss(fake, mu)
    AA = lrn.amp_agent.AMPAgent
    logits = torch.randn(B, 1, generator=gen) * 3
    fake_amp = types.SimpleNamespace(_eval_disc=lambda x: x, _norm_disc_reward=lambda: False, _disc_reward_scale=2, ppo_device="cpu")
    disc_r = AA._calc_disc_rewards(fake_amp, logits.clone())
    bce_neg = AA._disc_loss_neg(fake_amp, logits)
    bce_pos = AA._disc_loss_pos(fake_amp, logits)
    qm, qv, pm, pv = (torch.randn(B, 32, generator=gen) for _ in range(4))
    kl = ref.loss_functions.kl_multi(qm, qv, pm, pv)
    rms = ref.running_mean_std.RunningMeanStd((7,))
    rms.train()
    x1 = torch.randn(300, 7, generator=gen) * 3 + 1
    x2 = torch.randn(200, 7, generator=gen) * 0.5 - 2
    y1 = rms(x1)
    y2 = rms(x2)
    rms.eval()
    y3 = rms(x1)
    # --- AMPAgent._disc_loss (amp_agent.py:895-952) on a small ReLU discriminator
    dm = torch.nn.Sequential(torch.nn.Linear(40, 32), torch.nn.ReLU(), torch.nn.Linear(32, 16), torch.nn.ReLU())
    dl = torch.nn.Linear(16, 1)
    torch.manual_seed(11)
    for m in list(dm) + [dl]:
        if isinstance(m, torch.nn.Linear):
            torch.nn.init.normal_(m.weight, std=0.3)
            torch.nn.init.normal_(m.bias, std=0.1)
    net = types.SimpleNamespace(get_disc_logit_weights=lambda: torch.flatten(dl.weight),
                                get_disc_weights=lambda: [torch.flatten(dm[0].weight), torch.flatten(dm[2].weight), torch.flatten(dl.weight)])
    fake_disc = types.SimpleNamespace(model=types.SimpleNamespace(a2c_network=net), _disc_logit_reg=0.01, _disc_grad_penalty=5,
                                      _disc_weight_decay=0.0001)
    for name in ("_disc_loss_neg", "_disc_loss_pos", "_compute_disc_acc"):
        setattr(fake_disc, name, types.MethodType(getattr(AA, name), fake_disc))
    d_agent, d_replay = torch.randn(24, 40, generator=gen), torch.randn(24, 40, generator=gen)
    d_demo = torch.randn(24, 40, generator=gen).requires_grad_(True)
    ev = lambda x: dl(dm(x))
    agent_cat = torch.cat([ev(d_agent), ev(d_replay)], dim=0)
    demo_logit = ev(d_demo)
    dinfo = AA._disc_loss(fake_disc, agent_cat, demo_logit, d_demo)
    dparams = [dm[0].weight, dm[0].bias, dm[2].weight, dm[2].bias, dl.weight, dl.bias]
    dgrads = torch.autograd.grad(dinfo["disc_loss"], dparams)

    ag = {
        "disc_w0": dm[0].weight, "disc_b0": dm[0].bias, "disc_w1": dm[2].weight, "disc_b1": dm[2].bias, "disc_w2": dl.weight, "disc_b2": dl.bias,
        "disc_agent": d_agent, "disc_replay": d_replay, "disc_demo": d_demo, "disc_loss": dinfo["disc_loss"],
        "disc_grad_penalty": dinfo["disc_grad_penalty"], "disc_logit_loss": dinfo["disc_logit_loss"],
        "disc_agent_acc": dinfo["disc_agent_acc"], "disc_demo_acc": dinfo["disc_demo_acc"],
        "disc_gw0": dgrads[0], "disc_gb0": dgrads[1], "disc_gw1": dgrads[2], "disc_gb1": dgrads[3], "disc_gw2": dgrads[4], "disc_gb2": dgrads[5],
        "rewards": rewards, "values": values, "next_values": next_values, "fdones": fdones, "advs": advs, "returns": returns,
        "adv_norm": adv_norm, "old_neglogp": old_nlp, "new_neglogp": new_nlp, "adv_b": adv_b, "actor_loss": a_info["actor_loss"],
        "actor_clipped": a_info["actor_clipped"], "critic_values": values[:, 1, :], "critic_returns": returns[:, 0, :],
        "critic_loss": c_info["critic_loss"], "mu": mu, "bound_loss": b_loss, "disc_logits": logits, "disc_reward": disc_r,
        "bce_neg": bce_neg, "bce_pos": bce_pos, "kl_qm": qm, "kl_qv": qv, "kl_pm": pm, "kl_pv": pv, "kl_multi": kl,
        "rms_x1": x1, "rms_x2": x2, "rms_y1": y1, "rms_y2": y2, "rms_y3": y3, "rms_mean": rms.running_mean,
        "rms_var": rms.running_var, "rms_count": rms.count,
    }
    np.savez_compressed(os.path.join(HERE, "agent.npz"), **{k: np_(v) for k, v in ag.items()})
    os.chdir(cwd)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
