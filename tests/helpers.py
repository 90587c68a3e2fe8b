"""Shared test helpers: golden-fixture loading and oracle table construction (test infra)."""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_npz(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: torch.from_numpy(z[k]) if z[k].ndim > 0 else z[k].item() for k in z.files}


def oracle_tables(device="cpu"):
    from oracle.pulse_oracle import MotionTables
    z = load_npz("motionlib.npz")
    return MotionTables(**{k: v.to(device) for k, v in z.items()})


def synthetic_tables(num_motions, seed=0, min_frames=5, max_frames=400, median_frames=150, fps=30.0):
    """AMASS-shaped flat MotionLib tables written directly: unit quaternions, smooth
    positions; bypasses the reference's 60 ms/clip loader.  Returns oracle MotionTables on CPU."""
    from oracle.pulse_oracle import MotionTables
    rng = np.random.default_rng(seed)
    nf = np.clip(np.exp(rng.normal(np.log(median_frames), 0.6, size=num_motions)).astype(np.int64), min_frames, max_frames)
    F = int(nf.sum())
    g = torch.Generator().manual_seed(seed)

    def unit(x):
        return torch.nn.functional.normalize(x, dim=-1)

    base_q = unit(torch.randn(num_motions, 24, 4, generator=g)).repeat_interleave(torch.from_numpy(nf), dim=0)
    grs = unit(base_q + 0.15 * torch.randn(F, 24, 4, generator=g))
    lrs = unit(torch.randn(num_motions, 24, 4, generator=g).repeat_interleave(torch.from_numpy(nf), dim=0)
               + 0.15 * torch.randn(F, 24, 4, generator=g))
    base_p = (torch.randn(num_motions, 24, 3, generator=g) * 0.4 + torch.tensor([0.0, 0.0, 0.9])).repeat_interleave(torch.from_numpy(nf), dim=0)
    gts = base_p + 0.05 * torch.randn(F, 24, 3, generator=g)
    gvs = torch.randn(F, 24, 3, generator=g)
    gavs = torch.randn(F, 24, 3, generator=g)
    dvs = torch.randn(F, 23, 3, generator=g)
    aa = torch.randn(F, 72, generator=g)
    nf_t = torch.from_numpy(nf)
    starts = torch.cumsum(nf_t, 0) - nf_t
    lengths = torch.tensor([(1.0 / fps) * (int(n) - 1) for n in nf], dtype=torch.float32)
    dt = torch.full((num_motions,), 1.0 / fps, dtype=torch.float32)
    return MotionTables(gts=gts, grs=grs, lrs=lrs, gvs=gvs, gavs=gavs, dvs=dvs, motion_aa=aa, lengths=lengths,
                        num_frames=nf_t, dt=dt, length_starts=starts, fps=torch.full((num_motions,), fps),
                        motion_bodies=torch.zeros(num_motions, 17), motion_limb_weights=torch.zeros(num_motions, 10))


def synthetic_step_inputs(tb, n_envs, seed=0, dt=None):
    """Per-env task buffers + simulator state near the reference pose, on CPU."""
    from oracle import pulse_oracle as po
    dt = po.STEP_DT if dt is None else dt
    g = torch.Generator().manual_seed(seed + 1)
    M = tb.num_motions
    motion_ids = torch.arange(n_envs) % M
    L = tb.lengths[motion_ids]
    progress = torch.randint(0, 40, (n_envs,), generator=g)
    start = po.sample_time_interval(tb, motion_ids, torch.rand(n_envs, generator=g))
    start_off = torch.zeros(n_envs)
    goff = torch.zeros(n_envs, 3)
    goff[::4, :2] = torch.randn((n_envs + 3) // 4, 2, generator=g)
    cycle = torch.zeros(n_envs, dtype=torch.int32)
    cycle[::7] = 5
    t = po.im_motion_times(progress, start, start_off, dt, plus_one=False)
    pose = po.motion_state(tb, motion_ids, t, goff)
    noise = torch.full((n_envs, 1, 1), 0.03)
    noise[::5] = 0.12
    body_pos = pose["rg_pos"] + noise * torch.randn(n_envs, 24, 3, generator=g)
    dq = torch.nn.functional.normalize(torch.cat([0.1 * torch.randn(n_envs, 24, 3, generator=g), torch.ones(n_envs, 24, 1)], -1), dim=-1)
    body_rot = torch.nn.functional.normalize(po.quat_mul(pose["rb_rot"], dq), dim=-1)
    body_vel = pose["body_vel"] + 0.5 * torch.randn(n_envs, 24, 3, generator=g)
    body_ang = pose["body_ang_vel"] + 0.5 * torch.randn(n_envs, 24, 3, generator=g)
    return {
        "motion_ids": motion_ids, "progress_buf": progress, "start_times": start, "start_offset": start_off,
        "global_offset": goff, "cycle_counter": cycle,
        "body_state": torch.cat([body_pos, body_rot, body_vel, body_ang], dim=-1).contiguous(),
        "dof_pos": pose["dof_pos"] + 0.05 * torch.randn(n_envs, 69, generator=g),
        "dof_vel": pose["dof_vel"] + 0.5 * torch.randn(n_envs, 69, generator=g),
        "dof_force": 30 * torch.randn(n_envs, 69, generator=g),
        "reset_buf_in": torch.zeros(n_envs, dtype=torch.long),
    }


def vae_golden():
    """tests/golden/vae.npz (reference-generated, make_golden_vae.py) + the oracle's view of its network weights."""
    from oracle import pulse_oracle as po
    g = load_npz("vae.npz")
    S, Tk, A, E, T, NE = [int(x) for x in g["dims"]]
    sd = {k[4:]: v for k, v in g.items() if k.startswith("net.")}
    nets = po.VaeNets.from_state_dict(sd, S)
    pnn_cols = [([g[f"pnn.actors.{c}.{i}.weight"] for i in (0, 2, 4)], [g[f"pnn.actors.{c}.{i}.bias"] for i in (0, 2, 4)]) for c in range(3)]
    composer = ([g[f"composer.{i}.weight"] for i in (0, 2, 4)], [g[f"composer.{i}.bias"] for i in (0, 2, 4)])
    return g, sd, nets, dict(S=S, Tk=Tk, A=A, E=E, T=T, NE=NE), pnn_cols, composer


def vae_param_list(nets):
    """Trainable tensors of the kin loss in reference naming order -> {reference parameter name: tensor}."""
    out = {}
    for name, (ws, bs) in (("z_mlp", nets.enc), ("z_prior", nets.prior), ("actor_mlp", (nets.dec[0][:-1], nets.dec[1][:-1]))):
        for i, (w, b) in enumerate(zip(ws, bs)):
            out[f"{name}.{2 * i}.weight"], out[f"{name}.{2 * i}.bias"] = w, b
    for name, (w, b) in (("z_mu", nets.enc_mu), ("z_logvar", nets.enc_logvar), ("z_prior_mu", nets.prior_mu),
                         ("z_prior_logvar", nets.prior_logvar), ("mu", (nets.dec[0][-1], nets.dec[1][-1]))):
        out[f"{name}.weight"], out[f"{name}.bias"] = w, b
    return out
