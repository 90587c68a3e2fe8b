"""Fixture for SURVEY 8f-4: every `compute_imitation_observations*` variant of the UNMODIFIED reference (humanoid_im.py:1222-1540) on
seeded random inputs -- full body and the 3-point VR subset, one and three future samples, upright and non-upright starts.

  python tests/golden/make_golden_obs_versions.py     (needs /root/reference; writes tests/golden/obs_versions.npz)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

CASES = [  # tag, version, tracked bodies, time_steps, upright
    ("v1_full_t1", 1, list(range(24)), 1, True), ("v1_3pt_t3", 1, [13, 18, 23], 3, True),
    ("v2_full_t1", 2, list(range(24)), 1, True), ("v2_sub_t1", 2, [0, 4, 8, 13, 18, 23], 1, False),
    ("v3_full_t3", 3, list(range(24)), 3, True),
    ("v6_full_t1", 6, list(range(24)), 1, True), ("v6_full_t3", 6, list(range(24)), 3, False), ("v6_3pt_t1", 6, [13, 18, 23], 1, True),
    ("v7_3pt_t1", 7, [13, 18, 23], 1, True), ("v7_3pt_t3", 7, [13, 18, 23], 3, True), ("v7_full_t3", 7, list(range(24)), 3, False),
    ("v8_full_t1", 8, list(range(24)), 1, True), ("v8_3pt_t1", 8, [13, 18, 23], 1, False),
    ("v9_full_t1", 9, list(range(24)), 1, True), ("v9_3pt_t3", 9, [13, 18, 23], 3, True), ("v9_sub_t3", 9, [0, 4, 8, 13, 18, 23], 3, False),
]


def inputs(N, T, seed):
    g = torch.Generator().manual_seed(seed)
    unit = lambda q: q / q.norm(dim=-1, keepdim=True)
    body_state = torch.zeros(N, 24, 13)
    body_state[..., 0:3] = torch.randn(N, 24, 3, generator=g) * 0.4 + torch.tensor([0.0, 0.0, 0.9])
    body_state[..., 3:7] = unit(torch.randn(N, 24, 4, generator=g))
    body_state[..., 7:13] = torch.randn(N, 24, 6, generator=g)
    ref_pos = body_state[:, None, :, 0:3] + torch.randn(N, T, 24, 3, generator=g) * 0.1
    ref_rot = unit(torch.randn(N, T, 24, 4, generator=g))
    ref_vel = torch.randn(N, T, 24, 3, generator=g)
    ref_ang = torch.randn(N, T, 24, 3, generator=g)
    dof_pos = torch.randn(N, 69, generator=g) * 0.5
    ref_dof_pos = torch.randn(N * T, 69, generator=g) * 0.5
    return body_state, ref_pos.reshape(N * T, 24, 3), ref_rot.reshape(N * T, 24, 4), ref_vel.reshape(N * T, 24, 3), ref_ang.reshape(N * T, 24, 3), dof_pos, ref_dof_pos


def reference_obs(him, version, track, T, upright, body_state, rp, rr, rv, rw, dof_pos, ref_dof_pos):
    """The call `_compute_task_obs` makes for this version (humanoid_im.py:757-833) on the subset rows."""
    tr = torch.tensor(track)
    bp, br, bv, bw = (body_state[:, tr, a:b] for a, b in ((0, 3), (3, 7), (7, 10), (10, 13)))
    root_pos, root_rot = body_state[:, 0, 0:3], body_state[:, 0, 3:7]
    rps, rrs, rvs, rws = rp[:, tr], rr[:, tr], rv[:, tr], rw[:, tr]
    if version == 1:
        return him.compute_imitation_observations(root_pos, root_rot, bp, br, bv, bw, rps, rrs, rvs, rws, T, upright)
    if version == 2:
        rds = ref_dof_pos.reshape(-1, 23, 3)[..., tr[1:] - 1, :]
        ds = dof_pos.reshape(-1, 23, 3)[..., tr[1:] - 1, :]
        return him.compute_imitation_observations_v2(root_pos, root_rot, bp, br, bv, bw, ds, rps, rrs, rvs, rws, rds, T, upright)
    if version == 3:
        return him.compute_imitation_observations_v3(root_pos, root_rot, bp, br, bv, bw, rps, rrs, rvs, rws, T, upright)
    if version == 6:
        return him.compute_imitation_observations_v6(root_pos, root_rot, bp, br, bv, bw, rps, rrs, rvs, rws, T, upright)
    if version == 7:
        return him.compute_imitation_observations_v7(root_pos, root_rot, bp, bv, rps, rvs, T, upright)
    if version == 8:
        return him.compute_imitation_observations_v8(root_pos, root_rot, bp, br, bv, bw, rps, rrs, rvs, rws, T, upright)
    if version == 9:
        return him.compute_imitation_observations_v9(root_pos, root_rot, bp, br, bv, bw, rps, rrs, rvs[:, 0], rws[:, 0], T, upright)
    raise ValueError(version)


def main():
    from oracle.refshim.load_reference import load_reference
    ref = load_reference()
    him = ref["humanoid_im"] if isinstance(ref, dict) else ref.humanoid_im
    out = {}
    N = 37
    for k, (tag, version, track, T, upright) in enumerate(CASES):
        data = inputs(N, T, 100 + k)
        obs = reference_obs(him, version, track, T, upright, *data)
        out[tag] = obs.numpy().astype(np.float32)
    out["num_envs"] = np.int64(N)
    np.savez_compressed(os.path.join(HERE, "obs_versions.npz"), **out)
    print({k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.ndim})


if __name__ == "__main__":
    main()
