"""pulse_im_task_obs (every observation version, tracked-body subsets, future samples) through the C ABI against
(1) the fixtures the UNMODIFIED reference wrote (tests/golden/obs_versions.npz) and (2) the oracle on MotionLib-driven inputs via
HumanoidImCompute.task_obs (the fut_tracks sample times included).  Bar: 1e-4 (north_star: observations within 1e-4)."""
import ctypes as C
import importlib.util
import os

import numpy as np
import pytest
import torch

from tests.helpers import synthetic_step_inputs, synthetic_tables

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ATOL = 1e-4


def _gen():
    spec = importlib.util.spec_from_file_location("make_golden_obs_versions", os.path.join(HERE, "golden", "make_golden_obs_versions.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_every_version_matches_the_reference_fixture():
    from pulse_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda:0")
    m = _gen()
    z = np.load(os.path.join(HERE, "golden", "obs_versions.npz"))
    N = int(z["num_envs"])
    for k, (tag, version, track, T, upright) in enumerate(m.CASES):
        body_state, rp, rr, rv, rw, dof_pos, ref_dof = (x.to(dev).contiguous() for x in m.inputs(N, T, 100 + k))
        full = torch.full((N, 27, 13), 3.0, device=dev)          # a simulator view with extra bodies per env
        full[:, :24] = body_state
        dof_state = torch.zeros(N, 69, 2, device=dev)            # Isaac Gym dof-state layout: position = [:, :, 0]
        dof_state[:, :, 0] = dof_pos
        tr = torch.tensor(track, dtype=torch.int32, device=dev)
        size = lib.pulse_task_obs_size(version, len(track), T)
        assert size == z[tag].shape[1], tag
        obs = torch.full((N, size + 5), -7.0, device=dev)
        a = _lib.TaskObsArgs(body_state=full.data_ptr(), body_env_stride=full.stride(0), track_ids=tr.data_ptr(), num_track=len(track), time_steps=T,
                             version=version, upright=int(upright), ref_pos=rp.data_ptr(), ref_rot=rr.data_ptr(), ref_vel=rv.data_ptr(),
                             ref_ang_vel=rw.data_ptr(), dof_pos=dof_state.data_ptr(), dof_env_stride=dof_state.stride(0), dof_elem_stride=2,
                             ref_dof_pos=ref_dof.data_ptr(), obs=obs.data_ptr(), obs_stride=obs.stride(0), num_envs=N)
        _lib.check(lib.pulse_im_task_obs(C.byref(a), _lib.current_stream(dev)), "pulse_im_task_obs")
        torch.cuda.synchronize()
        torch.testing.assert_close(obs[:, :size].cpu(), torch.from_numpy(z[tag]), atol=ATOL, rtol=0, msg=lambda s, tag=tag: f"{tag}: {s}")
        assert float(obs[:, size:].min()) == -7.0 and float(obs[:, size:].max()) == -7.0      # nothing written past the version's size


@pytest.mark.parametrize("version,track,T", [(7, [13, 18, 23], 3), (6, list(range(24)), 1), (9, [0, 4, 8, 13, 18, 23], 3), (1, list(range(24)), 2)])
def test_task_obs_with_motionlib_queries_matches_oracle(version, track, T):
    from oracle import pulse_oracle as po
    from pulse_b200.humanoid_im import HumanoidImCompute, ImConfig
    from tests.test_gpu_step import _mlib
    dev = torch.device("cuda:0")
    tb = synthetic_tables(48, seed=3)
    n = 1500
    zin = synthetic_step_inputs(tb, n, seed=11)
    comp = HumanoidImCompute(_mlib(tb), ImConfig())
    sample_dt = 0.5
    tr = torch.tensor(track, dtype=torch.int32, device=dev)
    size = comp.lib.pulse_task_obs_size(version, len(track), T)
    obs = torch.zeros(n, size, device=dev)
    comp.task_obs(version=version, body_state=zin["body_state"].to(dev), progress_buf=zin["progress_buf"].to(dev), motion_ids=zin["motion_ids"].to(dev),
                  motion_start_times=zin["start_times"].to(dev), motion_start_offset=zin["start_offset"].to(dev),
                  global_offset=zin["global_offset"].to(dev), track_ids=tr, obs_buf=obs, time_steps=T, sample_dt=sample_dt)
    torch.cuda.synchronize()
    # oracle: the reference's sample times (humanoid_im.py:723-732), MotionLib query, observation
    t0 = (zin["progress_buf"] + 1) * comp.cfg.dt
    times = (t0[:, None] + (torch.arange(T) * sample_dt)[None, :] + zin["start_times"][:, None] + zin["start_offset"][:, None]).reshape(-1) if T > 1 \
        else t0 + zin["start_times"] + zin["start_offset"]
    ref = po.motion_state(tb, zin["motion_ids"].repeat_interleave(T), times.float(), zin["global_offset"].repeat_interleave(T, dim=0))
    bs = zin["body_state"]
    trl = torch.tensor(track)
    want = po.imitation_obs(version, bs[:, 0, 0:3], bs[:, 0, 3:7], bs[:, trl, 0:3], bs[:, trl, 3:7], bs[:, trl, 7:10], bs[:, trl, 10:13],
                            ref["rg_pos"][:, trl], ref["rb_rot"][:, trl], ref["body_vel"][:, trl], ref["body_ang_vel"][:, trl], T, True)
    torch.testing.assert_close(obs.cpu(), want, atol=ATOL, rtol=0)
