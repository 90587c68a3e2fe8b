"""Downstream latent-space speed and strike tasks: host-side mirrors of `HumanoidSpeed(Z)` / `HumanoidStrike(Z)`
(phc/env/tasks/humanoid_speed.py, humanoid_strike.py) for the post-physics path -- reward, reset, observation in ONE launch
(`pulse_ztask_step`) -- and the task-state updates (`_update_task` / `_reset_task`).  Like the reach task (pulse_b200/reach.py) the policy
acts in the frozen PULSE latent space (`PulseVAE.compute_z_actions`); Isaac Gym keeps the physics and owns the state tensors.
"""
import ctypes as C
import math
from typing import Optional, Sequence

import torch

from . import _lib
from .reach import SMPL_BODY_NAMES

SPEED_OBS, STRIKE_OBS = 361, 373      # 358 self observation + 3 / + 15


def _mask(names: Sequence[str]) -> int:
    m = 0
    for n in names:
        m |= 1 << SMPL_BODY_NAMES.index(n)
    return m


class _ZTaskBase:
    kind, obs_size = 0, 0

    def __init__(self, num_envs: int, device, contact_bodies, max_episode_length: int, enable_early_termination: bool, termination_height: float,
                 dt: float):
        self.device, self.num_envs = torch.device(device), int(num_envs)
        self.contact_body_mask = _mask(contact_bodies)
        self.strike_body_mask = 0
        self.max_episode_length, self.enable_early_termination, self.dt = int(max_episode_length), bool(enable_early_termination), float(dt)
        dev = self.device
        self.termination_heights = torch.full((24,), termination_height, device=dev)
        self._prev_root_pos = torch.zeros(num_envs, 3, device=dev)
        self.obs_buf = torch.zeros(num_envs, self.obs_size, device=dev)
        self.rew_buf = torch.zeros(num_envs, device=dev)
        self.reset_buf = torch.zeros(num_envs, dtype=torch.int64, device=dev)
        self._terminate_buf = torch.zeros(num_envs, dtype=torch.int64, device=dev)
        self.lib = _lib.load()

    def pre_physics_step(self, root_states: torch.Tensor) -> None:
        """`self._prev_root_pos[:] = self._humanoid_root_states[..., 0:3]` (humanoid_speed.py:73-76, humanoid_strike.py pre_physics_step)."""
        self._prev_root_pos.copy_(root_states[:, 0:3])

    def _args(self, rigid_body_state, progress_buf, contact_forces):
        if rigid_body_state.dim() != 3 or rigid_body_state.shape[1] < 24 or rigid_body_state.stride(1) != 13 or rigid_body_state.stride(2) != 1:
            raise _lib.PulseError("rigid_body_state must be a [N, B>=24, 13] view with row stride 13")
        return _lib.ZTaskStepArgs(
            kind=self.kind, enable_early_termination=int(self.enable_early_termination), body_state=rigid_body_state.data_ptr(),
            body_env_stride=rigid_body_state.stride(0), contact_forces=contact_forces.data_ptr() if contact_forces is not None else None,
            contact_env_stride=contact_forces.stride(0) if contact_forces is not None else 0, termination_heights=self.termination_heights.data_ptr(),
            contact_body_mask=self.contact_body_mask, strike_body_mask=self.strike_body_mask, progress_buf=progress_buf.data_ptr(),
            max_episode_length=self.max_episode_length, prev_root_pos=self._prev_root_pos.data_ptr(), dt=self.dt,
            obs_buf=self.obs_buf.data_ptr(), obs_stride=self.obs_buf.stride(0), rew_buf=self.rew_buf.data_ptr(), reset_buf=self.reset_buf.data_ptr(),
            terminate_buf=self._terminate_buf.data_ptr())

    def _launch(self, a) -> None:
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pulse_ztask_step(C.byref(a), self.num_envs, _lib.current_stream(self.device)), "pulse_ztask_step")


class SpeedTaskB200(_ZTaskBase):
    """HumanoidSpeed (humanoid_speed.py:17-240): run along +x at a commanded speed."""
    kind, obs_size = _lib.ZTASK_SPEED, SPEED_OBS

    def __init__(self, num_envs: int, device="cuda:0", contact_bodies: Sequence[str] = ("R_Ankle", "L_Ankle", "R_Toe", "L_Toe"),
                 tar_speed_min: float = 0.0, tar_speed_max: float = 5.0, speed_change_steps_min: int = 100, speed_change_steps_max: int = 200,
                 max_episode_length: int = 300, enable_early_termination: bool = True, termination_height: float = 0.15, dt: float = 1.0 / 30.0,
                 power_reward: bool = False, power_coefficient: float = 0.0005):
        super().__init__(num_envs, device, contact_bodies, max_episode_length, enable_early_termination, termination_height, dt)
        self._tar_speed_min, self._tar_speed_max = tar_speed_min, tar_speed_max
        self._speed_change_steps_min, self._speed_change_steps_max = speed_change_steps_min, speed_change_steps_max
        self.power_reward, self.power_coefficient = power_reward, power_coefficient
        dev = self.device
        self._tar_speed = torch.ones(num_envs, device=dev)                                    # :41
        self._speed_change_steps = torch.zeros(num_envs, dtype=torch.int64, device=dev)       # :39
        self.reward_raw = torch.zeros(num_envs, 2 if power_reward else 1, device=dev)

    def get_task_obs_size(self) -> int:
        return 3

    def update_task(self, progress_buf: torch.Tensor, rand01: Optional[torch.Tensor] = None, steps: Optional[torch.Tensor] = None) -> None:
        """_update_task / _reset_task (:157-175) without the `nonzero` host sync: every env draws, the envs whose progress reached
        `_speed_change_steps` take the draw (same distribution; the reference draws for the selected subset only)."""
        n, dev = self.num_envs, self.device
        rand01 = torch.rand(n, device=dev) if rand01 is None else rand01
        steps = torch.randint(self._speed_change_steps_min, self._speed_change_steps_max, (n,), device=dev) if steps is None else steps
        m = progress_buf >= self._speed_change_steps
        self._tar_speed.copy_(torch.where(m, (self._tar_speed_max - self._tar_speed_min) * rand01 + self._tar_speed_min, self._tar_speed))
        self._speed_change_steps.copy_(torch.where(m, progress_buf + steps, self._speed_change_steps))

    def post_physics_step(self, rigid_body_state: torch.Tensor, progress_buf: torch.Tensor, contact_forces: Optional[torch.Tensor] = None,
                          dof_force: Optional[torch.Tensor] = None, dof_vel: Optional[torch.Tensor] = None) -> None:
        """_compute_reward (:199-222) + _compute_reset (Humanoid's) + _compute_observations in one launch."""
        a = self._args(rigid_body_state, progress_buf, contact_forces)
        a.tar_speed, a.reward_raw, a.raw_stride = self._tar_speed.data_ptr(), self.reward_raw.data_ptr(), self.reward_raw.stride(0)
        if self.power_reward:
            if dof_force is None or dof_vel is None:
                raise _lib.PulseError("power_reward needs dof_force and dof_vel")
            a.dof_force, a.dof_force_stride, a.power_coefficient = dof_force.data_ptr(), dof_force.stride(0), self.power_coefficient
            a.dof_vel, a.dof_env_stride, a.dof_elem_stride = dof_vel.data_ptr(), dof_vel.stride(0), dof_vel.stride(1)
        self._launch(a)


class StrikeTaskB200(_ZTaskBase):
    """HumanoidStrike (humanoid_strike.py:17-240): walk to a standing target and knock it over."""
    kind, obs_size = _lib.ZTASK_STRIKE, STRIKE_OBS

    def __init__(self, num_envs: int, device="cuda:0", contact_bodies: Sequence[str] = ("R_Ankle", "L_Ankle", "R_Toe", "L_Toe"),
                 strike_bodies: Sequence[str] = ("R_Wrist", "R_Hand"), tar_dist_min: float = 0.5, tar_dist_max: float = 10.0, near_dist: float = 1.5,
                 near_prob: float = 0.5, max_episode_length: int = 300, enable_early_termination: bool = True, termination_height: float = 0.15,
                 dt: float = 1.0 / 30.0):
        super().__init__(num_envs, device, contact_bodies, max_episode_length, enable_early_termination, termination_height, dt)
        self.strike_body_mask = _mask(strike_bodies)
        self._tar_dist_min, self._tar_dist_max, self._near_dist, self._near_prob = tar_dist_min, tar_dist_max, near_dist, near_prob

    def get_task_obs_size(self) -> int:
        return 15

    def reset_target(self, env_ids: torch.Tensor, root_states: torch.Tensor, target_states: torch.Tensor, rand: Optional[torch.Tensor] = None) -> None:
        """_reset_target (humanoid_strike.py:124-145): place the target at a random distance / bearing around the character, upright,
        random yaw, at rest.  `target_states` is the [N, 13] view of the target actor's root state, written in place; `rand` [n, 4]
        injects the four uniform draws (near, distance, bearing, yaw)."""
        n = int(env_ids.shape[0])
        if n == 0:
            return
        r = torch.rand(n, 4, device=self.device) if rand is None else rand
        dist_max = torch.where(r[:, 0] < self._near_prob, torch.full_like(r[:, 0], self._near_dist), torch.full_like(r[:, 0], self._tar_dist_max))
        dist = (dist_max - self._tar_dist_min) * r[:, 1] + self._tar_dist_min
        theta, yaw = 2 * math.pi * r[:, 2], 2 * math.pi * r[:, 3]
        target_states[env_ids, 0] = dist * torch.cos(theta) + root_states[env_ids, 0]
        target_states[env_ids, 1] = dist * torch.sin(theta) + root_states[env_ids, 1]
        target_states[env_ids, 2] = 0.9
        zero = torch.zeros_like(yaw)
        target_states[env_ids, 3:7] = torch.stack([zero, zero, torch.sin(0.5 * yaw), torch.cos(0.5 * yaw)], dim=-1)   # quat_from_angle_axis(yaw, z)
        target_states[env_ids, 7:13] = 0.0

    def post_physics_step(self, rigid_body_state: torch.Tensor, progress_buf: torch.Tensor, target_states: torch.Tensor,
                          tar_contact_forces: torch.Tensor, contact_forces: Optional[torch.Tensor] = None) -> None:
        """_compute_reward (:176-185) + _compute_reset (:201-207) + _compute_observations in one launch.  target_states [N, 13] and
        tar_contact_forces [N, 3] are views of the simulator tensors (:109, :116), read through their env strides."""
        a = self._args(rigid_body_state, progress_buf, contact_forces)
        a.target_states, a.target_env_stride = target_states.data_ptr(), target_states.stride(0)
        a.tar_contact_forces, a.tar_contact_env_stride = tar_contact_forces.data_ptr(), tar_contact_forces.stride(0)
        self._launch(a)
