"""The dataset side of the reference's MotionLib (`phc.utils.motion_lib_base.MotionLibBase.__init__`, `load_motions`,
`update_hard_sampling_weight`, `update_soft_sampling_weight`, `update_sampling_prob`; motion_lib_base.py:112-177, :179-323,
:346-384) on top of the device-side loader (`MotionLibB200.from_clips`): which clips are drawn for the envs,
with which heading, and how the PMCP sampling weights move.  Host logic only (a few vectors of length `num_unique_motions`);
the per-frame work happens in `pulse_motionlib_load_clips`.

Mirrors the SINGLE-PROCESS path of `load_motions` (`num_jobs = 1`: `mp.cpu_count() <= 8`, `multi_thread` off or `flags.debug`,
motion_lib_base.py:237-243): the worker re-seeds numpy with `np.random.randint(5000) * pid` = 0 and draws one heading per clip
(motion_lib_smpl.py:106, :131-140).  In the reference's multi-process path every worker chunk has its own seed, so the
headings (not their distribution) differ from this class.
"""
import random
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .motion_lib import MotionLibB200


class MotionDatasetB200:
    def __init__(self, motion_data: Dict[str, dict], parents: Sequence[int], local_translation, device="cuda:0", min_length: int = -1,
                 randomize_heading: bool = True):
        """motion_data: {key: clip} in the on-disk schema (what `joblib.load(motion_file)` returns, motion_lib_base.py:139-160);
        clips shorter than `min_length` frames are dropped like the reference does (:150-156)."""
        if min_length != -1:
            motion_data = {k: v for k, v in motion_data.items() if len(v["pose_quat_global"]) >= min_length}
        self._motion_data_keys = np.array(list(motion_data.keys()))
        self._motion_data_list = [motion_data[k] for k in self._motion_data_keys]
        self._num_unique_motions = len(self._motion_data_list)
        if self._num_unique_motions == 0:
            raise _lib.PulseError("no motion clips")
        self._device = torch.device(device)
        self._parents, self._local_translation = list(parents), np.asarray(local_translation, dtype=np.float32)
        self.randomize_heading = randomize_heading
        # host-side bookkeeping vectors (the reference keeps them on its device; they are read by multinomial only)
        self._sampling_prob = torch.ones(self._num_unique_motions) / self._num_unique_motions          # :170
        self._termination_history = torch.zeros(self._num_unique_motions)                                # :166
        self._success_rate = torch.zeros(self._num_unique_motions)
        self._sampling_history = torch.zeros(self._num_unique_motions)
        self._curr_motion_ids: Optional[torch.Tensor] = None
        self.curr_motion_keys = None

    # ------------------------------------------------------------------ PMCP sampling weights
    def update_hard_sampling_weight(self, failed_keys) -> None:
        """:346-358: train only on the failed sequences (uniform over them), or uniform over all when there are none."""
        if len(failed_keys) > 0:
            all_keys = self._motion_data_keys.tolist()
            indexes = [all_keys.index(k) for k in failed_keys]
            self._sampling_prob[:] = 0
            self._sampling_prob[indexes] = 1 / len(indexes)
        else:
            self._sampling_prob = torch.ones(self._num_unique_motions) / self._num_unique_motions

    def update_soft_sampling_weight(self, failed_keys) -> None:
        """:360-374: failure counts accumulate; the sampling probability is proportional to them."""
        if len(failed_keys) > 0:
            all_keys = self._motion_data_keys.tolist()
            indexes = [all_keys.index(k) for k in failed_keys]
            self._termination_history[indexes] += 1
            self.update_sampling_prob(self._termination_history)
        else:
            self._sampling_prob = torch.ones(self._num_unique_motions) / self._num_unique_motions

    def update_sampling_prob(self, termination_history: torch.Tensor) -> bool:
        """:376-383."""
        if len(termination_history) == len(self._termination_history) and termination_history.sum() > 0:
            self._sampling_prob[:] = termination_history / termination_history.sum()
            self._termination_history = termination_history
            return True
        return False

    # ------------------------------------------------------------------ which clips, which headings
    def select(self, num_motion_to_load: int, random_sample: bool = True, start_idx: int = 0) -> torch.Tensor:
        """:205-213: multinomial over the sampling probabilities, or consecutive clips from `start_idx` (evaluation)."""
        if random_sample:
            ids = torch.multinomial(self._sampling_prob, num_samples=num_motion_to_load, replacement=True)
        else:
            ids = torch.remainder(torch.arange(num_motion_to_load) + start_idx, self._num_unique_motions)
        self._curr_motion_ids = ids
        self.curr_motion_keys = self._motion_data_keys[ids.numpy()]
        self._sampling_batch_prob = self._sampling_prob[ids] / self._sampling_prob[ids].sum()
        return ids

    @staticmethod
    def draw_headings(n: int) -> np.ndarray:
        """One heading per clip exactly as the single-process worker draws them (motion_lib_smpl.py:106, :134-135)."""
        np.random.seed(np.random.randint(5000) * 0)
        return np.array([np.pi * (2 * np.random.random() - 1.0) for _ in range(n)])

    def crop(self, clip: dict, max_len: int) -> dict:
        """:118-123 of motion_lib_smpl.py: a random window of `max_len` frames when the clip is longer (python's `random`)."""
        seq_len = len(clip["pose_quat_global"])
        if max_len == -1 or seq_len < max_len:
            return clip
        start = random.randint(0, seq_len - max_len)
        sl = slice(start, start + max_len)
        return dict(clip, pose_quat_global=clip["pose_quat_global"][sl], pose_aa=clip["pose_aa"][sl], root_trans_offset=clip["root_trans_offset"][sl])

    def load_motions(self, num_motion_to_load: int, random_sample: bool = True, start_idx: int = 0, max_len: int = -1, eval_mode: bool = False) -> MotionLibB200:
        """`load_motions(skeleton_trees, gender_betas, limb_weights, random_sample, start_idx, max_len)` for one skeleton shape:
        draws the clips, crops, draws the headings (none in `flags.im_eval` / `flags.test` = eval_mode) and builds the tables on
        the device.  Returns the MotionLibB200 the step kernels read."""
        ids = self.select(num_motion_to_load, random_sample, start_idx)
        clips: List[dict] = [self.crop(self._motion_data_list[int(i)], max_len) for i in ids]
        headings = self.draw_headings(len(clips)) if (self.randomize_heading and not eval_mode) else None
        lib = MotionLibB200.from_clips(clips, self._parents, self._local_translation, self._device, headings=headings)
        lib._sampling_batch_prob = self._sampling_batch_prob.to(self._device)
        return lib
