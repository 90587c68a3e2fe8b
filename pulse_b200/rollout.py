"""Rollout post-processing on the device: GAE / returns / advantage normalisation.

Host-side mirror of `CommonAgent.discount_values` (phc/learning/common_agent.py:493-505),
`mb_returns = mb_advs + mb_values` (amp_agent.py:427) and `_calc_advs` (:589-599).
"""
import ctypes as C

import torch

from . import _lib


def discount_values(mb_fdones: torch.Tensor, mb_values: torch.Tensor, mb_rewards: torch.Tensor, mb_next_values: torch.Tensor,
                    gamma: float = 0.99, tau: float = 0.95, normalize_advantage: bool = False):
    """Inputs time-major [T,N] / [T,N,1] as in the rl_games ExperienceBuffer.

    Returns (advantages, returns) ENV-MAJOR flat [N*T] -- the `swap_and_flatten01` layout the PPO
    dataset slices minibatches from.  With normalize_advantage=True the advantages are additionally
    normalised over the whole batch ((A-mean)/(std+1e-8), unbiased std) as `_calc_advs` does.
    """
    lib = _lib.load()
    T, N = mb_fdones.shape[0], mb_fdones.shape[1]
    dev = mb_rewards.device
    flat = lambda x: x.reshape(T, N).to(torch.float32).contiguous()
    r, v, nv, d = flat(mb_rewards), flat(mb_values), flat(mb_next_values), flat(mb_fdones)
    adv = torch.empty(N * T, device=dev, dtype=torch.float32)
    ret = torch.empty(N * T, device=dev, dtype=torch.float32)
    stats = torch.zeros(2, device=dev, dtype=torch.float64)
    a = _lib.GaeArgs(rewards=r.data_ptr(), values=v.data_ptr(), next_values=nv.data_ptr(), fdones=d.data_ptr(), gamma=gamma, tau=tau,
                     advantages=adv.data_ptr(), returns=ret.data_ptr(), adv_sum=stats.data_ptr())
    with torch.cuda.device(dev):
        st = _lib.current_stream(dev)
        _lib.check(lib.pulse_gae(C.byref(a), T, N, st), "pulse_gae")
        if normalize_advantage:
            _lib.check(lib.pulse_normalize_advantages(adv.data_ptr(), stats.data_ptr(), N * T, st), "pulse_normalize_advantages")
    return adv, ret
