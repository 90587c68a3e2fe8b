"""Golden LOSSES and gradient samples of the PULSE VAE distillation minibatch at im_z_fit.yaml WIDTHS, produced by the
UNMODIFIED reference (build container only):

  python tests/golden/make_golden_vae_full.py      -> tests/golden/vae_full.npz   (small: scalars + gradient samples)

`AMPZBuilder.Network` (phc/learning/amp_network_z_builder.py:24-557) is built from `im_z_fit.yaml`'s network block AS IS
(encoder 934-1536-1024-512-160, prior 358-1536-1024-512, decoder 390-3096-2048-1024-69, embedding 32), its parameters are
overwritten with `tests.helpers.vae_full_fixture()` (integer-exact regeneration -- 16 M parameters cannot be committed),
and `AMPAgent._optimize_kin` (phc/learning/amp_agent.py:771-849) runs unbound on a stand-in agent with a recording
optimizer and the fixture's noise injected through the network's "z_noise" path.  Stored: every loss term, the first rows
of every weight gradient, every bias gradient, per-parameter gradient norms, sample rows of the forward outputs, and the
fixture checksum.  rl_games' ObjectFactory is restated [3P-memory] exactly as in make_golden_vae.py.
"""
import os
import sys
import types

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.refshim import load_reference as LR  # noqa: E402
from tests.golden.make_golden_vae import ObjectFactory, RecordingOptimizer, np_  # noqa: E402
from tests.helpers import VAE_FULL, vae_full_fixture  # noqa: E402


def main():
    ref = LR.load_reference()
    lrn = LR.load_learning()
    lrn.network_builder.object_factory = types.SimpleNamespace(ObjectFactory=ObjectFactory)
    ref.flags.trigger_input = False
    d = VAE_FULL
    S, Tk, A, E, T, NE = d["S"], d["Tk"], d["A"], d["E"], d["T"], d["NE"]
    B = T * NE
    params = yaml.safe_load(open(os.path.join(LR.REFERENCE_ROOT, "phc/data/cfg/learning/im_z_fit.yaml")))["params"]["network"]
    assert tuple(params["mlp"]["units"]) == d["dec_units"] and tuple(params["task_mlp"]["units"]) == d["task_units"], params
    builder = lrn.amp_network_z_builder.AMPZBuilder()
    builder.load(params)
    detail = {"proj_norm": True, "embedding_size": E, "embedding_norm": 1, "z_readout": False, "z_type": "vae", "use_vae_prior": True,
              "use_vae_clamped_prior": True, "vae_var_clamp_max": 2}       # env_im_vae.yaml:21-27
    ms = types.SimpleNamespace(running_mean=torch.zeros(S + Tk), running_var=torch.ones(S + Tk))
    net = builder.build("amp_z", actions_num=A, input_shape=(S + Tk,), num_seqs=1, value_size=1, amp_input_shape=(1960,), self_obs_size=S,
                        task_obs_size=Tk, task_obs_size_detail=detail, mean_std=ms)
    sd, batch, chk = vae_full_fixture()
    own = net.state_dict()
    with torch.no_grad():
        for k, v in sd.items():
            assert own[k].shape == v.shape, (k, own[k].shape, v.shape)
            own[k].copy_(v)
    net.train()
    obs, gt_action, noise, progress = batch["obs"], batch["gt_action"], batch["noise"], batch["progress"]
    kin_flat = torch.cat([gt_action, progress.float().view(B, 1)], dim=-1)
    out = {"checksum": torch.tensor(chk, dtype=torch.float64), "dims": torch.tensor([S, Tk, A, E, T, NE])}
    task = types.SimpleNamespace(distill=True, z_type="vae", use_vae_prior=True, use_vae_fixed_prior=False, use_ar1_prior=True,
                                 use_vae_prior_regu=False, kld_coefficient=0.01, kld_coefficient_min=0.001, kld_anneal=True, ar1_coefficient=0.005)
    AA = lrn.amp_agent.AMPAgent
    names = [n for n, _ in net.named_parameters()]
    for tag, regu in (("", False), ("regu_", True)):
        task.use_vae_prior_regu = regu
        opt = RecordingOptimizer(net.parameters())
        agent = types.SimpleNamespace(vec_env=types.SimpleNamespace(env=types.SimpleNamespace(task=task)),
                                      model=types.SimpleNamespace(a2c_network=net, parameters=net.parameters),
                                      kin_dict_info={"gt_action": ((B, A), (B, A)), "progress_buf": ((B,), (B, 1))},
                                      minibatch_size=B, horizon_length=T, kin_optimizer=opt, grad_norm=50.0, epoch_num=10)
        agent._assamble_kin_dict = types.MethodType(AA._assamble_kin_dict, agent)
        net.z_noise = noise.clone()      # what the rollout's eval_actor left behind; the "z_noise" path below re-uses the given draw (:89-90)
        info = AA._optimize_kin(agent, {"obs": obs.clone(), "kin_dict": kin_flat.clone(), "z_noise": noise.clone()})
        for k, v in info.items():
            out[tag + "info." + k] = v if torch.is_tensor(v) else torch.tensor(float(v))
        if not regu:
            for n, gr in zip(names, opt.grads):
                if gr is None:
                    continue
                out["gnorm." + n] = gr.double().norm().float()
                out["grad." + n] = gr[:4].clone() if gr.dim() == 2 else gr.clone()
    with torch.no_grad():
        mu, sigma, extra = net.eval_actor({"obs": obs, "z_noise": noise}, return_extra=True)
        pm, plv = net.compute_prior({"obs": obs})
    rows = slice(0, 64)
    out.update(pred_action=mu[rows], vae_mu=extra["vae_mu"][rows], vae_log_var=extra["vae_log_var"][rows], prior_mu=pm[rows], prior_log_var=plv[rows])
    path = os.path.join(HERE, "vae_full.npz")
    np.savez_compressed(path, **{k: np_(v) for k, v in out.items()})
    print("vae_full.npz", os.path.getsize(path) // 1024, "KiB;", len(out), "arrays; checksum", chk)
    for k in sorted(out):
        if "info." in k:
            print(" ", k, float(out[k]))


if __name__ == "__main__":
    main()
