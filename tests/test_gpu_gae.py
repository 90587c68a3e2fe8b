"""GPU parity of the GAE / return scan and advantage normalisation vs golden fixtures and oracle."""
import pytest
import torch

from tests.helpers import load_npz

pytestmark = pytest.mark.gpu


def test_gae_matches_reference_golden():
    from oracle import pulse_oracle as po
    from pulse_b200.rollout import discount_values
    z = load_npz("agent.npz")
    dev = torch.device("cuda:0")
    adv, ret = discount_values(z["fdones"].to(dev), z["values"].to(dev), z["rewards"].to(dev), z["next_values"].to(dev))
    torch.cuda.synchronize()
    T, N = z["fdones"].shape
    torch.testing.assert_close(adv.cpu(), po.swap_and_flatten01(z["advs"]).reshape(-1), atol=1e-6, rtol=1e-6)
    torch.testing.assert_close(ret.cpu(), po.swap_and_flatten01(z["returns"]).reshape(-1), atol=1e-6, rtol=1e-6)
    adv_n, _ = discount_values(z["fdones"].to(dev), z["values"].to(dev), z["rewards"].to(dev), z["next_values"].to(dev),
                               normalize_advantage=True)
    torch.testing.assert_close(adv_n.cpu(), z["adv_norm"], atol=2e-5, rtol=1e-5)


@pytest.mark.parametrize("T,N", [(32, 16384), (1, 5), (64, 33)])
def test_gae_matches_oracle_sizes(T, N):
    from oracle import pulse_oracle as po
    from pulse_b200.rollout import discount_values
    g = torch.Generator().manual_seed(T * 1000 + N)
    r, v, nv = (torch.randn(T, N, 1, generator=g) for _ in range(3))
    d = (torch.rand(T, N, generator=g) < 0.1).float()
    ref = po.discount_values(d, v, r, nv)
    dev = torch.device("cuda:0")
    adv, ret = discount_values(d.to(dev), v.to(dev), r.to(dev), nv.to(dev))
    torch.testing.assert_close(adv.cpu(), po.swap_and_flatten01(ref).reshape(-1), atol=1e-6, rtol=1e-6)
    torch.testing.assert_close(ret.cpu(), po.swap_and_flatten01(ref + v).reshape(-1), atol=1e-6, rtol=1e-6)
