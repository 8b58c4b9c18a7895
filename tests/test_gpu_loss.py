"""GPU parity: balanced seeding loss forward/backward vs the numpy restatement (float32 result,
rtol 1e-5 -- Theano's float32 reduction order is unspecified, pylayers.py:136-139)."""
import numpy as np
import pytest

from dsrg_b200 import api, synth
from oracle import loss_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("H,W", [(41, 41), (64, 33)])
def test_seedloss_forward_backward(torch_cuda, H, W):
    torch = torch_cuda
    B, M = 4, 21
    batch = synth.make_batch(B, H, W, cues="random", start=80)
    probs = batch["probs"].copy()
    probs[probs < 1e-4] = 1e-4
    seeds = batch["cues"].copy()
    seeds[1, 1:] = 0   # image without foreground seeds -> max(count, 1e-4)
    seeds[2, 0] = 0    # image without background seeds
    eng = api.Engine(B, H, W, M)
    d_p, d_s = torch.from_numpy(probs).cuda(), torch.from_numpy(seeds).cuda()
    terms = torch.zeros(2, device="cuda")
    eng.seedloss_forward_dev(d_p, d_s, terms)
    loss = -float(terms.sum().item()) / B
    want = loss_oracle.balanced_seed_loss(probs, seeds)
    assert abs(loss - want) <= 1e-5 * abs(want)
    grad = torch.empty_like(d_p)
    eng.seedloss_backward_dev(d_p, d_s, grad)
    wg = loss_oracle.balanced_seed_loss_grad(probs, seeds)
    np.testing.assert_allclose(grad.cpu().numpy(), wg, rtol=1e-5, atol=1e-9)
    eng.close()
