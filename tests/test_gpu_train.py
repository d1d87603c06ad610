"""Row a16: training forward + backward (library path, jamun_b200/train.py) against the oracle's autograd."""
import pytest
import torch

from conftest import make_oracle_batch

pytestmark = pytest.mark.gpu
SIGMA = 0.04


def _setup(models, sizes, seed=3):
    from jamun_b200 import data, synthetic

    o32, o64, prod = models
    t = synthetic.make_tensors(sizes)
    gen = torch.Generator().manual_seed(seed)
    y = t["pos"] + SIGMA * torch.randn(t["pos"].shape, generator=gen)
    batch = data.Batch.from_tensors(t).to("cuda")
    yb = batch.clone("pos")
    yb.pos = y.cuda()
    return o32, prod, t, y, batch, yb


def test_differentiable_forward_matches_kernel_path(models):
    o32, prod, t, y, batch, yb = _setup(models, [22, 15, 9, 30])
    prod.eval()
    with torch.no_grad():
        ref = prod.xhat(yb, SIGMA).pos
        got = prod.xhat_with_grad(yb, SIGMA).pos
    assert torch.allclose(got, ref, rtol=1e-4, atol=1e-5), (got - ref).abs().max()


def test_training_gradients_match_oracle_autograd(models):
    from oracle import jamun_oracle as O

    o32, prod, t, y, batch, yb = _setup(models, [22, 15, 9])
    ob = make_oracle_batch(t)
    o32.zero_grad()
    xh_o = o32.xhat(ob.with_pos(y), SIGMA)
    loss_o, _ = O.compute_loss(o32, ob, xh_o, SIGMA)
    loss_o.mean().backward()
    ref = {k: p.grad.clone() for k, p in o32.named_parameters() if p.grad is not None}
    o32.zero_grad()

    prod.train()
    prod.zero_grad()
    xh_p = prod.xhat_with_grad(yb, SIGMA)
    loss_p, aux = prod.compute_loss(batch, xh_p, SIGMA)
    loss_p.mean().backward()
    prod.eval()
    assert torch.allclose(loss_p.detach().cpu(), loss_o.detach(), rtol=2e-4, atol=1e-6)
    got = {k.replace("g._orig_mod.", "g."): p.grad.detach().cpu() for k, p in prod.named_parameters() if p.grad is not None}
    prod.zero_grad()
    assert len(ref) > 40 and set(ref) <= set(got), sorted(set(ref) - set(got))[:5]
    for k, g_ref in ref.items():
        scale = max(g_ref.abs().max().item(), 1e-12)
        err = (got[k] - g_ref).abs().max().item()
        assert err <= 2e-3 * scale + 1e-7, f"{k}: grad err {err} scale {scale}"


def test_training_step_runs_and_reduces_loss(models):
    """A few SGD steps through Denoiser.training_step on one fixed noisy batch lower the loss."""

    import jamun_b200 as J
    from jamun_b200 import data, synthetic

    o32, _, prod0 = models
    prod = J.default_denoiser()
    prod.load_state_dict(o32.state_dict())
    prod = prod.to("cuda").train()
    prod.add_fixed_noise = True  # same noise draw every step (denoiser.py:94-103)
    t = synthetic.make_tensors([20, 20, 20, 20])
    batch = data.Batch.from_tensors(t).to("cuda")
    opt = torch.optim.SGD(prod.parameters(), lr=1e-3)
    losses = []
    for it in range(4):
        opt.zero_grad()
        out = prod.training_step(batch, it)
        out["loss"].backward()
        opt.step()
        losses.append(float(out["loss"].detach()))
    assert all(torch.isfinite(torch.tensor(losses)))
    assert losses[-1] < losses[0], losses
