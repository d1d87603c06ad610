"""Closed-form stand-in for the E3Conv network, g(y_scaled, c_noise, r_cut) -> [N, 3] (test scaffolding).

tests/golden/make_reference_golden.py runs the reference's own Denoiser / walk-jump code around this function, the tests
run the oracle and this library's kernels around the same function; anything that differs is then in the code *around*
the network (normalisation, centring, cut-off, xhat / score tail, clip, integrator, jump, noise, Kabsch, loss).
It depends on every argument the real network receives, so a wrong c_in, c_noise or cut-off shows up in the output."""
import torch

TOY_W = torch.tensor([[0.7, -0.3, 0.2], [0.1, 0.9, -0.5], [-0.4, 0.25, 0.6]], dtype=torch.float32)


def toy_g(p: torch.Tensor, c_noise, r_cut) -> torch.Tensor:
    w = TOY_W.to(device=p.device, dtype=p.dtype)
    c = torch.as_tensor(c_noise).to(device=p.device, dtype=p.dtype).reshape(-1)[0]
    r = torch.as_tensor(r_cut).to(device=p.device, dtype=p.dtype)
    return torch.sin(p @ w) * (1 + c) + 0.1 * r * p.roll(1, -1)
