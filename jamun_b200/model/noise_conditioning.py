"""NoiseConditionalScaling / NoiseConditionalSkipConnection parameter containers
(mirror of /root/reference/src/jamun/model/noise_conditioning.py:27-73).  At sampling time their outputs are
per-(model, sigma) constants; jamun_noise_mlp evaluates them once per plan."""
from __future__ import annotations

import torch
import torch.nn as nn

from ..irreps import Irreps


class NoiseConditionalScaling(nn.Module):
    def __init__(self, irreps_in, noise_input_dims: int = 1, num_layers: int = 1):
        super().__init__()
        if noise_input_dims != 1 or num_layers != 1:
            raise NotImplementedError("only the default Linear(1,n)-SELU-Linear(n,n) scale predictor is built")
        self.irreps_in = self.irreps_out = Irreps(irreps_in)
        n = self.irreps_in.num_irreps
        self.scale_predictor = nn.Sequential(nn.Linear(1, n), nn.SELU(), nn.Linear(n, n))
        with torch.no_grad():
            self.scale_predictor[-1].weight.fill_(0.0)
            self.scale_predictor[-1].bias.fill_(1.0)

    def _expand(self, scales):
        reps = torch.tensor([ir.dim for m, ir in self.irreps_in for _ in range(m)], device=scales.device)
        return torch.repeat_interleave(scales, reps, dim=-1)

    def scales(self, c_noise, sigmoid: bool = False):
        from .. import ops

        dev = self.scale_predictor[0].weight.device
        ops_ = [t.detach() for t in self.mlp_operands()]
        return ops.noise_mlp(*ops_, float(torch.as_tensor(c_noise).reshape(-1)[0]), sigmoid).to(dev)

    def forward(self, x, c_noise):
        """Module-level compatibility forward: per-irrep scales from jamun_noise_mlp, broadcast multiply."""
        return x * self._expand(self.scales(c_noise))

    def mlp_operands(self):
        l0, l2 = self.scale_predictor[0], self.scale_predictor[2]
        return l0.weight.reshape(-1).contiguous(), l0.bias.contiguous(), l2.weight.contiguous(), l2.bias.contiguous()


class NoiseConditionalSkipConnection(nn.Module):
    def __init__(self, irreps_in, noise_input_dims: int = 1):
        super().__init__()
        self.weights = NoiseConditionalScaling(irreps_in, noise_input_dims=noise_input_dims)
        self.irreps_in = self.irreps_out = Irreps(irreps_in)

    def forward(self, x1, x2, c_noise):
        w = self.weights._expand(self.weights.scales(c_noise, sigmoid=True))
        return x1 * w + x2 * (1 - w)
