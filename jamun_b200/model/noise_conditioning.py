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

    def mlp_operands(self):
        l0, l2 = self.scale_predictor[0], self.scale_predictor[2]
        return l0.weight.reshape(-1).contiguous(), l0.bias.contiguous(), l2.weight.contiguous(), l2.bias.contiguous()


class NoiseConditionalSkipConnection(nn.Module):
    def __init__(self, irreps_in, noise_input_dims: int = 1):
        super().__init__()
        self.weights = NoiseConditionalScaling(irreps_in, noise_input_dims=noise_input_dims)
        self.irreps_in = self.irreps_out = Irreps(irreps_in)
