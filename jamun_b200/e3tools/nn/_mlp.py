"""ScalarMLP / EquivariantMLPBlock / EquivariantMLP (mirror of /root/reference/src/jamun/e3tools/nn/_mlp.py:10-114)."""
from __future__ import annotations

from typing import Callable, Optional

import torch

from ...irreps import Irreps
from ._gate import Gate
from ._linear import Linear


class ScalarMLP(torch.nn.Sequential):
    def __init__(self, in_features: int, out_features: int, hidden_features, activation_layer=torch.nn.ReLU,
                 norm_layer: Optional[Callable] = None, dropout: float = 0.0, bias: bool = True):
        if norm_layer is not None or dropout != 0.0 or not bias:
            raise NotImplementedError("radial MLP variants with norm/dropout/no-bias are outside the kernels' scope")
        layers, d = [], in_features
        for hdim in hidden_features:
            layers += [torch.nn.Linear(d, hdim, bias=bias), activation_layer(), torch.nn.Dropout(dropout)]
            d = hdim
        layers += [torch.nn.Linear(d, out_features, bias=bias), torch.nn.Dropout(dropout)]
        super().__init__(*layers)


class EquivariantMLPBlock(torch.nn.Module):
    def __init__(self, irreps_in, irreps_out, act=None, act_gates=None, norm_layer=None):
        super().__init__()
        if norm_layer is not None:
            raise NotImplementedError("norm_layer is not instantiated on the default path (SURVEY #15)")
        self.irreps_in, self.irreps_out = Irreps(irreps_in), Irreps(irreps_out)
        self.gate = Gate(self.irreps_out, act=act, act_gates=act_gates)
        self.lin = Linear(self.irreps_in, self.gate.irreps_in)
        self.norm = None

    def forward(self, x):
        return self.gate(self.lin(x))


class EquivariantMLP(torch.nn.Sequential):
    def __init__(self, irreps_in, irreps_out, irreps_hidden_list, act=None, act_gates=None, norm_layer=None):
        layers, cur = [], Irreps(irreps_in)
        for hid in irreps_hidden_list:
            layers.append(EquivariantMLPBlock(cur, Irreps(hid), act=act, act_gates=act_gates, norm_layer=norm_layer))
            cur = Irreps(hid)
        layers.append(Linear(cur, irreps_out))
        super().__init__(*layers)
        self.irreps_in, self.irreps_out = Irreps(irreps_in), Irreps(irreps_out)
