"""LinearSelfInteraction (mirror of /root/reference/src/jamun/e3tools/nn/_interaction.py:5-30)."""
from __future__ import annotations

import torch

from ._linear import Linear


class LinearSelfInteraction(torch.nn.Module):
    def __init__(self, f):
        super().__init__()
        self.f = f
        self.irreps_in, self.irreps_out = f.irreps_in, f.irreps_out
        self.skip_connection = Linear(self.irreps_in, self.irreps_out)
        self.self_interaction = Linear(self.irreps_out, self.irreps_out)

    def forward(self, x, *args):
        s = self.skip_connection(x)
        x = self.f(x, *args)
        x = self.self_interaction(x)
        return x + s
