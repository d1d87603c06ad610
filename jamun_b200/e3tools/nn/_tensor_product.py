"""Instruction table of e3nn.o3.FullyConnectedTensorProduct(shared_weights=False, internal_weights=False)
for l <= 1; the contraction itself runs inside jamun_conv_fwd."""
from __future__ import annotations

from typing import List, Tuple

import torch

from ...irreps import Irreps


class FullyConnectedTensorProduct(torch.nn.Module):
    def __init__(self, irreps_in1, irreps_in2, irreps_out, shared_weights: bool = False, internal_weights: bool = False):
        super().__init__()
        if shared_weights or internal_weights:
            raise NotImplementedError("only per-edge external weights are on the walk-jump path")
        self.irreps_in1, self.irreps_in2, self.irreps_out = Irreps(irreps_in1), Irreps(irreps_in2), Irreps(irreps_out)
        self.instructions: List[Tuple[int, int, int, int]] = []  # (i1, i2, i_out, weight offset)
        off = 0
        for i1, (m1, ir1) in enumerate(self.irreps_in1):
            for i2, (m2, ir2) in enumerate(self.irreps_in2):
                for io, (mo, iro) in enumerate(self.irreps_out):
                    if iro in ir1 * ir2:
                        self.instructions.append((i1, i2, io, off))
                        off += m1 * m2 * mo
        self.weight_numel = off

    def forward(self, x1, x2, weight):
        raise NotImplementedError("evaluated inside jamun_conv_fwd")
