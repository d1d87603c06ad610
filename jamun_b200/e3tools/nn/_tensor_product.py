"""e3nn.o3.FullyConnectedTensorProduct(shared_weights=False, internal_weights=False) for l <= 1.

The instruction table fixes the layout of ``radial_nn``'s output (``weight_numel``, per-instruction offsets) that
``Conv.pack`` re-lays out for the aggregate-then-transform kernels; the sampling path never materialises per-edge weights.
``forward(x1, x2, weight)`` is the reference's seam ``tp(x_src, sh, weight)`` (/root/reference/src/jamun/e3tools/nn/_conv.py:94)
and runs in ``jamun_tensor_product`` (one CTA per row)."""
from __future__ import annotations

import math
import struct
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
        self._table = {}

    def instruction_table(self, device) -> torch.Tensor:
        """[n_instr, 12] int32 records of jamun_tensor_product (offsets, multiplicities, l's, weight offset, path coefficient)."""
        key = str(device)
        if key not in self._table:
            if max(self.irreps_in1.lmax, self.irreps_in2.lmax, self.irreps_out.lmax) > 1:
                raise NotImplementedError("jamun_tensor_product evaluates l <= 1 only")
            s1, s2, so = self.irreps_in1.slices(), self.irreps_in2.slices(), self.irreps_out.slices()
            fan = [0] * len(self.irreps_out)
            for i1, i2, io, _ in self.instructions:
                fan[io] += self.irreps_in1[i1][0] * self.irreps_in2[i2][0]
            rows = []
            for i1, i2, io, off in self.instructions:
                (m1, ir1), (m2, ir2), (mo, iro) = self.irreps_in1[i1], self.irreps_in2[i2], self.irreps_out[io]
                coeff = math.sqrt(iro.dim / fan[io])
                bits = struct.unpack("<i", struct.pack("<f", coeff))[0]
                rows.append([s1[i1].start, m1, ir1.l, s2[i2].start, m2, ir2.l, so[io].start, mo, iro.l, off, bits, 0])
            self._table[key] = torch.tensor(rows, dtype=torch.int32).reshape(-1, 12).to(device)
        return self._table[key]

    def forward(self, x1, x2, weight):
        """x1 [Z, irreps_in1.dim], x2 [Z, irreps_in2.dim], weight [Z, weight_numel] (e3nn layouts) -> [Z, irreps_out.dim]."""
        from ... import ops

        if weight.shape[-1] != self.weight_numel:
            raise ValueError(f"weight has {weight.shape[-1]} columns, expected weight_numel={self.weight_numel}")
        return ops.tensor_product(x1.float().contiguous(), x2.float().contiguous(), weight.float().contiguous(),
                                  self.instruction_table(x1.device), self.irreps_out.dim)
