"""Per-irrep channel-mixing linear (stand-in for e3nn.o3.Linear).

Parameter container + packing: the flat ``weight`` has e3nn's layout (paths in i_in-major order, each
``(mul_in, mul_out)`` row-major, N(0,1) init, 1/sqrt(fan_in) applied at run time) so reference
checkpoints load unchanged.  Used by /root/reference/src/jamun/e3tools/nn/_interaction.py:23-24 and
_mlp.py:69,109.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import torch

from ...irreps import Irreps


class Linear(torch.nn.Module):
    def __init__(self, irreps_in, irreps_out):
        super().__init__()
        self.irreps_in = Irreps(irreps_in)
        self.irreps_out = Irreps(irreps_out)
        self.paths: List[Tuple[int, int, int, int, int]] = []  # (i_in, i_out, offset, mul_in, mul_out)
        off = 0
        for i_in, (mi, iri) in enumerate(self.irreps_in):
            for i_out, (mo, iro) in enumerate(self.irreps_out):
                if iri == iro:
                    self.paths.append((i_in, i_out, off, mi, mo))
                    off += mi * mo
        self.weight_numel = off
        self.weight = torch.nn.Parameter(torch.randn(off))

    def packed(self, l: int) -> torch.Tensor:
        """Dense [sum mul_in(l), sum mul_out(l)] matrix for irreps of degree l, 1/sqrt(fan_in) folded in.

        Rows/cols follow the order of the degree-l blocks in irreps_in / irreps_out."""
        in_blocks = [(i, m) for i, (m, ir) in enumerate(self.irreps_in) if ir.l == l]
        out_blocks = [(i, m) for i, (m, ir) in enumerate(self.irreps_out) if ir.l == l]
        rows = sum(m for _, m in in_blocks)
        cols = sum(m for _, m in out_blocks)
        W = self.weight.new_zeros(rows, cols)
        fan = {io: 0 for io, _ in out_blocks}
        for i_in, i_out, _, mi, _ in self.paths:
            if i_out in fan:
                fan[i_out] += mi
        r0 = {}
        o = 0
        for i, m in in_blocks:
            r0[i] = o
            o += m
        c0 = {}
        o = 0
        for i, m in out_blocks:
            c0[i] = o
            o += m
        for i_in, i_out, off, mi, mo in self.paths:
            if i_in in r0 and i_out in c0:
                blk = self.weight[off:off + mi * mo].reshape(mi, mo) / math.sqrt(fan[i_out])
                W[r0[i_in]:r0[i_in] + mi, c0[i_out]:c0[i_out] + mo] = blk
        return W.contiguous()

    def forward(self, x):
        """Module-level compatibility forward (e3nn layout in/out); the sampler evaluates these inside the fused block kernels."""
        from ... import ops

        n = x.shape[0]
        sl_in, outs = self.irreps_in.slices(), []
        ls = sorted({ir.l for _, ir in self.irreps_out})
        res = {}
        for l in ls:
            d = 2 * l + 1
            blocks = [x[:, sl].reshape(n, m, d) for sl, (m, ir) in zip(sl_in, self.irreps_in) if ir.l == l]
            cols = sum(m for m, ir in self.irreps_out if ir.l == l)
            if not blocks:
                res[l] = x.new_zeros(n, cols, d)
                continue
            xin = torch.cat(blocks, dim=1).permute(0, 2, 1).reshape(n * d, -1).contiguous().float()
            w = self.packed(l).detach().T.contiguous()
            res[l] = ops.linear_act(xin, w, None, act=0).reshape(n, d, cols).permute(0, 2, 1)
        off = {l: 0 for l in ls}
        for m, ir in self.irreps_out:
            outs.append(res[ir.l][:, off[ir.l]:off[ir.l] + m].reshape(n, -1))
            off[ir.l] += m
        return torch.cat(outs, dim=1)
