"""Conv / ConvBlock (mirror of /root/reference/src/jamun/e3tools/nn/_conv.py:15-221).

These classes own the parameters under the reference's names and expose ``pack()``: the re-layout of
``radial_nn`` into the operands of the aggregate-then-transform kernels ("conv math").
"""
from __future__ import annotations

import functools
import math
from typing import Callable, Optional

import torch

from ...irreps import Irrep, Irreps
from ._gate import Gated
from ._interaction import LinearSelfInteraction
from ._mlp import ScalarMLP
from ._tensor_product import FullyConnectedTensorProduct


class Conv(torch.nn.Module):
    def __init__(self, irreps_in, irreps_out, irreps_sh, edge_attr_dim: int,
                 radial_nn: Optional[Callable[..., torch.nn.Module]] = None,
                 tensor_product: Optional[Callable[..., torch.nn.Module]] = None):
        super().__init__()
        self.irreps_in, self.irreps_out, self.irreps_sh = Irreps(irreps_in), Irreps(irreps_out), Irreps(irreps_sh)
        self.edge_attr_dim = edge_attr_dim
        if tensor_product is None:
            tensor_product = functools.partial(FullyConnectedTensorProduct, shared_weights=False, internal_weights=False)
        self.tp = tensor_product(self.irreps_in, self.irreps_sh, self.irreps_out)
        if radial_nn is None:
            radial_nn = functools.partial(ScalarMLP, hidden_features=[edge_attr_dim], activation_layer=torch.nn.SiLU)
        self.radial_nn = radial_nn(edge_attr_dim, self.tp.weight_numel)

    def pack(self, embed_bondedness: torch.Tensor):
        """-> dict(w0r [32,64] (k-major), b0eff [2,64], m0 [65,U0,152], m1 [65,U1,32], alpha0, alpha1, s_in, v_in)."""
        if not isinstance(self.tp, FullyConnectedTensorProduct) or not isinstance(self.radial_nn, ScalarMLP):
            raise NotImplementedError("the B200 conv kernel is built for FullyConnectedTensorProduct + ScalarMLP")
        if self.irreps_sh != Irreps("1x0e+1x1e") or self.irreps_out.simplify() != Irreps("152x0e+32x1e"):
            raise NotImplementedError(f"conv irreps sh={self.irreps_sh} out={self.irreps_out} outside kernel scope")
        lins = [m for m in self.radial_nn if isinstance(m, torch.nn.Linear)]
        acts = [m for m in self.radial_nn if not isinstance(m, (torch.nn.Linear, torch.nn.Dropout))]
        if len(lins) != 2 or len(acts) != 1 or not isinstance(acts[0], torch.nn.SiLU) or lins[0].out_features != 64:
            raise NotImplementedError("radial MLP must be Linear(64,64)-SiLU-Linear(64,P)")
        s_in, v_in = self.irreps_in.scalars_vectors()
        nb = embed_bondedness.shape[1]
        W0, b0 = lins[0].weight, lins[0].bias
        w0r = W0[:, nb:].T.contiguous()
        b0eff = (b0[None, :] + embed_bondedness @ W0[:, :nb].T).contiguous()
        Mfull = torch.cat([lins[1].weight.T, lins[1].bias[None, :]], dim=0)  # [65, P]
        K = Mfull.shape[0]
        e0, e1 = Irrep(0, 1), Irrep(1, 1)
        io0 = [i for i, (_, ir) in enumerate(self.irreps_out) if ir == e0]
        io1 = [i for i, (_, ir) in enumerate(self.irreps_out) if ir == e1]
        if len(io0) != 1 or len(io1) != 1:
            raise NotImplementedError("conv output must be one 0e block and one 1e block")
        mo0, mo1 = self.irreps_out[io0[0]][0], self.irreps_out[io1[0]][0]
        blk = {}
        for i1, i2, io, off in self.tp.instructions:
            m1 = self.irreps_in1_mul(i1)
            mo = self.irreps_out[io][0]
            blk[(i1, self.irreps_sh[i2][1].l, self.irreps_out[io][1].l)] = Mfull[:, off:off + m1 * mo].reshape(K, m1, mo)
        sc = [i for i, (_, ir) in enumerate(self.irreps_in) if ir == e0]
        vc = [i for i, (_, ir) in enumerate(self.irreps_in) if ir == e1]
        m0 = torch.cat([blk[(i, 0, 0)] for i in sc] + [blk[(i, 1, 0)] for i in vc], dim=1).contiguous()
        m1 = torch.cat([blk[(i, 1, 1)] for i in sc] + [blk[(i, 0, 1)] for i in vc] + [blk[(i, 1, 1)] for i in vc],
                       dim=1).contiguous()
        assert m0.shape == (K, s_in + v_in, mo0) and m1.shape == (K, s_in + 2 * v_in, mo1)
        return dict(w0r=w0r, b0eff=b0eff, m0=m0, m1=m1, alpha0=math.sqrt(1.0 / (s_in + v_in)),
                    alpha1=math.sqrt(3.0 / (s_in + 2 * v_in)), s_in=s_in, v_in=v_in)

    def irreps_in1_mul(self, i1: int) -> int:
        return self.irreps_in[i1][0]

    # ---- module-level forwards with the reference's signatures (compatibility path; the sampler uses engine.conv_tc).
    # Host-side glue (edge sort, layout permutations) is torch; the arithmetic runs in jamun_linear_act + jamun_conv_fwd.
    def _run(self, x_nodes, src, dst, n_recv, edge_attr, edge_sh):
        from ... import ops

        dev = x_nodes.device
        pk = self.pack(torch.zeros(2, self.edge_attr_dim // 2, device=dev))
        lins = [m for m in self.radial_nn if isinstance(m, torch.nn.Linear)]
        order = torch.sort(dst, stable=True).indices
        rowptr = torch.zeros(n_recv + 1, dtype=torch.long, device=dev)
        rowptr[1:] = torch.cumsum(torch.bincount(dst, minlength=n_recv), 0)
        col = src[order].to(torch.int32).contiguous()
        h = ops.linear_act(edge_attr[order].contiguous().float(), lins[0].weight.detach().contiguous(),
                           lins[0].bias.detach().contiguous(), act=1)
        sh = edge_sh[order].float()
        rhat = torch.zeros(sh.shape[0], 4, device=dev)
        rhat[:, :3] = sh[:, 1:4] / math.sqrt(3.0)
        s_in, v_in = pk["s_in"], pk["v_in"]
        xs = ops.layout_to_soa(x_nodes.contiguous().float(), s_in, v_in) if v_in else x_nodes.contiguous().float()
        out = torch.empty(n_recv, 248, device=dev)
        ops.conv_fwd(xs, s_in, v_in, rowptr.to(torch.int32), col, h, rhat.contiguous(), pk["m0"].detach().contiguous(),
                     pk["m1"].detach().contiguous(), pk["alpha0"], pk["alpha1"], out)
        return ops.layout_from_soa(out, 152, 32)

    def apply_per_edge(self, node_attr_src, edge_attr, edge_sh):
        """Per-edge messages tp(x_src, sh, radial_nn(edge_attr)) [E, 248] (each edge is its own receiver)."""
        E = node_attr_src.shape[0]
        idx = torch.arange(E, device=node_attr_src.device)
        return self._run(node_attr_src, idx, idx, E, edge_attr, edge_sh)

    def forward(self, node_attr, edge_index, edge_attr, edge_sh):
        """[N, irreps_in.dim], [2,E], [E, edge_attr_dim], [E, 4] -> scatter-mean of the messages, [N, 248] (e3nn layout)."""
        src, dst = edge_index
        return self._run(node_attr, src, dst, node_attr.shape[0], edge_attr, edge_sh)


class ConvBlock(torch.nn.Module):
    def __init__(self, irreps_in, irreps_out, irreps_sh, edge_attr_dim: int, act=None, act_gates=None,
                 conv: Optional[Callable[..., torch.nn.Module]] = None):
        super().__init__()
        self.irreps_in, self.irreps_out, self.irreps_sh = Irreps(irreps_in), Irreps(irreps_out), Irreps(irreps_sh)
        if conv is None:
            conv = Conv
        wrapped_conv = functools.partial(conv, irreps_sh=irreps_sh, edge_attr_dim=edge_attr_dim)
        self.gated_conv = LinearSelfInteraction(
            Gated(wrapped_conv, irreps_in=self.irreps_in, irreps_out=self.irreps_out, act=act, act_gates=act_gates))

    @property
    def conv(self) -> Conv:
        return self.gated_conv.f.f

    def forward(self, node_attr, edge_index, edge_attr, edge_sh):
        return self.gated_conv(node_attr, edge_index, edge_attr, edge_sh)

    def pack(self, embed_bondedness: torch.Tensor):
        """Operands of jamun_conv_fwd + jamun_block_tail for this block."""
        if self.irreps_out.simplify() != Irreps("120x0e+32x1e"):
            raise NotImplementedError(f"hidden irreps {self.irreps_out} outside kernel scope (120x0e+32x1e)")
        out = self.conv.pack(embed_bondedness)
        lsi = self.gated_conv
        out.update(wself_s=lsi.self_interaction.packed(0), wself_v=lsi.self_interaction.packed(1),
                   wskip_s=lsi.skip_connection.packed(0))
        out["wskip_v"] = lsi.skip_connection.packed(1) if out["v_in"] > 0 else None
        gate = lsi.f.gate
        out.update(c_act=gate.c_act, c_gate=gate.c_gate)
        return out
